// ekf_coop.cuh -- the cooperative mapping: THREE lanes per filter.
//
// Why.  With one thread per filter the packed covariance (960 B in FP64) caps an SM at 224 filters = 7 warps,
// 1.75 per scheduler, and nothing overlaps the FP64 pipe's idle stretches (profiles/r1_07_final_sr.md: 28 % issue
// slots used).  Here a filter is advanced by three lanes -- the same lane index of three consecutive warps (a
// "group" of 96 threads owns 32 filters) -- so an SM runs 3x the warps on the same covariance storage, every lane's
// dependency chains are a third as long, and shared-memory accesses stay conflict-free (a warp still touches one
// element of 32 consecutive filters).
//
// How the work splits.  The error state is five 3-vectors (dr dv dth dab dwb), the covariance 5x5 blocks of 3x3.
// Lane c (0,1,2) owns COLUMN c of every block: a row operation P[X,.] <- sum_Z F[X,Z] P[Z,.] (what each factor of
// F = E3 E2 E1 does, relative_pose_EKF.cpp:412-414) acts on columns independently, so each lane applies it to its
// own 5 columns with no communication; only the 3x3 diagonal block of the row being updated needs a transposed
// view, which goes through a 6-word exchange buffer.  Phases are separated by a named barrier of the group.
//
// Rotated frame.  Lane c relabels the axes of EVERY frame cyclically, local index a <-> global (a + c) mod 3.  A
// cyclic relabelling is a proper rotation, so every vector / rotation-matrix / cross-product / quaternion
// identity of the filter holds verbatim in local indices; all three lanes run the same instruction stream and
// "my column" is always local column 0.  Only loads, stores and exchanges translate indices.
//
// Private diagonal.  Element (c,c) of each 3x3 block is touched by lane c only (it is in column c and, mirrored,
// in row c).  Those 15 values live in the lane's registers; shared memory holds the other 75 (600 B per filter)
// and every column access costs two shared-memory words instead of three.
//
// Roles.  The nominal state is not replicated: role 0 (lane c = 0, whose local frame is the global one) holds it,
// runs the nominal kinematics of a tick and publishes the Jacobian pieces (B = -dT C, the specific force a, dT w and
// the coefficients of Phi) through a double-buffered exchange; meanwhile roles 1 and 2 draw the next tick's noisy
// IMU sample (Philox + Box-Muller), so neither job is replicated and neither serialises the other.  Roles are
// whole warps, so the split costs no divergence.
//
// Addressing.  Word (a,b), a != b, of a 3x3 block lives at slot ((a-b) mod 3 - 1)*3 + b, word {a,b} of a symmetric
// diagonal block at slot 3-a-b.  With that numbering every access of lane c is one of three base pointers
// (word c, word (c+1)%3, word (c+2)%3 of the filter) plus a compile-time offset.
//
// Nothing here is a port of the reference: what is computed is prediction_step / correction_step
// (relative_pose_EKF.cpp:346-502); how it is computed is specific to this mapping.
#pragma once

#include "ekf_kernels.cuh"

namespace qekf {
namespace coop {

// ------------------------------------------------------------------------------------------------
// storage layout of one filter's shared words (units of T; word e of filter f lives at base[e*S + f])
// ------------------------------------------------------------------------------------------------
QEKF_FN constexpr int odslot(int a, int b) { return ((a - b + 3) % 3 - 1) * 3 + b; }   // entry (a,b), a != b, of a 3x3 block
QEKF_FN constexpr int dgslot(int a, int b) { return 3 - a - b; }                        // entry {a,b}, a != b, of a symmetric 3x3
QEKF_FN constexpr int fullslot(int a, int b) { return ((a - b + 3) % 3) * 3 + b; }      // entry (a,b) of a full 3x3 (9 words)

template <int NB> QEKF_FN constexpr int bidx(int X, int Y) { return X * NB - (X * (X - 1)) / 2 + (Y - X); }   // X <= Y
template <int NB> QEKF_FN constexpr int sbase(int X, int Y)   // X <= Y: first shared word of block (X,Y)
{
    int o = 0;
    for (int x = 0; x < NB; ++x)
        for (int y = x; y < NB; ++y) {
            if (x == X && y == Y) return o;
            o += (x == y) ? 3 : 6;
        }
    return o;
}
// blocks of the dr / dth block columns, whose private entries are published before a correction
template <int NB> QEKF_FN constexpr int pubidx(int X, int Y)
{
    if (X == BR) return Y;                                   // (r,r) (r,v) (r,th) (r,ab) (r,wb)
    if (X == BV && Y == BTH) return NB;
    if (X == BTH) return NB + 1 + (Y - BTH);                 // (th,th) (th,ab) (th,wb)
    return -1;
}

template <int NB> struct Lay {
    static constexpr int NPRIV = NB * (NB + 1) / 2;          // 15 / 6 private (register) entries per lane
    static constexpr int NSH = 3 * NB + 3 * NB * (NB - 1);   // 75 / 27 shared covariance words
    static constexpr int UB = NSH;                           // noisy IMU samples [2 ticks][6], global order
    static constexpr int JB = UB + 12;                       // kinematics exchange [2 ticks][JB_N]
    static constexpr int JB_B = 0, JB_A = 9, JB_DTH = 12, JB_PC = 15, JB_N = 18;   // B (fullslot), a, dT w, (cs, s1, s2)
    static constexpr int XV = JB + 2 * JB_N;                 // exchange of the (v,v) row update, 6 words (odslot)
    static constexpr int XT = XV + 6;                        // exchange of Phi*P(th,th)
    static constexpr int CB = XT + 6;                        // ---- words used by corrections / statistics only ----
    static constexpr int NPUB = (NB == 5) ? 9 : 5;
    static constexpr int PB = CB;                            // published private entries [NPUB][3 lanes]
    static constexpr int DX = PB + 3 * NPUB;                 // injected error state [NB][3 lanes]
    static constexpr int QR = DX + 3 * NB;                   // nominal attitude (vector part [3], w) and position [3], global order
    static constexpr int SLOTS = QR + 7;                     // 190 / 121
};

// Shared-memory words are addressed through SPtr: on the device a 32-bit shared-window address used by explicit
// ld.shared / st.shared (generic pointers cost 64-bit address arithmetic and, where the compiler cannot prove the
// address space, generic loads); on the host a plain pointer whose accesses the test harness can trace (race
// detection between the three lanes of a filter).
#ifdef __CUDA_ARCH__
template <typename T> struct SPtr {
    uint32_t a;
    __device__ __forceinline__ SPtr operator+(int n) const { return SPtr{ a + (uint32_t)(n * (int)sizeof(T)) }; }
    static __device__ __forceinline__ SPtr from(T *g) { return SPtr{ (uint32_t)__cvta_generic_to_shared(g) }; }
    __device__ __forceinline__ T *generic() const { return (T *)__cvta_shared_to_generic((size_t)a); }
};
__device__ __forceinline__ double sm_ld(SPtr<double> p)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(p.a));
    return v;
}
__device__ __forceinline__ float sm_ld(SPtr<float> p)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(p.a));
    return v;
}
__device__ __forceinline__ void sm_st(SPtr<double> p, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(p.a), "d"(v)); }
__device__ __forceinline__ void sm_st(SPtr<float> p, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(p.a), "f"(v)); }
#else
template <typename T> struct SPtr {
    T *p;
    SPtr operator+(int n) const { return SPtr{ p + n }; }
    static SPtr from(T *g) { return SPtr{ g }; }
    T *generic() const { return p; }
};
#if defined(QEKF_COOP_TRACE)
void coop_trace(const void *addr, int is_write);
template <typename T> inline T sm_ld(SPtr<T> p) { coop_trace(p.p, 0); return *p.p; }
template <typename T> inline void sm_st(SPtr<T> p, T v) { coop_trace(p.p, 1); *p.p = v; }
#else
template <typename T> inline T sm_ld(SPtr<T> p) { return *p.p; }
template <typename T> inline void sm_st(SPtr<T> p, T v) { *p.p = v; }
#endif
#endif

// One lane's view of its filter's covariance.
template <typename T, int NB_, int S_> struct PLane {
    using real = T;
    static constexpr int nb = NB_, stride = S_;
    using L = Lay<NB_>;
    T priv[L::NPRIV];             // entry (c,c) of every block, X <= Y at bidx(X,Y)
    SPtr<T> sh;                   // word 0 of this filter
    SPtr<T> pc, pa, pb;           // words c, i1 = (c+1)%3, i2 = (c+2)%3 of this filter
    int c;
    bool wc1, wc2, wr1, wr2;      // who stores a diagonal block's shared entries: column form (i < c), row form (i > c)

    QEKF_FN void setup(T *filter_base, int c_)
    {
        c = c_;
        const int i1 = (c_ + 1) % 3, i2 = (c_ + 2) % 3;
        sh = SPtr<T>::from(filter_base);
        pc = sh + c * S_; pa = sh + i1 * S_; pb = sh + i2 * S_;
        wc1 = i1 < c; wc2 = i2 < c; wr1 = i1 > c; wr2 = i2 > c;
    }
    QEKF_FN int gi(int a) const { return (c + a) % 3; }   // local -> global component (cold paths only)
};

// column c of block (X,Y), i.e. P[3X + (a+c)%3, 3Y + c] for local a = 0,1,2 -- whatever the storage orientation
template <int X, int Y, class PL> QEKF_FN void col_ld(const PL &P, typename PL::real v[3])
{
    constexpr int NB = PL::nb, S = PL::stride;
    if constexpr (X < Y) {
        constexpr int b = sbase<NB>(X, Y) * S;
        v[0] = P.priv[bidx<NB>(X, Y)]; v[1] = sm_ld(P.pc + b); v[2] = sm_ld(P.pc + (b + 3 * S));
    } else if constexpr (X > Y) {
        constexpr int b = sbase<NB>(Y, X) * S;
        v[0] = P.priv[bidx<NB>(Y, X)]; v[1] = sm_ld(P.pa + (b + 3 * S)); v[2] = sm_ld(P.pb + b);
    } else {
        constexpr int b = sbase<NB>(X, X) * S;
        v[0] = P.priv[bidx<NB>(X, X)]; v[1] = sm_ld(P.pb + b); v[2] = sm_ld(P.pa + b);
    }
}
// store column c of block (X,Y); of a diagonal block only the upper entries (i < c) -- the others belong to
// the columns of the other lanes
template <int X, int Y, class PL> QEKF_FN void col_st(PL &P, const typename PL::real v[3])
{
    constexpr int NB = PL::nb, S = PL::stride;
    if constexpr (X < Y) {
        constexpr int b = sbase<NB>(X, Y) * S;
        P.priv[bidx<NB>(X, Y)] = v[0]; sm_st(P.pc + b, v[1]); sm_st(P.pc + (b + 3 * S), v[2]);
    } else if constexpr (X > Y) {
        constexpr int b = sbase<NB>(Y, X) * S;
        P.priv[bidx<NB>(Y, X)] = v[0]; sm_st(P.pa + (b + 3 * S), v[1]); sm_st(P.pb + b, v[2]);
    } else {
        constexpr int b = sbase<NB>(X, X) * S;
        P.priv[bidx<NB>(X, X)] = v[0];
        if (P.wc1) sm_st(P.pb + b, v[1]);
        if (P.wc2) sm_st(P.pa + b, v[2]);
    }
}
// store ROW c of diagonal block (X,X): v[k] = P[3X + c, 3X + (k+c)%3]; upper entries (i > c) only
template <int X, class PL> QEKF_FN void row_st(PL &P, const typename PL::real v[3])
{
    constexpr int NB = PL::nb, S = PL::stride;
    constexpr int b = sbase<NB>(X, X) * S;
    P.priv[bidx<NB>(X, X)] = v[0];
    if (P.wr1) sm_st(P.pb + b, v[1]);
    if (P.wr2) sm_st(P.pa + b, v[2]);
}
// a whole block (X <= Y) of the dr / dth block columns in local indices, b[a*3+k] = P[3X+gi(a), 3Y+gi(k)]; the
// private entries of the other two lanes come from the published copies (corr_publish)
template <int X, int Y, class PL> QEKF_FN void full_ld(const PL &P, typename PL::real b[9])
{
    constexpr int NB = PL::nb, S = PL::stride;
    using L = typename PL::L;
    constexpr int sb = sbase<NB>(X, Y) * S;
    constexpr int pw = (L::PB + 3 * pubidx<NB>(X, Y)) * S;
    static_assert(pubidx<NB>(X, Y) >= 0, "block is not published");
    b[0] = P.priv[bidx<NB>(X, Y)];
    b[4] = sm_ld(P.pa + pw);
    b[8] = sm_ld(P.pb + pw);
    if constexpr (X < Y) {
        b[3] = sm_ld(P.pc + sb); b[6] = sm_ld(P.pc + (sb + 3 * S));
        b[1] = sm_ld(P.pa + (sb + 3 * S)); b[2] = sm_ld(P.pb + sb);
        b[5] = sm_ld(P.pb + (sb + 3 * S));
        b[7] = sm_ld(P.pa + sb);
    } else {
        b[1] = b[3] = sm_ld(P.pb + sb);
        b[2] = b[6] = sm_ld(P.pa + sb);
        b[5] = b[7] = sm_ld(P.pc + sb);
    }
}
// a 3-vector / the rows of a full 3x3 published in GLOBAL order, read in local order
template <class PL> QEKF_FN void vec_ld(const PL &P, int w, typename PL::real v[3])
{
    constexpr int S = PL::stride;
    v[0] = sm_ld(P.pc + w * S); v[1] = sm_ld(P.pa + w * S); v[2] = sm_ld(P.pb + w * S);
}
template <class PL> QEKF_FN void mat_ld(const PL &P, int w, typename PL::real m[9])   // published at fullslot(a,b)
{
    constexpr int S = PL::stride;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int d = ((a - k + 3) % 3) * 3;
            m[a * 3 + k] = sm_ld((k == 0 ? P.pc : (k == 1 ? P.pa : P.pb)) + (w + d) * S);
        }
}

// ------------------------------------------------------------------------------------------------
// global packed covariance [NP][ld]  <->  (shared words, private entries)
// ------------------------------------------------------------------------------------------------
// global packed index of shared word w of block (X,Y)
template <int NB> QEKF_FN constexpr int packed_of_word(int X, int Y, int w)
{
    constexpr int N = 3 * NB;
    int a = 0, b = 0;
    if (X == Y) { a = (w == 0) ? 1 : 0; b = (w == 2) ? 1 : 2; }          // slot = the index that is missing
    else { b = w % 3; a = (b + w / 3 + 1) % 3; }                          // slot = ((a-b) mod 3 - 1)*3 + b
    return sym_idx<N>(3 * X + a, 3 * Y + b);
}
// Each lane moves the shared words w with w % 3 == c of every block, and its own private entries.
// `gP` points at this filter's element 0 of a [NP][ld] array.
template <class PL, typename TG> QEKF_FN void cov_load(PL &P, const TG *gP, int64_t ld)
{
    constexpr int NB = PL::nb, S = PL::stride, N = 3 * NB;
#pragma unroll
    for (int X = 0; X < NB; ++X)
#pragma unroll
        for (int Y = X; Y < NB; ++Y) {
            P.priv[bidx<NB>(X, Y)] = (typename PL::real)gP[(int64_t)sym_idx<N>(3 * X + P.c, 3 * Y + P.c) * ld];
#pragma unroll
            for (int w = 0; w < (X == Y ? 3 : 6); ++w)
                if ((w % 3) == P.c)
                    sm_st(P.sh + (sbase<NB>(X, Y) + w) * S, (typename PL::real)gP[(int64_t)packed_of_word<NB>(X, Y, w) * ld]);
        }
}
template <class PL, typename TG> QEKF_FN void cov_store(const PL &P, TG *gP, int64_t ld)
{
    constexpr int NB = PL::nb, S = PL::stride, N = 3 * NB;
#pragma unroll
    for (int X = 0; X < NB; ++X)
#pragma unroll
        for (int Y = X; Y < NB; ++Y) {
            gP[(int64_t)sym_idx<N>(3 * X + P.c, 3 * Y + P.c) * ld] = (TG)P.priv[bidx<NB>(X, Y)];
#pragma unroll
            for (int w = 0; w < (X == Y ? 3 : 6); ++w)
                if ((w % 3) == P.c)
                    gP[(int64_t)packed_of_word<NB>(X, Y, w) * ld] = (TG)sm_ld(P.sh + (sbase<NB>(X, Y) + w) * S);
        }
}
// cov_pert = diag(cov_init)  (initialize_state, cpp:340-343)
template <class PL, typename T> QEKF_FN void cov_init(PL &P, const Consts<T> &c)
{
    constexpr int NB = PL::nb, S = PL::stride;
#pragma unroll
    for (int X = 0; X < NB; ++X)
#pragma unroll
        for (int Y = X; Y < NB; ++Y) {
            P.priv[bidx<NB>(X, Y)] = (X == Y) ? c.cov_init[X] : T(0);
#pragma unroll
            for (int w = 0; w < (X == Y ? 3 : 6); ++w)
                if ((w % 3) == P.c) sm_st(P.sh + (sbase<NB>(X, Y) + w) * S, T(0));
        }
}

// ------------------------------------------------------------------------------------------------
// the parameters seen through the lane's relabelling
// ------------------------------------------------------------------------------------------------
// Relabelled copies of the launch-wide constants, one per lane role, [3][RC_N] words: in shared memory on the
// device (a role-indexed read of the kernel parameters would be a dynamically indexed parameter access).
enum { RC_G = 0, RC_ABS = 3, RC_WBS = 6, RC_Q = 9, RC_RA = 21, RC_RC = 24, RC_RAS = 30, RC_D = 36, RC_CVC = 45, RC_RVCV = 54,
       RC_QVC = 57, RC_N = 61 };

template <typename T, class PAR> struct RotPar {
    const PAR &p;
    const Consts<T> &c;
    int i0, i1, i2;
    SPtr<T> gs9;          // g, ab_static, wb_static already relabelled (the head of this role's constant copy)
    QEKF_FN int ix(int a) const { return a == 0 ? i0 : (a == 1 ? i1 : i2); }
    QEKF_FN int sym3(int e) const      // packed symmetric 3x3 index (00 01 02 11 12 22) of the relabelled entry
    {
        const int a = (e < 3) ? 0 : (e < 5 ? 1 : 2), b = (e < 3) ? e : (e < 5 ? e - 2 : 2);
        const int ga = ix(a), gb = ix(b);
        const int lo = ga < gb ? ga : gb, hi = ga < gb ? gb : ga;
        return lo * 3 - (lo * (lo - 1)) / 2 + (hi - lo);
    }
    QEKF_FN T Q(int i) const { return p.Q(3 * (i / 3) + ix(i % 3)); }
    QEKF_FN T Ra(int i) const { return p.Ra(ix(i)); }
    QEKF_FN T RC(int e) const { return p.RC(sym3(e)); }
    QEKF_FN T RA(int e) const { return p.RA(sym3(e)); }
    QEKF_FN T D(int i) const { return p.D(3 * ix(i / 3) + ix(i % 3)); }
    QEKF_FN T C_vc(int i) const { return p.C_vc(3 * ix(i / 3) + ix(i % 3)); }
    QEKF_FN T r_v_cv(int i) const { return p.r_v_cv(ix(i)); }
    QEKF_FN T q_vc(int i) const { return p.q_vc(i < 3 ? ix(i) : 3); }
    QEKF_FN double meas_delay() const { return p.meas_delay(); }
    QEKF_FN double dyn_offset() const { return p.dyn_offset(); }
    QEKF_FN T g(int a) const { return sm_ld(gs9 + (RC_G + a)); }
    QEKF_FN T ab_static(int a) const { return sm_ld(gs9 + (RC_ABS + a)); }
    QEKF_FN T wb_static(int a) const { return sm_ld(gs9 + (RC_WBS + a)); }
    QEKF_FN T dT() const { return c.dT; }
    QEKF_FN T small_ang_tol() const { return c.small_ang_tol; }
    QEKF_FN T cov_init(int i) const { return c.cov_init[i]; }
};
// the launch-wide parameters entirely from this role's relabelled copy
template <typename T> struct ParS {
    const Consts<T> &c;
    SPtr<T> rc;
    QEKF_FN T Q(int i) const { return sm_ld(rc + (RC_Q + i)); }
    QEKF_FN T Ra(int i) const { return sm_ld(rc + (RC_RA + i)); }
    QEKF_FN T RC(int e) const { return sm_ld(rc + (RC_RC + e)); }
    QEKF_FN T RA(int e) const { return sm_ld(rc + (RC_RAS + e)); }
    QEKF_FN T D(int i) const { return sm_ld(rc + (RC_D + i)); }
    QEKF_FN T C_vc(int i) const { return sm_ld(rc + (RC_CVC + i)); }
    QEKF_FN T r_v_cv(int i) const { return sm_ld(rc + (RC_RVCV + i)); }
    QEKF_FN T q_vc(int i) const { return sm_ld(rc + (RC_QVC + i)); }
    QEKF_FN double meas_delay() const { return c.meas_delay; }
    QEKF_FN double dyn_offset() const { return c.dyn_offset; }
    QEKF_FN T g(int a) const { return sm_ld(rc + (RC_G + a)); }
    QEKF_FN T ab_static(int a) const { return sm_ld(rc + (RC_ABS + a)); }
    QEKF_FN T wb_static(int a) const { return sm_ld(rc + (RC_WBS + a)); }
    QEKF_FN T dT() const { return c.dT; }
    QEKF_FN T small_ang_tol() const { return c.small_ang_tol; }
    QEKF_FN T cov_init(int i) const { return c.cov_init[i]; }
};
// fill one role's copy rc[RC_N] (plain memory) from the launch-wide constants
template <typename T> QEKF_FN void fill_role_consts(const Consts<T> &c, int role, T *rc)
{
    const int ix[3] = { role, (role + 1) % 3, (role + 2) % 3 };
    auto sym3 = [&](int a, int b) {
        const int ga = ix[a], gb = ix[b], lo = ga < gb ? ga : gb, hi = ga < gb ? gb : ga;
        return lo * 3 - (lo * (lo - 1)) / 2 + (hi - lo);
    };
    for (int a = 0; a < 3; ++a) {
        rc[RC_G + a] = c.g[ix[a]]; rc[RC_ABS + a] = c.ab_static[ix[a]]; rc[RC_WBS + a] = c.wb_static[ix[a]];
        rc[RC_RA + a] = c.Ra[ix[a]]; rc[RC_RVCV + a] = c.r_v_cv[ix[a]]; rc[RC_QVC + a] = c.q_vc[ix[a]];
        for (int b = 0; b < 4; ++b) rc[RC_Q + 3 * b + a] = c.Q[3 * b + ix[a]];
        for (int b = 0; b < 3; ++b) { rc[RC_D + 3 * a + b] = c.D[3 * ix[a] + ix[b]]; rc[RC_CVC + 3 * a + b] = c.C_vc[3 * ix[a] + ix[b]]; }
    }
    rc[RC_QVC + 3] = c.q_vc[3];
    {
        int e = 0;
        for (int a = 0; a < 3; ++a)
            for (int b = a; b < 3; ++b) { rc[RC_RC + e] = c.RC[sym3(a, b)]; rc[RC_RAS + e] = c.RA[sym3(a, b)]; ++e; }
    }
}

// relabel a global 3-vector into the lane's local indices without indexing registers dynamically
template <typename T> QEKF_FN void rot3(const T g[3], int c, T l[3])
{
    l[0] = c == 0 ? g[0] : (c == 1 ? g[1] : g[2]);
    l[1] = c == 0 ? g[1] : (c == 1 ? g[2] : g[0]);
    l[2] = c == 0 ? g[2] : (c == 1 ? g[0] : g[1]);
}

// what a tick carries from one phase to the next
template <typename T> struct TickCarry {
    T B[9], a[3];     // B = -dT C (row-major), a = specific force, in local indices; A = B skew(a) is applied as B (a x .)
    T accR[3], accT[3];
    T Mv0, PD0;
};

template <typename T> QEKF_FN void cross_add(const T a[3], const T x[3], const T y[3], T t[3])   // t = a x x + y
{
    t[0] = M<T>::fma_(a[1], x[2], M<T>::fma_(-a[2], x[1], y[0]));
    t[1] = M<T>::fma_(a[2], x[0], M<T>::fma_(-a[0], x[2], y[1]));
    t[2] = M<T>::fma_(a[0], x[1], M<T>::fma_(-a[1], x[0], y[2]));
}
template <typename T> QEKF_FN void cross_set(const T a[3], const T x[3], T t[3])                  // t = a x x
{
    t[0] = M<T>::fma_(a[1], x[2], -a[2] * x[1]);
    t[1] = M<T>::fma_(a[2], x[0], -a[0] * x[2]);
    t[2] = M<T>::fma_(a[0], x[1], -a[1] * x[0]);
}
template <typename T> QEKF_FN void mv_acc(const T A[9], const T t[3], T o[3])                     // o += A t
{
#pragma unroll
    for (int a = 0; a < 3; ++a) o[a] = M<T>::fma_(A[a * 3 + 2], t[2], M<T>::fma_(A[a * 3 + 1], t[1], M<T>::fma_(A[a * 3], t[0], o[a])));
}

// ------------------------------------------------------------------------------------------------
// prediction_step, phase 0 (role 0 only): nominal kinematics, and the Jacobian pieces published for all three
// lanes at word offset `jw` (this tick's half of the exchange)               relative_pose_EKF.cpp:346-401
// ------------------------------------------------------------------------------------------------
template <typename T, class PL>
QEKF_FN void kin_step(Nominal<T> &s, T accel[3], const T u[6], const Consts<T> &c, PL &P, int jw)
{
    constexpr int S = PL::stride;
    using L = typename PL::L;
    const T d = c.dT;
    T a[3], w[3], C[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        a[i] = u[i] - s.ab[i] - c.ab_static[i];
        w[i] = u[3 + i] - s.wb[i] - c.wb_static[i];
    }
    quat_to_rot(s.q, C);
    T acc[3];
    mv(C, a, acc);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        acc[i] += c.g[i];
        accel[i] = acc[i];
        s.r[i] = M<T>::fma_(d, s.v[i], s.r[i]);     // uses the old v (explicit Euler)
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) s.v[i] = M<T>::fma_(d, acc[i], s.v[i]);
    const SPtr<T> j = P.sh + jw * S;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) sm_st(j + (L::JB_B + fullslot(r, k)) * S, -d * C[r * 3 + k]);
    T dth[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        dth[i] = d * w[i];
        sm_st(j + (L::JB_A + i) * S, a[i]);
        sm_st(j + (L::JB_DTH + i) * S, dth[i]);
    }
    PhiCoef<T> pc;
    attitude_step(s.q, dth, c.small_ang_tol, pc);
    sm_st(j + (L::JB_PC + 0) * S, pc.cs);
    sm_st(j + (L::JB_PC + 1) * S, pc.s1);
    sm_st(j + (L::JB_PC + 2) * S, pc.s2);
}

// this lane's base pointers shifted by a run-time word offset (one half of a double-buffered exchange)
template <typename T> struct Tri { SPtr<T> c, a, b, h; };
template <class PL> QEKF_FN Tri<typename PL::real> tri(const PL &P, int w)
{
    constexpr int S = PL::stride;
    return Tri<typename PL::real>{ P.pc + w * S, P.pa + w * S, P.pb + w * S, P.sh + w * S };
}
// Phi (row-major, local indices) from the published dT w and coefficients
template <class PL, typename T = typename PL::real> QEKF_FN void phi_ld(const PL &P, int jw, T Phi[9])
{
    constexpr int S = PL::stride;
    using L = typename PL::L;
    const Tri<T> j = tri(P, jw);
    const T dth[3] = { sm_ld(j.c + L::JB_DTH * S), sm_ld(j.a + L::JB_DTH * S), sm_ld(j.b + L::JB_DTH * S) };
    const PhiCoef<T> pc{ sm_ld(j.h + (L::JB_PC + 0) * S), sm_ld(j.h + (L::JB_PC + 1) * S), sm_ld(j.h + (L::JB_PC + 2) * S) };
    phi_matrix(pc, dth, Phi);
}

// Phase 1: E1 on the columns (r,Y) += dT (v,Y); the parts of the (r,r) and (v,v) updates that need OLD values.
template <bool BIAS, class PL, class RP, typename T = typename PL::real>
QEKF_FN void pred_stage1(PL &P, const RP &rp, TickCarry<T> &k, int jw)
{
    constexpr int S = PL::stride;
    using L = typename PL::L;
    const T d = rp.c.dT;
    {   // B and a of this tick, relabelled: entry (r,k) of B sits at fullslot(r,k) = ((r-k) mod 3)*3 + k
        const Tri<T> j = tri(P, jw);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            k.B[r * 3 + 0] = sm_ld(j.c + (L::JB_B + ((r + 3) % 3) * 3) * S);
            k.B[r * 3 + 1] = sm_ld(j.a + (L::JB_B + ((r + 2) % 3) * 3) * S);
            k.B[r * 3 + 2] = sm_ld(j.b + (L::JB_B + ((r + 1) % 3) * 3) * S);
        }
        k.a[0] = sm_ld(j.c + L::JB_A * S); k.a[1] = sm_ld(j.a + L::JB_A * S); k.a[2] = sm_ld(j.b + L::JB_A * S);
    }
    T vv[3], x[3], y[3], t[3];
    col_ld<BV, BV>(P, vv);
    col_ld<BR, BV>(P, x);
#pragma unroll
    for (int i = 0; i < 3; ++i) x[i] = M<T>::fma_(d, vv[i], x[i]);
    col_st<BR, BV>(P, x);
    // (r,r)' = rr + dT (N + N^T) - dT^2 vv with N = the new (r,v); column c of N is x, row c comes after the barrier
    col_ld<BR, BR>(P, y);
#pragma unroll
    for (int i = 0; i < 3; ++i) k.accR[i] = M<T>::fma_(d, x[i], M<T>::fma_(-d * d, vv[i], y[i]));
    col_ld<BV, BTH>(P, y);
    col_ld<BR, BTH>(P, x);
#pragma unroll
    for (int i = 0; i < 3; ++i) x[i] = M<T>::fma_(d, y[i], x[i]);
    col_st<BR, BTH>(P, x);
    if constexpr (BIAS) {
        col_ld<BV, BAB>(P, y);
        col_ld<BR, BAB>(P, x);
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = M<T>::fma_(d, y[i], x[i]);
        col_st<BR, BAB>(P, x);
        col_ld<BV, BWB>(P, y);
        col_ld<BR, BWB>(P, x);
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = M<T>::fma_(d, y[i], x[i]);
        col_st<BR, BWB>(P, x);
    }
    // column c of M = vv + A (th,v) + B (ab,v), from the old row c of (v,th), (v,ab)
    col_ld<BTH, BV>(P, x);
    if constexpr (BIAS) {
        col_ld<BAB, BV>(P, y);
        cross_add(k.a, x, y, t);
    } else {
        cross_set(k.a, x, t);
    }
    mv_acc(k.B, t, vv);
    k.Mv0 = vv[0];
    sm_st(P.pc + L::XV * S, vv[1]);
    sm_st(P.pc + (L::XV + 3) * S, vv[2]);
}

// Phase 2: finish (r,r); E2 on the columns (v,Y) += A (th,Y) + B (ab,Y), Y != v.
template <bool BIAS, class PL, class RP, typename T = typename PL::real>
QEKF_FN void pred_stage2(PL &P, const RP &rp, TickCarry<T> &k)
{
    const T d = rp.c.dT;
    T n[3], x[3], y[3], t[3];
    col_ld<BV, BR>(P, n);                 // row c of the new (r,v)
#pragma unroll
    for (int i = 0; i < 3; ++i) x[i] = M<T>::fma_(d, n[i], k.accR[i]);
    col_st<BR, BR>(P, x);
    col_ld<BTH, BR>(P, x);
    if constexpr (BIAS) { col_ld<BAB, BR>(P, y); cross_add(k.a, x, y, t); } else { cross_set(k.a, x, t); }
    mv_acc(k.B, t, n);
    col_st<BV, BR>(P, n);
    col_ld<BV, BTH>(P, n);
    col_ld<BTH, BTH>(P, x);
    if constexpr (BIAS) { col_ld<BAB, BTH>(P, y); cross_add(k.a, x, y, t); } else { cross_set(k.a, x, t); }
    mv_acc(k.B, t, n);
    col_st<BV, BTH>(P, n);
    if constexpr (BIAS) {
        col_ld<BV, BAB>(P, n);
        col_ld<BTH, BAB>(P, x);
        col_ld<BAB, BAB>(P, y);
        cross_add(k.a, x, y, t);
        mv_acc(k.B, t, n);
        col_st<BV, BAB>(P, n);
        col_ld<BV, BWB>(P, n);
        col_ld<BTH, BWB>(P, x);
        col_ld<BAB, BWB>(P, y);
        cross_add(k.a, x, y, t);
        mv_acc(k.B, t, n);
        col_st<BV, BWB>(P, n);
    }
}

// Phase 3: finish (v,v) (row c, + C Qa C^T); E3 on the columns (th,Y) <- Phi (th,Y) - dT (wb,Y), Y != th.
template <bool BIAS, class PL, class RP, typename T = typename PL::real>
QEKF_FN void pred_stage3(PL &P, const RP &rp, TickCarry<T> &k, int jw)
{
    constexpr int S = PL::stride;
    using L = typename PL::L;
    const T d = rp.c.dT;
    T nt[3], y[3], t[3], vr[3], x[3];
    col_ld<BTH, BV>(P, nt);               // row c of the new (v,th)
    if constexpr (BIAS) { col_ld<BAB, BV>(P, y); cross_add(k.a, nt, y, t); } else { cross_set(k.a, nt, t); }
    vr[0] = k.Mv0;
    vr[1] = sm_ld(P.pa + (L::XV + 3) * S);
    vr[2] = sm_ld(P.pb + L::XV * S);
    mv_acc(k.B, t, vr);
    {   // + row c of C diag(Qa) C^T = (B diag(Qa) B^T) / dT^2
        const T inv = T(1) / (d * d);
        const T q0 = k.B[0] * (rp.Q(0) * inv), q1 = k.B[1] * (rp.Q(1) * inv), q2 = k.B[2] * (rp.Q(2) * inv);
#pragma unroll
        for (int a = 0; a < 3; ++a)
            vr[a] = M<T>::fma_(q2, k.B[a * 3 + 2], M<T>::fma_(q1, k.B[a * 3 + 1], M<T>::fma_(q0, k.B[a * 3], vr[a])));
    }
    row_st<BV>(P, vr);
    T Phi[9];
    phi_ld(P, jw, Phi);
    // (th,r)
    col_ld<BTH, BR>(P, x);
    if constexpr (BIAS) {
        col_ld<BWB, BR>(P, y);
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = -d * y[i];
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = T(0);
    }
    mv_acc(Phi, x, t);
    col_st<BTH, BR>(P, t);
    // (th,v)
    if constexpr (BIAS) {
        col_ld<BWB, BV>(P, y);
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = -d * y[i];
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = T(0);
    }
    mv_acc(Phi, nt, t);
    col_st<BTH, BV>(P, t);
    if constexpr (BIAS) {
        col_ld<BTH, BAB>(P, x);
        col_ld<BWB, BAB>(P, y);
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = -d * y[i];
        mv_acc(Phi, x, t);
        col_st<BTH, BAB>(P, t);
        col_ld<BTH, BWB>(P, x);
        col_ld<BWB, BWB>(P, y);
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = -d * y[i];
        mv_acc(Phi, x, t);
        col_st<BTH, BWB>(P, t);
        // (th,th)' = Phi D Phi^T - dT (N + N^T) - dT^2 (wb,wb), N = the new (th,wb): the column-c part
#pragma unroll
        for (int i = 0; i < 3; ++i) k.accT[i] = -d * M<T>::fma_(d, y[i], t[i]);
    }
    // column c of Phi D; its transposed view comes through the exchange
    col_ld<BTH, BTH>(P, x);
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = T(0);
    mv_acc(Phi, x, t);
    k.PD0 = t[0];
    sm_st(P.pc + L::XT * S, t[1]);
    sm_st(P.pc + (L::XT + 3) * S, t[2]);
}

// Phase 4: finish (th,th) (row c), process noise on the diagonals.
template <bool BIAS, class PL, class RP, typename T = typename PL::real>
QEKF_FN void pred_stage4(PL &P, const RP &rp, TickCarry<T> &k, int jw)
{
    constexpr int S = PL::stride, NB = PL::nb;
    using L = typename PL::L;
    const T d = rp.c.dT;
    T pd[3] = { k.PD0, sm_ld(P.pa + (L::XT + 3) * S), sm_ld(P.pb + L::XT * S) };
    T Phi[9], tr[3];
    phi_ld(P, jw, Phi);
    if constexpr (BIAS) {
        T nw[3];
        col_ld<BWB, BTH>(P, nw);          // row c of the new (th,wb)
#pragma unroll
        for (int i = 0; i < 3; ++i) tr[i] = M<T>::fma_(-d, nw[i], k.accT[i]);
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) tr[i] = T(0);
    }
    mv_acc(Phi, pd, tr);
    tr[0] += rp.Q(3);
    row_st<BTH>(P, tr);
    if constexpr (BIAS) {
        P.priv[bidx<NB>(BAB, BAB)] += rp.Q(6);
        P.priv[bidx<NB>(BWB, BWB)] += rp.Q(9);
    }
}

// ------------------------------------------------------------------------------------------------
// correction_step                                               relative_pose_EKF.cpp:417-502
//
// With B = P G^T and k_j = S^-1 B[j,:]^T (row j of the gain K):  dx_j = k_j . dy,
//   P'[i,j] = P[i,j] - B[i,:] . k_j                     for i, j outside the dr / dth block columns,
//   P' G^T  = K R_k  (= B - K (S - R_k))                for the dr / dth block columns themselves,
// so lane c needs, besides its own columns, the full (X,r), (X,th) blocks (X = v, ab, wb) and (r,r), (r,th),
// (th,th).  Phase A publishes the private entries those blocks hold; phase B does all the arithmetic and
// updates the blocks among {v, ab, wb} (which read only unmodified dr / dth columns); phase C stores the new
// dr / dth columns and injects the error state.
// ------------------------------------------------------------------------------------------------
template <typename T, int NB> struct CorrCarry {
    T nr[NB][3], nt[NB][3];    // new columns (r,Y), (th,Y)
    T dx[NB];                  // component c of the injected error of every block
};

// `s`: the nominal state (role 0) or nullptr (roles 1, 2)
template <class PL, typename T = typename PL::real>
QEKF_FN void corr_publish(PL &P, const Nominal<T> *s)
{
    constexpr int S = PL::stride, NB = PL::nb;
    using L = typename PL::L;
#pragma unroll
    for (int X = 0; X < NB; ++X)
#pragma unroll
        for (int Y = X; Y < NB; ++Y)
            if (pubidx<NB>(X, Y) >= 0) sm_st(P.pc + (L::PB + 3 * pubidx<NB>(X, Y)) * S, P.priv[bidx<NB>(X, Y)]);
    if (s) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sm_st(P.sh + (L::QR + i) * S, s->q[i]);
#pragma unroll
        for (int i = 0; i < 3; ++i) sm_st(P.sh + (L::QR + 4 + i) * S, s->r[i]);
    }
}

// gain row, injected error and the new dr / dth entries of this lane's column of block Y
template <int Y, bool DIRECT, class PL, typename T>
QEKF_FN void corr_column(PL &P, const T Sinv[21], const T Rk[21], const T Gam[9], const T dy[6], T kY[6],
                         CorrCarry<T, PL::nb> &cc)
{
    constexpr int S = PL::stride;
    using L = typename PL::L;
    T x[3], t[3], b[6];
    col_ld<BR, Y>(P, x);
    col_ld<BTH, Y>(P, t);
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        T v = x[m];
        if (!DIRECT) v = M<T>::fma_(Gam[m * 3 + 2], t[2], M<T>::fma_(Gam[m * 3 + 1], t[1], M<T>::fma_(Gam[m * 3], t[0], v)));
        b[m] = v;
        b[3 + m] = t[m];
    }
    T dx = T(0);
#pragma unroll
    for (int m = 0; m < 6; ++m) {
        T v = T(0);
#pragma unroll
        for (int l = 0; l < 6; ++l) v = M<T>::fma_(b[l], Sinv[sym_idx<6>(l, m)], v);
        kY[m] = v;
        dx = M<T>::fma_(v, dy[m], dx);
    }
    cc.dx[Y] = dx;
    sm_st(P.pc + (L::DX + 3 * Y) * S, dx);
    T w[6];
#pragma unroll
    for (int m = 0; m < 6; ++m) {
        T v = T(0);
#pragma unroll
        for (int l = 0; l < 6; ++l) v = M<T>::fma_(kY[l], Rk[sym_idx<6>(l, m)], v);
        w[m] = v;
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        T v = w[m];
        if (!DIRECT) v = M<T>::fma_(-Gam[m * 3 + 2], w[5], M<T>::fma_(-Gam[m * 3 + 1], w[4], M<T>::fma_(-Gam[m * 3], w[3], v)));
        cc.nr[Y][m] = v;
        cc.nt[Y][m] = w[3 + m];
    }
}

// B_X = [ P(X,r) + P(X,th) Gam^T , P(X,th) ]  (3x6, local indices), X in {v, ab, wb}
template <int X, bool DIRECT, class PL, typename T> QEKF_FN void corr_bx(const PL &P, const T Gam[9], T Bx[18])
{
    T xr[9], xt[9];
    full_ld<BR, X>(P, xr);                                   // stored (r,X): transposed use
    if constexpr (X == BV) full_ld<BV, BTH>(P, xt); else full_ld<BTH, X>(P, xt);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            const T th_am = (X == BV) ? xt[a * 3 + m] : xt[m * 3 + a];
            Bx[a * 6 + 3 + m] = th_am;
        }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            T v = xr[m * 3 + a];
            if (!DIRECT) {
#pragma unroll
                for (int n = 0; n < 3; ++n) v = M<T>::fma_(Bx[a * 6 + 3 + n], Gam[m * 3 + n], v);
            }
            Bx[a * 6 + m] = v;
        }
}
// column c of block (X,Y) -= B_X k_Y.  Of a diagonal block a lane loads only the entries it stores (the others are
// being rewritten by the lanes that own them).
template <int X, int Y, class PL, typename T> QEKF_FN void corr_down(PL &P, const T Bx[18], const T kY[6])
{
    T v[3];
    if constexpr (X == Y) {
        constexpr int b = sbase<PL::nb>(X, X) * PL::stride;
        v[0] = P.priv[bidx<PL::nb>(X, X)];
        v[1] = P.wc1 ? sm_ld(P.pb + b) : T(0);
        v[2] = P.wc2 ? sm_ld(P.pa + b) : T(0);
    } else {
        col_ld<X, Y>(P, v);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        T s = v[a];
#pragma unroll
        for (int m = 0; m < 6; ++m) s = M<T>::fma_(-Bx[a * 6 + m], kY[m], s);
        v[a] = s;
    }
    col_st<X, Y>(P, v);
}

// Phase B.  `tag` is the tag pose in local indices, `obs` comes back in local indices.
template <bool BIAS, bool DIRECT, class PL, class RP, typename T = typename PL::real>
QEKF_FN void corr_stage1(PL &P, const T tag[7], const RP &rp, Observation<T> &obs, CorrCarry<T, PL::nb> &cc)
{
    constexpr int S = PL::stride;
    using L = typename PL::L;
    T dy[6], Rk[21], Gam[9], Sinv[21];
    {   // the published nominal attitude and position, relabelled
        T q[4], r[3];
        vec_ld(P, L::QR, q);
        q[3] = sm_ld(P.sh + (L::QR + 3) * S);
        vec_ld(P, L::QR + 4, r);
        correction_front<T, DIRECT>(q, r, tag, rp, obs, dy, Rk, Gam);
    }
    {   // S = G P G^T + R_k, inverted through Cholesky (as the thread-per-filter path)
        T rr[9], rt[9], tt[9], Brt[36], Sm[21];
        full_ld<BR, BR>(P, rr);
        full_ld<BR, BTH>(P, rt);
        full_ld<BTH, BTH>(P, tt);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                Brt[a * 6 + b] = rr[a * 3 + b];
                Brt[a * 6 + 3 + b] = rt[a * 3 + b];
                Brt[(3 + a) * 6 + b] = rt[b * 3 + a];
                Brt[(3 + a) * 6 + 3 + b] = tt[a * 3 + b];
            }
        if (!DIRECT) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    T x0 = Brt[a * 6 + b], x1 = Brt[(3 + a) * 6 + b];
#pragma unroll
                    for (int kk = 0; kk < 3; ++kk) {
                        x0 = M<T>::fma_(rt[a * 3 + kk], Gam[b * 3 + kk], x0);
                        x1 = M<T>::fma_(tt[a * 3 + kk], Gam[b * 3 + kk], x1);
                    }
                    Brt[a * 6 + b] = x0;
                    Brt[(3 + a) * 6 + b] = x1;
                }
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = i; j < 6; ++j) {
                    T x = Brt[i * 6 + j];
                    if (i < 3) {
#pragma unroll
                        for (int kk = 0; kk < 3; ++kk) x = M<T>::fma_(Gam[i * 3 + kk], Brt[(3 + kk) * 6 + j], x);
                    }
                    Sm[sym_idx<6>(i, j)] = x + Rk[sym_idx<6>(i, j)];
                }
        } else {
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = i; j < 6; ++j) Sm[sym_idx<6>(i, j)] = Brt[i * 6 + j] + Rk[sym_idx<6>(i, j)];
        }
        sym6_inverse(Sm, Sinv);
    }
    T kr[6], kv[6], kt[6];
    corr_column<BR, DIRECT>(P, Sinv, Rk, Gam, dy, kr, cc);
    corr_column<BTH, DIRECT>(P, Sinv, Rk, Gam, dy, kt, cc);
    corr_column<BV, DIRECT>(P, Sinv, Rk, Gam, dy, kv, cc);
    T Bx[18];
    if constexpr (BIAS) {
        T ka[6], kw[6];
        corr_column<BAB, DIRECT>(P, Sinv, Rk, Gam, dy, ka, cc);
        corr_column<BWB, DIRECT>(P, Sinv, Rk, Gam, dy, kw, cc);
        corr_bx<BV, DIRECT>(P, Gam, Bx);
        corr_down<BV, BV>(P, Bx, kv);
        corr_down<BV, BAB>(P, Bx, ka);
        corr_down<BV, BWB>(P, Bx, kw);
        corr_bx<BAB, DIRECT>(P, Gam, Bx);
        corr_down<BAB, BAB>(P, Bx, ka);
        corr_down<BAB, BWB>(P, Bx, kw);
        corr_bx<BWB, DIRECT>(P, Gam, Bx);
        corr_down<BWB, BWB>(P, Bx, kw);
    } else {
        corr_bx<BV, DIRECT>(P, Gam, Bx);
        corr_down<BV, BV>(P, Bx, kv);
    }
}

// Phase C: the new dr / dth columns.
template <bool BIAS, class PL, typename T = typename PL::real>
QEKF_FN void corr_stage2(PL &P, const CorrCarry<T, PL::nb> &cc)
{
    col_st<BR, BR>(P, cc.nr[BR]);
    col_st<BR, BV>(P, cc.nr[BV]);
    col_st<BTH, BV>(P, cc.nt[BV]);
    col_st<BR, BTH>(P, cc.nr[BTH]);
    col_st<BTH, BTH>(P, cc.nt[BTH]);
    if constexpr (BIAS) {
        col_st<BR, BAB>(P, cc.nr[BAB]);
        col_st<BTH, BAB>(P, cc.nt[BAB]);
        col_st<BR, BWB>(P, cc.nr[BWB]);
        col_st<BTH, BWB>(P, cc.nt[BWB]);
    }
}
// ... and (role 0) the injection of the error state into the nominal one (cpp:484-498)
template <bool BIAS, class PL, typename T = typename PL::real>
QEKF_FN void corr_inject(const PL &P, Nominal<T> &s)
{
    constexpr int S = PL::stride;
    using L = typename PL::L;
    T dth[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        s.r[i] += sm_ld(P.sh + (L::DX + 3 * BR + i) * S);
        s.v[i] += sm_ld(P.sh + (L::DX + 3 * BV + i) * S);
        dth[i] = sm_ld(P.sh + (L::DX + 3 * BTH + i) * S);
    }
    {
        T qe[4], qn[4], nn, sh, ch;
        quat_exp(dth, qe, nn, sh, ch);
        quat_mul(s.q, qe, qn);
        quat_normclip(qn);
#pragma unroll
        for (int i = 0; i < 4; ++i) s.q[i] = qn[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if constexpr (BIAS) {
            s.ab[i] += sm_ld(P.sh + (L::DX + 3 * BAB + i) * S);
            s.wb[i] += sm_ld(P.sh + (L::DX + 3 * BWB + i) * S);
        } else {
            s.ab[i] = T(0); s.wb[i] = T(0);
        }
    }
}


// ------------------------------------------------------------------------------------------------
// the replay loop of one lane                   filter_update, relative_pose_EKF.cpp:127-303 (single-rate)
// ------------------------------------------------------------------------------------------------
// barrier of the 96 threads (3 warps) that share 32 filters; the host harness plugs in a thread barrier
struct GroupSync {
    int id;                  // device: named barrier 1..15
    void (*fn)(void *);      // host harness: thread barrier
    void *ctx;
    int nthreads = 96;       // device: threads that meet at the barrier
#ifdef __CUDA_ARCH__
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
    __device__ __forceinline__ void cta_sync() const { __syncthreads(); }
    // CTA-wide "does anybody still ...": one barrier per loop iteration, which also keeps the groups of a CTA within
    // one iteration of each other so that they share the instruction cache lines of the (large, unrolled) tick
    __device__ __forceinline__ bool cta_any(bool x) const { return __syncthreads_or(x) != 0; }
#else
    void sync() const { fn(ctx); }
    void cta_sync() const { fn(ctx); }
    bool cta_any(bool x) const { return x; }
#endif
};

// where the group sits in its CTA, and the CTA-wide scratch the statistics sample uses (it aliases the
// correction-only words, which no group uses while all of them are at a sampling point)
template <typename T> struct CtaCtx {
    int group, n_groups;
    T *scratch;              // [NP + 6][32]
};

struct GroupVote {
    int active, want, fenced, at_fence, lanes;
    bool out_of_patience;
};
// The three warps of a group hold identical copies of the sequencing state, so a vote over the group's 32 filters
// is a vote over the lanes of one warp: no shared memory, no barrier.
QEKF_FN GroupVote group_vote(bool active, bool want, bool oop, bool fenced, bool at_fence)
{
    GroupVote r;
#ifdef __CUDA_ARCH__
    const unsigned full = 0xffffffffu;
    r.active = __popc(__ballot_sync(full, active));
    r.want = __popc(__ballot_sync(full, want));
    r.fenced = __popc(__ballot_sync(full, fenced));
    r.at_fence = __popc(__ballot_sync(full, at_fence));
    r.out_of_patience = __ballot_sync(full, oop) != 0u;
    r.lanes = 32;
#else
    r.active = active; r.want = want; r.fenced = fenced; r.at_fence = at_fence; r.out_of_patience = oop; r.lanes = 1;
#endif
    return r;
}
QEKF_FN bool group_any(bool x)
{
#ifdef __CUDA_ARCH__
    return __ballot_sync(0xffffffffu, x) != 0u;
#else
    return x;
#endif
}

// covariance accessor that swallows stores (initialize_state is reused for the nominal part only)
template <typename T, int N> struct PNull {
    static constexpr int n = N;
    QEKF_FN void st(int, int, T) {}
};

// Roles 1 and 2 draw the noisy IMU sample of tick k -- the realisation synth_imu draws -- into half `uw` of the
// exchange: role 1 components 0..3 (two Box-Muller pairs), role 2 components 4, 5.
template <class PL, typename T = typename PL::real>
QEKF_FN void synth_imu_role(const NoiseSpec &ns, int64_t gid, int64_t k, const double *clean6, const double bias[4], PL &P, int uw)
{
    constexpr int S = PL::stride;
    uint32_t w[4];
    const uint32_t k0 = (uint32_t)ns.seed, k1 = (uint32_t)(ns.seed >> 32);
    const uint32_t g0 = (uint32_t)(uint64_t)gid, g1 = (uint32_t)((uint64_t)gid >> 32);
    const SPtr<T> u = P.sh + uw * S;
    uint32_t uu[6];
    philox4x32_10((uint32_t)k, STREAM_IMU, g0, g1, k0, k1, w);      // (one block; each role turns its share into normals)
    uniforms21x6(w, uu);
    if (P.c == 1) {
        float z0, z1, z2, z3;
        box_muller(uu[0], uu[1], z0, z1);
        box_muller(uu[2], uu[3], z2, z3);
        sm_st(u + 0 * S, (T)(clean6[0] + bias[0] + (double)ns.sig_a * (double)z0));
        sm_st(u + 1 * S, (T)(clean6[1] + bias[1] + (double)ns.sig_a * (double)z1));
        sm_st(u + 2 * S, (T)(clean6[2] + bias[2] + (double)ns.sig_a * (double)z2));
        sm_st(u + 3 * S, (T)(clean6[3] + bias[3] + (double)ns.sig_w * (double)z3));
    } else {
        float z4, z5;
        box_muller(uu[4], uu[5], z4, z5);
        sm_st(u + 4 * S, (T)(clean6[4] + bias[0] + (double)ns.sig_w * (double)z4));
        sm_st(u + 5 * S, (T)(clean6[5] + bias[1] + (double)ns.sig_w * (double)z5));
    }
}

// One statistics sample of the group's 32 filters: the three lanes assemble the packed covariance in the scratch,
// role 0 (which holds the nominal state) runs the thread-per-filter sampling code on that copy.
template <typename T, bool BIAS, class PL>
QEKF_COLD void coop_stats_sample(const RunArgs<T> &a, int64_t i, int64_t k_done, const PL P, const Nominal<T> s, T *scr,
                                 bool mine, const GroupSync gs)
{
    constexpr int NB = PL::nb, N = 3 * NB, S = PL::stride;
#ifdef __CUDA_ARCH__
    const SPtr<T> my = SPtr<T>::from(scr + (threadIdx.x & 31));
    constexpr int SS = 32;
#else
    const SPtr<T> my = SPtr<T>::from(scr);
    constexpr int SS = 1;
#endif
#pragma unroll
    for (int X = 0; X < NB; ++X)
#pragma unroll
        for (int Y = X; Y < NB; ++Y) {
            sm_st(my + sym_idx<N>(3 * X + P.c, 3 * Y + P.c) * SS, P.priv[bidx<NB>(X, Y)]);
#pragma unroll
            for (int w = 0; w < (X == Y ? 3 : 6); ++w)
                if ((w % 3) == P.c) sm_st(my + packed_of_word<NB>(X, Y, w) * SS, sm_ld(P.sh + (sbase<NB>(X, Y) + w) * S));
        }
    gs.sync();
    if (P.c == 0) {
        double bias[6];
        true_bias(a.ns, a.ns.gid0 + i, bias);
        stats_sample<T, BIAS, PShared<T, N, SS>, false>(a, i, k_done, s, PShared<T, N, SS>{ my.generic() }, bias, mine);
    }
}

// `live` = false marks the padding lanes of a ragged last CTA (they take part in votes and barriers only).
// `rp` = the parameters as this lane sees them (ParS, or RotPar over the per-filter table).
template <typename T, bool BIAS, bool DIRECT, bool SYNTH, class PL, class RP>
QEKF_FN void run_filter_coop(const RunArgs<T> &a, const int64_t i_in, PL &P, const RP &rp, const bool live,
                             const GroupSync gs, const CtaCtx<T> cta)
{
    constexpr int S = PL::stride, NB = PL::nb;
    using L = typename PL::L;
    const Consts<T> &c = a.c;
    const int64_t i = live ? i_in : 0;
    const int64_t k_end = a.k0 + a.n_steps;
    const bool do_stats = SYNTH && a.stats.acc != nullptr;
    const int32_t patience = c.limit_measurement_freq ? (c.upd_per_meas - 1) : 0;
    const bool lead = P.c == 0;            // role 0: holds the nominal state, does the kinematics, stores the scalars

    Nominal<T> s;                          // role 0 only
    T accel[3] = { T(0), T(0), T(0) };     // role 0 only
    double nb[4] = { 0, 0, 0, 0 };         // roles 1, 2: the true bias of the components they draw
    int32_t flags = 0, upds = 0;
    Inputs<T, SYNTH> in;
    int64_t k = k_end;
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) { s.r[cc] = T(0); s.v[cc] = T(0); s.q[cc] = T(0); s.ab[cc] = T(0); s.wb[cc] = T(0); }
    s.q[3] = T(1);
    if (live) {
        const int64_t ld = a.st.ld;
        if (lead) {
            const T *x = a.st.x + i;
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                s.r[cc] = x[(0 + cc) * ld]; s.v[cc] = x[(3 + cc) * ld]; s.q[cc] = x[(6 + cc) * ld];
                s.ab[cc] = x[(10 + cc) * ld]; s.wb[cc] = x[(13 + cc) * ld];
                accel[cc] = a.st.aux[cc * ld + i];
            }
            s.q[3] = x[9 * ld];
        }
        cov_load(P, a.st.P + i, ld);
        flags = a.st.flags[i];
        upds = a.st.upds[i];
        in.init(a, i);
        k = a.k0;
        if (SYNTH && !lead) {
            if (P.c == 1) { nb[0] = in.bias[0]; nb[1] = in.bias[1]; nb[2] = in.bias[2]; nb[3] = in.bias[3]; }
            else { nb[0] = in.bias[4]; nb[1] = in.bias[5]; }
            if (k < k_end) synth_imu_role(a.ns, in.gid, k, a.in.imu + k * 6, nb, P, L::UB + 6 * (int)(k & 1));
        }
    }
    gs.sync();

    uint32_t n_pred = 0, n_corr = 0, n_iter = 0, n_sexec = 0;
    int32_t m = a.m0;
    int32_t next_tag_step = (m < a.in.M) ? a.in.tag_step[m] : INT32_MAX;
    int32_t pend_m = -1;
    int32_t held = 0;
    bool at_fence = false;
    int32_t rdv_left = 0;                  // sampling points of this launch the group still has to attend
    if (do_stats) rdv_left = (int32_t)(k_end / a.stats.stride - a.k0 / a.stats.stride);

    for (;;) {
        const bool active = (k < k_end) && !at_fence;

        // ---- AprilTagSubCallback for the arrival scheduled at tick k (node.cpp:153-176) ----
        bool init_now = false;
        if (active && k == next_tag_step) {
            if (in.valid(a.in, a.ns, m, (int32_t)k)) {
                pend_m = m;
                flags |= FLAG_READY;
                if (!(flags & FLAG_INIT)) {
                    if (lead) {        // initialize_state (cpp:305-344): the nominal part, in role 0's (= the global) frame
                        T tg[7];
                        in.tag(a.in, a.ns, m, tg);
                        PNull<T, 3 * NB> pn;
                        initialize_state<T, BIAS>(s, pn, tg, rp, false);   // role 0's parameter view is not relabelled
                    }
                    cov_init(P, c);
                    flags |= FLAG_INIT;
                    init_now = true;
                }
            }
            ++m;
            next_tag_step = (m < a.in.M) ? a.in.tag_step[m] : INT32_MAX;
        }
        if (group_any(init_now)) gs.sync();          // the fresh covariance words are visible to the other lanes

        const bool want = active && (flags & FLAG_INIT) && (flags & FLAG_READY) &&
                          (!c.limit_measurement_freq || (upds + 1) >= c.upd_per_meas);
        const GroupVote v = group_vote(active, want, want && held >= patience, at_fence || k >= k_end, at_fence);
        if (v.active == 0 && v.at_fence == 0) break;
        ++n_iter;
        if (do_stats && v.fenced == v.lanes && v.at_fence != 0) {
            // every filter of the group is at the sampling point (or finished): meet the other groups, then
            // the groups sample one after the other in the shared scratch
            const bool mine = live && at_fence && (flags & FLAG_INIT);
            gs.cta_sync();
            for (int g = 0; g < cta.n_groups; ++g) {
                if (g == cta.group) {
                    coop_stats_sample<T, BIAS>(a, i, k - 1, P, s, cta.scratch, mine, gs);
                    if (mine && lead) ++n_sexec;
                }
                gs.cta_sync();
            }
            --rdv_left;
            at_fence = false;
        }
        bool serve = true;
        if (v.want != 0) serve = (2 * v.want > v.active) || v.out_of_patience;
        if (want && !serve) ++held;
        const bool adv = active && !(want && !serve);
        const bool exec = adv && (flags & FLAG_INIT);

        // ---- consume the measurement, corner-margin gate (cpp:150-186) ----
        bool perform = false;
        T tag[7];
        if (exec && want) {
            T tg[7];
            if (pend_m >= 0) {
                in.tag(a.in, a.ns, pend_m, tg);
            } else {
#pragma unroll
                for (int cc = 0; cc < 7; ++cc) tg[cc] = (T)a.st.pend[cc * a.st.ld + i];
            }
            flags &= ~FLAG_READY;
            perform = c.corner_margin_enbl ? corner_gate<T>(tg, c) : true;
            held = 0;
            rot3(tg, P.c, tag);
            rot3(tg + 3, P.c, tag + 3);
            tag[6] = tg[6];
        }

        if (v.active != 0) {
            const int par = (int)(k & 1);
            const int jw = L::JB + L::JB_N * par;
            // ---- phase 0: role 0 runs the nominal kinematics of tick k, roles 1 and 2 draw the sample of tick k+1 ----
            if (lead) {
                if (exec) {
                    T u[6];
                    if (SYNTH) {
#pragma unroll
                        for (int cc = 0; cc < 6; ++cc) u[cc] = sm_ld(P.sh + (L::UB + 6 * par + cc) * S);
                    } else {
#pragma unroll
                        for (int cc = 0; cc < 6; ++cc) u[cc] = (T)in.imu_i[(k * 6 + cc) * in.cs];
                    }
                    kin_step(s, accel, u, c, P, jw);
                    ++n_pred;
                }
            } else if (SYNTH) {
                if (adv && k + 1 < k_end) synth_imu_role(a.ns, in.gid, k + 1, a.in.imu + (k + 1) * 6, nb, P, L::UB + 6 * (1 - par));
            }
            gs.sync();
            // ---- prediction (cpp:402-414) in four phases ----
            TickCarry<T> tc;
            if (exec) pred_stage1<BIAS>(P, rp, tc, jw);
            gs.sync();
            if (exec) pred_stage2<BIAS>(P, rp, tc);
            gs.sync();
            if (exec) pred_stage3<BIAS>(P, rp, tc, jw);
            gs.sync();
            if (exec) pred_stage4<BIAS>(P, rp, tc, jw);
            // ---- single-rate correction (cpp:265-279) ----
            if (group_any(perform)) {
                CorrCarry<T, NB> cc;
                Observation<T> obs;
                if (perform) corr_publish(P, lead ? &s : nullptr);
                gs.sync();
                if (perform) corr_stage1<BIAS, DIRECT>(P, tag, rp, obs, cc);
                gs.sync();
                if (perform) {
                    corr_stage2<BIAS>(P, cc);
                    if (lead) {
                        corr_inject<BIAS>(P, s);
#pragma unroll
                        for (int q3 = 0; q3 < 3; ++q3) a.st.aux[(3 + q3) * a.st.ld + i] = obs.r_t_vt_obs[q3];
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) a.st.aux[(6 + q4) * a.st.ld + i] = obs.q_tv_obs[q4];
                        ++n_corr;
                    }
                }
                gs.sync();
            }
            if (exec) {
                if (perform) { upds = 0; flags |= FLAG_CORRECTED; }
                else { upds += 1; flags &= ~FLAG_CORRECTED; }
                flags |= FLAG_ACTIVE;
            }
        }
        if (adv) {
            ++k;
            if (do_stats && (k % a.stats.stride) == 0) at_fence = true;
        }
    }
    // a group that ran out of ticks early (padding, or nothing to do) still attends the remaining sampling points
    for (; rdv_left > 0; --rdv_left) {
        gs.cta_sync();
        for (int g = 0; g < cta.n_groups; ++g) gs.cta_sync();
    }
    gs.sync();                             // the last tick's phase-4 stores are visible to the lanes that store them home
    if (!live) return;

    if (lead && (flags & FLAG_READY) && pend_m >= 0) {
        double tg[7];
        in.tag_f64(a.in, a.ns, pend_m, tg);
#pragma unroll
        for (int cc = 0; cc < 7; ++cc) a.st.pend[cc * a.st.ld + i] = tg[cc];
        a.st.pend[7 * a.st.ld + i] = a.in.tag_stamp[pend_m];
    }
    {
        const int64_t ld = a.st.ld;
        if (lead) {
            T *x = a.st.x + i;
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                x[(0 + cc) * ld] = s.r[cc]; x[(3 + cc) * ld] = s.v[cc]; x[(6 + cc) * ld] = s.q[cc];
                x[(10 + cc) * ld] = s.ab[cc]; x[(13 + cc) * ld] = s.wb[cc];
                a.st.aux[cc * ld + i] = accel[cc];
            }
            x[9 * ld] = s.q[3];
            a.st.flags[i] = flags;
            a.st.upds[i] = upds;
        }
        cov_store(P, a.st.P + i, ld);
    }
    if (a.st.counts && lead) {
#ifdef __CUDA_ARCH__
        atomicAdd(a.st.counts + 0, (unsigned long long)n_pred);
        atomicAdd(a.st.counts + 1, (unsigned long long)n_corr);
        if ((i & 31) == 0) atomicAdd(a.st.counts + 2, (unsigned long long)n_iter);
        atomicAdd(a.st.counts + 4, (unsigned long long)n_sexec);
#else
        a.st.counts[0] += n_pred;
        a.st.counts[1] += n_corr;
#endif
    }
}

#ifdef __CUDACC__
// G groups of 3 warps per CTA; one CTA per SM.  Shared memory: [Lay::SLOTS][32 G] filter words, then the three
// roles' relabelled constants.
template <typename T, bool BIAS, bool DIRECT, bool SYNTH, bool PF, int G>
__global__ void __launch_bounds__(96 * G, 1) run_kernel_coop(const __grid_constant__ RunArgs<T> a)
{
    constexpr int NB = BIAS ? 5 : 3, F = 32 * G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    T *rcs = sm + (size_t)Lay<NB>::SLOTS * F;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = (int)(threadIdx.x & 31);
    const int g = warp / 3, c = warp - 3 * g;
    const int f = g * 32 + lane;
    const int64_t slot = (int64_t)blockIdx.x * F + f;
    const bool live = slot < a.st.n;
    const int64_t i = (live && a.st.perm) ? (int64_t)a.st.perm[slot] : slot;
    if (threadIdx.x < 3) fill_role_consts(a.c, (int)threadIdx.x, rcs + threadIdx.x * RC_N);
    __syncthreads();
    PLane<T, NB, F> P;
    P.setup(sm + f, c);
    const GroupSync gs{ g + 1, nullptr, nullptr, 96 };
    const CtaCtx<T> cta{ g, G, sm + (size_t)Lay<NB>::CB * F };
    const SPtr<T> rc = SPtr<T>::from(rcs + c * RC_N);
    if constexpr (PF) {
        const ParF<T> par = ParSel<T, true>::make(a.c, a.st, live ? i : 0);
        const RotPar<T, ParF<T>> rp{ par, a.c, P.c, (P.c + 1) % 3, (P.c + 2) % 3, rc };
        run_filter_coop<T, BIAS, DIRECT, SYNTH>(a, i, P, rp, live, gs, cta);
    } else {
        const ParS<T> rp{ a.c, rc };
        run_filter_coop<T, BIAS, DIRECT, SYNTH>(a, i, P, rp, live, gs, cta);
    }
}
template <int NB> constexpr size_t coop_smem_bytes(int groups, size_t tsize)
{
    return ((size_t)Lay<NB>::SLOTS * 32 * groups + 3 * RC_N) * tsize;
}
#endif

}  // namespace coop
}  // namespace qekf
