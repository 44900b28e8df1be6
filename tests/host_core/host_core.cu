// host_core.cu -- TEST HARNESS (not product): instantiates the product's per-filter code
// (quadrotor_landing_b200/csrc/ekf_core.cuh, run_filter in ekf_kernels.cuh) for the HOST so that the
// structured arithmetic and the tick sequencer can be checked against the oracle in the CPU-only test
// tier.  The product library never contains or calls this; on a GPU the same templates run inside
// run_kernel.
#include <cmath>
#include <cstdlib>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../quadrotor_landing_b200/csrc/ekf_kernels.cuh"
#include "../../quadrotor_landing_b200/csrc/ekf_params.hpp"

using namespace qekf;

// ---- instrumented real type: counts the floating-point operations the structured code executes -------
// (FMA = 2; add, sub, mul, div, sqrt, rsqrt = 1; sincos = 2; atan2 = 1).  Used to derive the
// "algorithmic flops per filter-step" that bench.py's roofline uses.
struct Cnt {
    double v;
    static long long flops;
    Cnt() : v(0) {}
    Cnt(double x) : v(x) {}
    Cnt(int x) : v(x) {}
    explicit operator double() const { return v; }
};
long long Cnt::flops = 0;
inline Cnt operator+(Cnt a, Cnt b) { Cnt::flops += 1; return Cnt(a.v + b.v); }
inline Cnt operator-(Cnt a, Cnt b) { Cnt::flops += 1; return Cnt(a.v - b.v); }
inline Cnt operator*(Cnt a, Cnt b) { Cnt::flops += 1; return Cnt(a.v * b.v); }
inline Cnt operator/(Cnt a, Cnt b) { Cnt::flops += 1; return Cnt(a.v / b.v); }
inline Cnt operator-(Cnt a) { return Cnt(-a.v); }
inline Cnt &operator+=(Cnt &a, Cnt b) { Cnt::flops += 1; a.v += b.v; return a; }
inline Cnt &operator-=(Cnt &a, Cnt b) { Cnt::flops += 1; a.v -= b.v; return a; }
inline Cnt &operator*=(Cnt &a, Cnt b) { Cnt::flops += 1; a.v *= b.v; return a; }
inline bool operator<(Cnt a, Cnt b) { return a.v < b.v; }
inline bool operator>(Cnt a, Cnt b) { return a.v > b.v; }
inline bool operator<=(Cnt a, Cnt b) { return a.v <= b.v; }
inline bool operator>=(Cnt a, Cnt b) { return a.v >= b.v; }
namespace qekf {
template <> struct M<Cnt> {
    static Cnt sqrt_(Cnt x) { Cnt::flops += 1; return Cnt(std::sqrt(x.v)); }
    static Cnt rsqrt_(Cnt x) { Cnt::flops += 1; return Cnt(1.0 / std::sqrt(x.v)); }
    static void sincos_(Cnt x, Cnt *s, Cnt *c) { Cnt::flops += 2; *s = Cnt(std::sin(x.v)); *c = Cnt(std::cos(x.v)); }
    static Cnt atan2_(Cnt y, Cnt x) { Cnt::flops += 1; return Cnt(std::atan2(y.v, x.v)); }
    static Cnt fma_(Cnt a, Cnt b, Cnt c) { Cnt::flops += 2; return Cnt(a.v * b.v + c.v); }
    static Cnt abs_(Cnt x) { return Cnt(std::fabs(x.v)); }
    static Cnt min_(Cnt a, Cnt b) { return Cnt(std::fmin(a.v, b.v)); }
};
}  // namespace qekf

namespace {

template <typename T, bool BIAS> void unpack(const double *x, const double *Pfull, Nominal<T> &s, PLocal<T, (BIAS ? 15 : 9)> &P)
{
    constexpr int N = BIAS ? 15 : 9;
    for (int k = 0; k < 3; ++k) { s.r[k] = (T)x[k]; s.v[k] = (T)x[3 + k]; s.ab[k] = (T)x[10 + k]; s.wb[k] = (T)x[13 + k]; }
    for (int k = 0; k < 4; ++k) s.q[k] = (T)x[6 + k];
    for (int i = 0; i < N; ++i)
        for (int j = i; j < N; ++j) P.st(i, j, (T)Pfull[i * N + j]);
}
template <typename T, bool BIAS> void pack(const Nominal<T> &s, const PLocal<T, (BIAS ? 15 : 9)> &P, double *x, double *Pfull)
{
    constexpr int N = BIAS ? 15 : 9;
    for (int k = 0; k < 3; ++k) { x[k] = s.r[k]; x[3 + k] = s.v[k]; x[10 + k] = s.ab[k]; x[13 + k] = s.wb[k]; }
    for (int k = 0; k < 4; ++k) x[6 + k] = s.q[k];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) Pfull[i * N + j] = P.ld(i, j);
}

template <typename T, bool BIAS> void predict_t(const qekf_params *p, const double *x, const double *Pin, const double *u,
                                                 double *xo, double *Po, double *acc)
{
    Consts<T> c = make_consts<T>(*p);
    Nominal<T> s;
    PLocal<T, (BIAS ? 15 : 9)> P;
    unpack<T, BIAS>(x, Pin, s, P);
    T uu[6], a[3];
    for (int k = 0; k < 6; ++k) uu[k] = (T)u[k];
    prediction_step<T, BIAS>(s, P, uu, ParU<T>{ c }, a);
    pack<T, BIAS>(s, P, xo, Po);
    for (int k = 0; k < 3; ++k) acc[k] = a[k];
}

template <typename T, bool BIAS, bool DIRECT> void correct_t(const qekf_params *p, const double *x, const double *Pin,
                                                              const double *tag, double *xo, double *Po, double *obs7)
{
    Consts<T> c = make_consts<T>(*p);
    Nominal<T> s;
    PLocal<T, (BIAS ? 15 : 9)> P;
    unpack<T, BIAS>(x, Pin, s, P);
    T tg[7];
    for (int k = 0; k < 7; ++k) tg[k] = (T)tag[k];
    Observation<T> obs;
    correction_step<T, BIAS, DIRECT>(s, P, tg, ParU<T>{ c }, obs);
    pack<T, BIAS>(s, P, xo, Po);
    for (int k = 0; k < 3; ++k) obs7[k] = obs.r_t_vt_obs[k];
    for (int k = 0; k < 4; ++k) obs7[3 + k] = obs.q_tv_obs[k];
}

struct McArgs {           // Monte-Carlo (synthetic-noise) extras; ns == nullptr selects explicit streams
    const NoiseSpec *ns;
    const double *truth;
    double *stats_acc;    // [STAT_REPL][n_bins][STAT_DIM]
    int32_t n_bins, stride;
};

// delayed-fusion state and per-filter overrides of a host batch (all optional)
struct ExtArgs {
    double *xc, *Pc, *ring;          // [16][N], [NP][N], [L][6][N]
    int32_t *nh, *hpos, *hlen;       // [N]
    int32_t ring_len, dmax;
    const double *pf[5];             // per-filter overrides [dim][N] per QEKF_PF_* field, or all NULL
};

template <typename T, bool BIAS, bool DIRECT>
void run_t(const qekf_params *p, int64_t N, int64_t k0, int64_t n_steps, const double *imu, int64_t M,
           const int32_t *tag_step, const double *tag_pose, const double *tag_stamp, const uint8_t *tag_valid,
           double t_start, double *x, double *Ppk, double *aux, double *pend, int32_t *flags, int32_t *upds,
           const McArgs *mc = nullptr, const ExtArgs *ext = nullptr, bool lazy = false)
{
    constexpr int NS = BIAS ? 15 : 9;
    constexpr int NP = NS * (NS + 1) / 2;
    std::vector<T> xs(16 * N), Ps((size_t)NP * N), as(AUX_DIM * N);
    for (size_t k = 0; k < xs.size(); ++k) xs[k] = (T)x[k];
    for (size_t k = 0; k < Ps.size(); ++k) Ps[k] = (T)Ppk[k];
    for (size_t k = 0; k < as.size(); ++k) as[k] = (T)aux[k];
    RunArgs<T> a;
    std::memset(&a, 0, sizeof a);
    a.st.x = xs.data(); a.st.P = Ps.data(); a.st.aux = as.data(); a.st.pend = pend;
    a.st.flags = flags; a.st.upds = upds; a.st.counts = nullptr; a.st.ld = N; a.st.n = N;
    std::memset(&a.in, 0, sizeof a.in);
    a.in.imu = imu; a.in.tag_step = tag_step; a.in.tag_pose = tag_pose; a.in.tag_stamp = tag_stamp;
    a.in.tag_valid = tag_valid; a.in.cs = N; a.in.is = 1; a.in.M = M; a.in.vs = N;
    a.in.t_start = t_start; a.in.update_freq = p->update_freq;
    a.c = make_consts<T>(*p);
    a.k0 = k0; a.n_steps = n_steps;
    int32_t m0 = 0;
    while (m0 < M && tag_step[m0] < k0) ++m0;
    a.m0 = m0;
    std::vector<uint8_t> mask;
    std::vector<double> sig;
    if (mc && mc->ns) {
        a.in.cs = 1; a.in.is = 0; a.in.vs = 0; a.in.tag_valid = nullptr;
        a.ns = *mc->ns;
        if (a.ns.edge_loss) {
            mask.resize((size_t)M);
            visibility_mask(tag_pose, M, make_consts<double>(*p), mask.data());
            a.in.tag_valid = mask.data();
        }
        if (a.ns.range_ref > 0) {
            sig.resize((size_t)M * 2);
            range_sigmas(tag_pose, M, a.ns, sig.data());
            a.in.tag_sigma = sig.data();
        }
        if (mc->stats_acc && mc->truth) {
            a.stats.acc = mc->stats_acc; a.stats.truth = mc->truth; a.stats.n_bins = mc->n_bins; a.stats.stride = mc->stride;
            a.stats.chi2_lo = BIAS ? 6.262137795043251 : 2.7003894999803584;
            a.stats.chi2_hi = BIAS ? 27.488392863442982 : 19.02276779864163;
        }
    }
    const bool mr = p->multirate_ekf != 0;
    const bool pf = ext && ext->pf[0];
    std::vector<T> xcs, Pcs, rings, pft;
    std::vector<double> pfd;
    if (mr) {
        xcs.resize(16 * N); Pcs.resize((size_t)NP * N); rings.resize((size_t)ext->ring_len * 6 * N);
        for (size_t k = 0; k < xcs.size(); ++k) xcs[k] = (T)ext->xc[k];
        for (size_t k = 0; k < Pcs.size(); ++k) Pcs[k] = (T)ext->Pc[k];
        for (size_t k = 0; k < rings.size(); ++k) rings[k] = (T)ext->ring[k];
        a.st.xc = xcs.data(); a.st.Pc = Pcs.data(); a.st.ring = rings.data();
        a.st.nh = ext->nh; a.st.hpos = ext->hpos; a.st.hlen = ext->hlen;
        a.st.ring_len = ext->ring_len; a.st.dmax_m1 = ext->dmax - 1;
    }
    if (pf) {
        pft.assign((size_t)PF_DIM * N, T(0));
        pfd.assign(2 * (size_t)N, 0.0);
        qekf_params q = *p;
        for (int64_t i = 0; i < N; ++i) {
            for (int k = 0; k < 3; ++k) {
                q.Q_a[k] = ext->pf[0][(0 + k) * N + i]; q.Q_w[k] = ext->pf[0][(3 + k) * N + i];
                q.Q_ab[k] = ext->pf[0][(6 + k) * N + i]; q.Q_wb[k] = ext->pf[0][(9 + k) * N + i];
                q.R_r[k] = ext->pf[1][(0 + k) * N + i]; q.R_ang[k] = ext->pf[1][(3 + k) * N + i];
                q.r_v_cv[k] = ext->pf[2][k * N + i];
            }
            for (int k = 0; k < 4; ++k) q.q_vc[k] = ext->pf[3][k * N + i];
            fill_pf_column<T>(q, pft.data() + i, N);
            pfd[i] = ext->pf[4][i]; pfd[N + i] = ext->pf[4][N + i];
        }
        a.st.pf = pft.data(); a.st.pf_delay = pfd.data();
    }
    const bool synth = mc && mc->ns;
    for (int64_t i = 0; i < N; ++i) {
        PLocal<T, NS> P;
        int32_t mr_ints[MR_SCRATCH_INTS > SR_SCRATCH_INTS ? MR_SCRATCH_INTS : SR_SCRATCH_INTS];
#define RUN_(S, F)                                                         \
        do {                                                               \
            if (mr && S && lazy) run_filter_mrs<T, BIAS, DIRECT, F>(a, i, P, mr_ints, 1); \
            else if (mr) run_filter_mr<T, BIAS, DIRECT, S, F>(a, i, P, mr_ints, 1); \
            else run_filter<T, BIAS, DIRECT, S, F>(a, i, P, mr_ints, 1);   \
        } while (0)
        if (synth && pf) RUN_(true, true);
        else if (synth) RUN_(true, false);
        else if (pf) RUN_(false, true);
        else RUN_(false, false);
#undef RUN_
    }
    for (size_t k = 0; k < xs.size(); ++k) x[k] = xs[k];
    for (size_t k = 0; k < Ps.size(); ++k) Ppk[k] = Ps[k];
    for (size_t k = 0; k < as.size(); ++k) aux[k] = as[k];
    if (mr) {
        for (size_t k = 0; k < xcs.size(); ++k) ext->xc[k] = xcs[k];
        for (size_t k = 0; k < Pcs.size(); ++k) ext->Pc[k] = Pcs[k];
        for (size_t k = 0; k < rings.size(); ++k) ext->ring[k] = rings[k];
    }
}

}  // namespace

// flops executed by one prediction_step / one correction_step of the structured product code, on a
// typical (large-angle-branch) input.  out[0] = predict, out[1] = correct.
template <bool BIAS, bool DIRECT> static void count_t(const qekf_params *p, double *out)
{
    Consts<double> cd = make_consts<double>(*p);
    Consts<Cnt> c;
    {   // Consts<Cnt> has the same field order; convert the real-valued fields one by one
        const double *src = reinterpret_cast<const double *>(&cd);
        Cnt *dst = reinterpret_cast<Cnt *>(&c);
        const size_t nreal = offsetof(Consts<double>, n_tags) / sizeof(double);
        for (size_t i = 0; i < nreal; ++i) dst[i] = Cnt(src[i]);
        c.n_tags = cd.n_tags; c.upd_per_meas = cd.upd_per_meas; c.limit_measurement_freq = cd.limit_measurement_freq;
        c.corner_margin_enbl = cd.corner_margin_enbl; c.dynamic_meas_delay = cd.dynamic_meas_delay;
    }
    constexpr int N = BIAS ? 15 : 9;
    Nominal<Cnt> s;
    PLocal<Cnt, N> P;
    for (int k = 0; k < 3; ++k) { s.r[k] = Cnt(0.1 * (k + 1)); s.v[k] = Cnt(0.05 * k); s.ab[k] = Cnt(0.01); s.wb[k] = Cnt(0.001); }
    s.r[2] = Cnt(2.5);
    s.q[0] = Cnt(0.05); s.q[1] = Cnt(-0.02); s.q[2] = Cnt(0.15); s.q[3] = Cnt(std::sqrt(1 - 0.05 * 0.05 - 0.02 * 0.02 - 0.15 * 0.15));
    for (int i = 0; i < N; ++i)
        for (int j = i; j < N; ++j) P.st(i, j, Cnt(i == j ? 0.1 : 0.001 * (i + j)));
    Cnt u[6] = { Cnt(0.1), Cnt(-0.2), Cnt(9.7), Cnt(0.01), Cnt(-0.02), Cnt(0.05) }, acc[3];
    Cnt::flops = 0;
    prediction_step<Cnt, BIAS>(s, P, u, ParU<Cnt>{ c }, acc);
    out[0] = (double)Cnt::flops;
    // a tag pose close to the prediction
    Cnt tag[7] = { Cnt(0.05), Cnt(-0.1), Cnt(2.4), Cnt(0.70), Cnt(-0.71), Cnt(0.03), Cnt(0.02) };
    Observation<Cnt> obs;
    Cnt::flops = 0;
    correction_step<Cnt, BIAS, DIRECT>(s, P, tag, ParU<Cnt>{ c }, obs);
    out[1] = (double)Cnt::flops;
}

// Built twice (in parallel): -DHC_ONLY_PREC=64 -> libhost_core64.so, -DHC_ONLY_PREC=32 -> libhost_core32.so.
#ifndef HC_ONLY_PREC
#define HC_ONLY_PREC 64
#endif
#if HC_ONLY_PREC == 64
typedef double hc_real;
#else
typedef float hc_real;
#endif
#define HC_DISPATCH(prec, p, CALL)                                                                  \
    do {                                                                                            \
        const bool b__ = (p)->est_bias != 0, d__ = (p)->direct_orien_method != 0;                   \
        if ((prec) != HC_ONLY_PREC) std::abort();                                                   \
        if (b__ && d__) { CALL(hc_real, true, true); } else if (b__) { CALL(hc_real, true, false); } \
        else if (d__) { CALL(hc_real, false, true); } else { CALL(hc_real, false, false); }         \
    } while (0)

extern "C" {

void hc_prediction_step(const qekf_params *p, int prec, const double *x, const double *P, const double *u,
                        double *xo, double *Po, double *acc)
{
#define C_(T, B, D) predict_t<T, B>(p, x, P, u, xo, Po, acc)
    HC_DISPATCH(prec, p, C_);
#undef C_
}

void hc_correction_step(const qekf_params *p, int prec, const double *x, const double *P, const double *tag,
                        double *xo, double *Po, double *obs7)
{
#define C_(T, B, D) correct_t<T, B, D>(p, x, P, tag, xo, Po, obs7)
    HC_DISPATCH(prec, p, C_);
#undef C_
}

// state arrays: x [16][N], Ppk [NP][N] packed upper triangle, aux [11][N], pend [8][N], flags [N], upds [N]
void hc_run(const qekf_params *p, int prec, int64_t N, int64_t k0, int64_t n_steps, const double *imu, int64_t M,
            const int32_t *tag_step, const double *tag_pose, const double *tag_stamp, const uint8_t *tag_valid,
            double t_start, double *x, double *Ppk, double *aux, double *pend, int32_t *flags, int32_t *upds)
{
#define C_(T, B, D) run_t<T, B, D>(p, N, k0, n_steps, imu, M, tag_step, tag_pose, tag_stamp, tag_valid, t_start, x, Ppk, aux, pend, flags, upds)
    HC_DISPATCH(prec, p, C_);
#undef C_
}

// hc_run with delayed-fusion state and/or per-filter overrides
void hc_run_ext(const qekf_params *p, int prec, int64_t N, int64_t k0, int64_t n_steps, const double *imu, int64_t M,
                const int32_t *tag_step, const double *tag_pose, const double *tag_stamp, const uint8_t *tag_valid,
                double t_start, double *x, double *Ppk, double *aux, double *pend, int32_t *flags, int32_t *upds,
                double *xc, double *Pc, double *ring, int32_t *nh, int32_t *hpos, int32_t *hlen, int32_t ring_len,
                int32_t dmax, const double *pf_q, const double *pf_r, const double *pf_rvcv, const double *pf_qvc,
                const double *pf_delay)
{
    ExtArgs e = { xc, Pc, ring, nh, hpos, hlen, ring_len, dmax, { pf_q, pf_r, pf_rvcv, pf_qvc, pf_delay } };
#define C_(T, B, D) run_t<T, B, D>(p, N, k0, n_steps, imu, M, tag_step, tag_pose, tag_stamp, tag_valid, t_start, x, Ppk, aux, pend, flags, upds, nullptr, &e)
    HC_DISPATCH(prec, p, C_);
#undef C_
}

// ring geometry the product derives for (params, per-filter delays): out[0] = dmax, out[1] = ring length
void hc_ring_geometry(const qekf_params *p, const double *pf_delay, int64_t N, int32_t *out)
{
    int d = step_of_delay(p->dynamic_meas_delay ? p->measurement_delay_max : p->measurement_delay, p->update_freq);
    if (!p->dynamic_meas_delay && pf_delay)
        for (int64_t i = 0; i < N; ++i) {
            const int di = step_of_delay(pf_delay[i], p->update_freq);
            if (di > d) d = di;
        }
    out[0] = d;
    out[1] = ring_length(d, *p);
}

void hc_count_flops(const qekf_params *p, double *out)
{
    const bool b = p->est_bias != 0, d = p->direct_orien_method != 0;
    if (b && d) count_t<true, true>(p, out);
    else if (b) count_t<true, false>(p, out);
    else if (d) count_t<false, true>(p, out);
    else count_t<false, false>(p, out);
}

// Monte-Carlo replay: shared clean streams imu [T][6], tag_pose [M][7]; noise generated per filter.
// noise: the product's device NoiseSpec filled from the qekf_noise_spec fields.
void hc_run_mc(const qekf_params *p, int prec, int64_t N, int64_t k0, int64_t n_steps, const double *imu_clean, int64_t M,
               const int32_t *tag_step, const double *tag_pose_clean, const double *tag_stamp, const double *truth,
               const qekf_noise_spec *n, double *stats_acc, int32_t n_bins, int32_t stride, double t_start, double *x,
               double *Ppk, double *aux, double *pend, int32_t *flags, int32_t *upds)
{
    NoiseSpec ns;
    std::memset(&ns, 0, sizeof ns);
    ns.seed = n->seed; ns.gid0 = n->first_global_id;
    ns.sig_a = (float)n->sigma_accel; ns.sig_w = (float)n->sigma_gyro;
    ns.sig_ba = (float)n->sigma_bias_accel; ns.sig_bw = (float)n->sigma_bias_gyro;
    ns.sig_p = (float)n->sigma_tag_pos; ns.sig_th = (float)n->sigma_tag_ang;
    ns.drop_k0 = n->dropout_k0; ns.drop_k1 = n->dropout_k1;
    ns.rdrop_len = n->rand_dropout_len; ns.rdrop_lo = n->rand_dropout_lo; ns.rdrop_hi = n->rand_dropout_hi;
    ns.edge_loss = n->edge_loss; ns.range_ref = n->range_ref; ns.range_exp_p = n->range_exp_pos; ns.range_exp_th = n->range_exp_ang;
    McArgs mc = { &ns, truth, stats_acc, n_bins, stride };
#define C_(T, B, D) run_t<T, B, D>(p, N, k0, n_steps, imu_clean, M, tag_step, tag_pose_clean, tag_stamp, nullptr, t_start, x, Ppk, aux, pend, flags, upds, &mc)
    HC_DISPATCH(prec, p, C_);
#undef C_
}

// Monte-Carlo replay with delayed fusion (and optional per-filter overrides).  lazy != 0: the loop that re-synthesises the
// history inputs instead of keeping them in the ring (run_filter_mrs); the caller guarantees that the entries after
// the checkpoints came from Monte-Carlo launches of the same noise model and scenario.
void hc_run_mc_ext(const qekf_params *p, int prec, int64_t N, int64_t k0, int64_t n_steps, const double *imu_clean, int64_t M,
                   const int32_t *tag_step, const double *tag_pose_clean, const double *tag_stamp, const double *truth,
                   const qekf_noise_spec *n, double *stats_acc, int32_t n_bins, int32_t stride, double t_start, double *x,
                   double *Ppk, double *aux, double *pend, int32_t *flags, int32_t *upds, double *xc, double *Pc, double *ring,
                   int32_t *nh, int32_t *hpos, int32_t *hlen, int32_t ring_len, int32_t dmax, const double *pf_q,
                   const double *pf_r, const double *pf_rvcv, const double *pf_qvc, const double *pf_delay, int lazy)
{
    NoiseSpec ns;
    std::memset(&ns, 0, sizeof ns);
    ns.seed = n->seed; ns.gid0 = n->first_global_id;
    ns.sig_a = (float)n->sigma_accel; ns.sig_w = (float)n->sigma_gyro;
    ns.sig_ba = (float)n->sigma_bias_accel; ns.sig_bw = (float)n->sigma_bias_gyro;
    ns.sig_p = (float)n->sigma_tag_pos; ns.sig_th = (float)n->sigma_tag_ang;
    ns.drop_k0 = n->dropout_k0; ns.drop_k1 = n->dropout_k1;
    ns.rdrop_len = n->rand_dropout_len; ns.rdrop_lo = n->rand_dropout_lo; ns.rdrop_hi = n->rand_dropout_hi;
    ns.edge_loss = n->edge_loss; ns.range_ref = n->range_ref; ns.range_exp_p = n->range_exp_pos; ns.range_exp_th = n->range_exp_ang;
    McArgs mc = { &ns, truth, stats_acc, n_bins, stride };
    ExtArgs e = { xc, Pc, ring, nh, hpos, hlen, ring_len, dmax, { pf_q, pf_r, pf_rvcv, pf_qvc, pf_delay } };
#define C_(T, B, D) run_t<T, B, D>(p, N, k0, n_steps, imu_clean, M, tag_step, tag_pose_clean, tag_stamp, nullptr, t_start, x, Ppk, aux, pend, flags, upds, &mc, &e, lazy != 0)
    HC_DISPATCH(prec, p, C_);
#undef C_
}

// the product's generator, host-instantiated: explicit streams for filters [first, first+count)
void hc_synthesize(const qekf_params *p, const qekf_noise_spec *n, int64_t T, const double *imu_clean, int64_t M, const int32_t *tag_step,
                   const double *tag_pose_clean, int64_t first, int64_t count, double *imu_out, double *tag_out,
                   uint8_t *valid_out, double *bias_out)
{
    RunArgs<double> a;
    std::memset(&a, 0, sizeof a);
    a.in.imu = imu_clean; a.in.tag_pose = tag_pose_clean; a.in.tag_step = tag_step; a.in.cs = 1; a.in.is = 0; a.in.M = M;
    a.ns.seed = n->seed; a.ns.gid0 = n->first_global_id;
    a.ns.sig_a = (float)n->sigma_accel; a.ns.sig_w = (float)n->sigma_gyro;
    a.ns.sig_ba = (float)n->sigma_bias_accel; a.ns.sig_bw = (float)n->sigma_bias_gyro;
    a.ns.sig_p = (float)n->sigma_tag_pos; a.ns.sig_th = (float)n->sigma_tag_ang;
    a.ns.drop_k0 = n->dropout_k0; a.ns.drop_k1 = n->dropout_k1;
    a.ns.rdrop_len = n->rand_dropout_len; a.ns.rdrop_lo = n->rand_dropout_lo; a.ns.rdrop_hi = n->rand_dropout_hi;
    a.ns.edge_loss = n->edge_loss; a.ns.range_ref = n->range_ref; a.ns.range_exp_p = n->range_exp_pos; a.ns.range_exp_th = n->range_exp_ang;
    std::vector<uint8_t> mask;
    std::vector<double> sig;
    if (p && a.ns.edge_loss) {
        mask.resize((size_t)M);
        visibility_mask(tag_pose_clean, M, make_consts<double>(*p), mask.data());
        a.in.tag_valid = mask.data();
    }
    if (a.ns.range_ref > 0) {
        sig.resize((size_t)M * 2);
        range_sigmas(tag_pose_clean, M, a.ns, sig.data());
        a.in.tag_sigma = sig.data();
    }
    for (int64_t j = 0; j < count; ++j) {
        Inputs<double, true> in;
        in.init(a, first + j);
        for (int c = 0; c < 6; ++c) bias_out[c * count + j] = in.bias[c];
        for (int64_t k = 0; k < T; ++k) {
            double raw[6], u[6];
            in.raw_imu(k, raw);
            synth_imu(a.ns, in.gid, k, raw, in.bias, u);
            for (int c = 0; c < 6; ++c) imu_out[(k * 6 + c) * count + j] = u[c];
        }
        for (int32_t m = 0; m < M; ++m) {
            double tg[7];
            in.tag_f64(a.in, a.ns, m, tg);
            for (int c = 0; c < 7; ++c) tag_out[((int64_t)m * 7 + c) * count + j] = tg[c];
            valid_out[(int64_t)m * count + j] = in.valid(a.in, a.ns, m, tag_step[m]) ? 1 : 0;
        }
    }
}

}  // extern "C"
