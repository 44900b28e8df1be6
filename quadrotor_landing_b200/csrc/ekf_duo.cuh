// ekf_duo.cuh -- the two-role mapping: TWO warps per 32 filters, each doing a different half of filter_update.
//
// Why.  One thread per filter caps an SM at 7 warps (the packed covariance fills the shared memory) and a warp's
// tick is one long dependent instruction stream.  Splitting a filter over lanes that run the SAME code (ekf_coop.cuh,
// three lanes) replicates the nominal state, the Jacobians and the sequencing in every lane: measured at half the
// speed (profiles/r2_01, r2_02).  Splitting it over two warps that run DIFFERENT code replicates nothing:
//
//   role A (covariance core)    the 9x9 block P11 of (dr, dv, dth): F11 P11 F11^T through the same three congruences
//                               the thread-per-filter kernel uses, then the terms that couple it to the bias columns;
//                               the whole measurement update (correction_step) on correction ticks; statistics samples.
//   role B (nominal + biases)   owns the nominal state: draws the IMU noise, runs the kinematics one tick ahead and
//                               publishes B = -dT C, a, dT w and the coefficients of Phi (18 words); propagates the
//                               9x6 bias columns P12' = F11 P12 + F12 P22, which depend on nothing role A writes.
//
// With F = [F11 F12; 0 I] (relative_pose_EKF.cpp:402-414),
//   P12' = F11 P12 + F12 P22,      P22' = P22 + Q22,
//   P11' = F11 P11 F11^T + N F12^T + (N F12^T)^T - F12 P22 F12^T + Q11,    N = P12',
// so a tick is:  phase 1  A: F11 P11 F11^T + Q11   ||  B: P12 <- N          (disjoint blocks, old values only)
//                phase 2  A: coupling terms, Q22   ||  B: noise + kinematics of the NEXT tick
// with one named barrier (64 threads) after each phase.  Roles are whole warps: no divergence, no replicated
// arithmetic; per-SM residency goes from 7 warps to 12 (192 filters) with 168 registers per thread.
#pragma once

#include "ekf_coop.cuh"     // GroupSync, group votes

namespace qekf {
namespace duo {

using coop::GroupSync;
using coop::GroupVote;
using coop::group_any;
using coop::group_vote;
using coop::PNull;

// exchange words per filter (element-major like the covariance: word w of filter f at X[w*F + f])
enum { X_B = 0, X_A = 9, X_DTH = 12, X_PC = 15, X_J = 18,      // per tick: B (row-major), a, dT w, (cs, s1, s2)
       X_NOM = 0,                                              // cold (aliases the above): the nominal state, 16 words
       X_WORDS = 18 };

template <typename T, int STRIDE> struct XBuf {
    T *base;      // already points at the filter
    QEKF_FN T ld(int w) const { return base[w * STRIDE]; }
    QEKF_FN void st(int w, T v) const { base[w * STRIDE] = v; }
};

template <typename T, class XB> QEKF_FN void nominal_put(const XB &X, const Nominal<T> &s)
{
#pragma unroll
    for (int i = 0; i < 3; ++i) { X.st(X_NOM + i, s.r[i]); X.st(X_NOM + 3 + i, s.v[i]); X.st(X_NOM + 10 + i, s.ab[i]); X.st(X_NOM + 13 + i, s.wb[i]); }
#pragma unroll
    for (int i = 0; i < 4; ++i) X.st(X_NOM + 6 + i, s.q[i]);
}
template <typename T, class XB> QEKF_FN void nominal_get(const XB &X, Nominal<T> &s)
{
#pragma unroll
    for (int i = 0; i < 3; ++i) { s.r[i] = X.ld(X_NOM + i); s.v[i] = X.ld(X_NOM + 3 + i); s.ab[i] = X.ld(X_NOM + 10 + i); s.wb[i] = X.ld(X_NOM + 13 + i); }
#pragma unroll
    for (int i = 0; i < 4; ++i) s.q[i] = X.ld(X_NOM + 6 + i);
}

// ---- role B: nominal kinematics of one tick (cpp:346-401), Jacobian pieces published ----
template <typename T, class XB>
QEKF_FN void kin_publish(Nominal<T> &s, T accel[3], const T u[6], const Consts<T> &c, const XB &X)
{
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
#pragma unroll
    for (int i = 0; i < 9; ++i) X.st(X_B + i, -d * C[i]);
    T dth[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        dth[i] = d * w[i];
        X.st(X_A + i, a[i]);
        X.st(X_DTH + i, dth[i]);
    }
    PhiCoef<T> pc;
    attitude_step(s.q, dth, c.small_ang_tol, pc);
    X.st(X_PC + 0, pc.cs); X.st(X_PC + 1, pc.s1); X.st(X_PC + 2, pc.s2);
}

// the published pieces -> A = B skew(a), Phi (and, for role A, QV = C diag(Q_a) C^T = B diag(Q_a) B^T / dT^2)
template <typename T, class XB, class PAR>
QEKF_FN void jac_get(const XB &X, const PAR &par, PredJac<T> &J, bool want_qv)
{
    const T d = par.c.dT;
    T a[3], dth[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) J.B[i] = X.ld(X_B + i);
#pragma unroll
    for (int i = 0; i < 3; ++i) { a[i] = X.ld(X_A + i); dth[i] = X.ld(X_DTH + i); }
    const PhiCoef<T> pc{ X.ld(X_PC + 0), X.ld(X_PC + 1), X.ld(X_PC + 2) };
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        J.A[i * 3 + 0] = J.B[i * 3 + 1] * a[2] - J.B[i * 3 + 2] * a[1];
        J.A[i * 3 + 1] = J.B[i * 3 + 2] * a[0] - J.B[i * 3 + 0] * a[2];
        J.A[i * 3 + 2] = J.B[i * 3 + 0] * a[1] - J.B[i * 3 + 1] * a[0];
    }
    phi_matrix(pc, dth, J.Phi);
    if (want_qv) {
        const T inv = T(1) / (d * d);
        const T q0 = par.Q(0) * inv, q1 = par.Q(1) * inv, q2 = par.Q(2) * inv;
        int e = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i; j < 3; ++j)
                J.QV[e++] = J.B[i * 3 + 0] * q0 * J.B[j * 3 + 0] + J.B[i * 3 + 1] * q1 * J.B[j * 3 + 1] +
                            J.B[i * 3 + 2] * q2 * J.B[j * 3 + 2];
    }
}

// ---- role B, phase 1: the bias columns  P12 <- F11 P12 + F12 P22  (old values only, in place top-down) ----
template <typename T, class PS> QEKF_FN void bias_columns(PS &P, const PredJac<T> &J, T d)
{
#pragma unroll
    for (int y = 0; y < 2; ++y) {
        const int Y = BAB + y;
        T rY[9], vY[9], tY[9], aY[9], wY[9];
        ldb(P, BR, Y, rY);
        ldb(P, BV, Y, vY);
        ldb(P, BTH, Y, tY);
        ldb(P, BAB, Y, aY);
        ldb(P, BWB, Y, wY);
#pragma unroll
        for (int i = 0; i < 9; ++i) rY[i] = M<T>::fma_(d, vY[i], rY[i]);
        stb(P, BR, Y, rY);
        mm_acc(vY, J.A, tY);
        mm_acc(vY, J.B, aY);
        stb(P, BV, Y, vY);
        T n[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) n[i] = -d * wY[i];
        mm_acc(n, J.Phi, tY);
        stb(P, BTH, Y, n);
    }
}

// ---- role A, phase 2: P11 += N F12^T + (N F12^T)^T - F12 P22 F12^T, then P22 += Q22 ----
//   (r,v) += N_ra B^T          (r,th) += -dT N_rw
//   (v,v) += N_va B^T + B N_va^T - B (ab,ab) B^T
//   (v,th) += -dT N_vw + B N_tha^T + dT B (ab,wb)
//   (th,th) += -dT (N_thw + N_thw^T) - dT^2 (wb,wb)
template <typename T, class PS, class PAR> QEKF_FN void couple_core(PS &P, const T B[9], const PAR &par)
{
    const T d = par.c.dT;
    {
        T n[9], x[9];
        ldb(P, BR, BAB, n);
        ldb(P, BR, BV, x);
        mmt_acc(x, n, B);
        stb(P, BR, BV, x);
        ldb(P, BR, BWB, n);
        ldb(P, BR, BTH, x);
#pragma unroll
        for (int i = 0; i < 9; ++i) x[i] = M<T>::fma_(-d, n[i], x[i]);
        stb(P, BR, BTH, x);
    }
    {
        T nva[9], aa[9], t[9], vv[9];
        ldb(P, BV, BAB, nva);
        ldb(P, BAB, BAB, aa);
        // t = N_va - B aa   ->   vv += t B^T + B N_va^T
#pragma unroll
        for (int i = 0; i < 9; ++i) t[i] = nva[i];
        {
            T baa[9];
            mm_set(baa, B, aa);
#pragma unroll
            for (int i = 0; i < 9; ++i) t[i] -= baa[i];
        }
        ldb(P, BV, BV, vv);
        mmt_acc(vv, t, B);
        mmt_acc(vv, B, nva);
        stb(P, BV, BV, vv);
    }
    {
        T x[9], n[9], aw[9];
        ldb(P, BV, BTH, x);
        ldb(P, BV, BWB, n);
        ldb(P, BAB, BWB, aw);
        // x += -dT N_vw + B (N_tha^T + dT aw)
#pragma unroll
        for (int i = 0; i < 9; ++i) x[i] = M<T>::fma_(-d, n[i], x[i]);
        ldb(P, BTH, BAB, n);
        T m[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) m[a * 3 + b] = M<T>::fma_(d, aw[a * 3 + b], n[b * 3 + a]);
        mm_acc(x, B, m);
        stb(P, BV, BTH, x);
    }
    {
        T n[9];
        ldb(P, BTH, BWB, n);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = a; b < 3; ++b) {
                T x = P.ld(3 * BTH + a, 3 * BTH + b);
                x = M<T>::fma_(-d, n[a * 3 + b] + n[b * 3 + a], x);
                x = M<T>::fma_(-d * d, P.ld(3 * BWB + a, 3 * BWB + b), x);
                P.st(3 * BTH + a, 3 * BTH + b, x);
            }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        P.st(9 + i, 9 + i, P.ld(9 + i, 9 + i) + par.Q(6 + i));
        P.st(12 + i, 12 + i, P.ld(12 + i, 12 + i) + par.Q(9 + i));
    }
}


// cov_pert = diag(cov_init)  (initialize_state, cpp:340-343)
template <typename T, class PS> QEKF_FN void cov_diag_init(PS &P, const Consts<T> &c)
{
    constexpr int N = PS::n;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i; j < N; ++j) P.st(i, j, (i == j) ? c.cov_init[i / 3] : T(0));
}

// correction_step for role A behind a real call: the nominal state travels through the exchange words
template <typename T, bool BIAS, bool DIRECT, class PS, class PAR, class XB>
QEKF_COLD void duo_correction(PS P, const XB X, const T *tag, const PAR par, Observation<T> *obs)
{
    Nominal<T> s;
    nominal_get(X, s);
    T tg[7];
#pragma unroll
    for (int cc = 0; cc < 7; ++cc) tg[cc] = tag[cc];
    Observation<T> o;
    correction_step<T, BIAS, DIRECT>(s, P, tg, par, o);
    nominal_put(X, s);
    *obs = o;
}

// The replay loop of one lane of one role (ROLE_A = covariance core + corrections, ROLE_B = nominal + bias columns).
// Both roles run the same sequencing on identical copies of the integer state, so every decision a barrier hangs on
// is taken identically by the two warps of a group.  filter_update, relative_pose_EKF.cpp:127-303 (single-rate).
enum { ROLE_A = 0, ROLE_B = 1 };

template <typename T, bool BIAS, bool DIRECT, bool SYNTH, bool PF, class PS, class XB>
QEKF_FN void run_filter_duo(const RunArgs<T> &a, const int64_t i_in, PS &P, const XB &X, const int role, const bool live,
                            const GroupSync gs, int *vbuf = nullptr)
{
    const Consts<T> &c = a.c;
    const int64_t i = live ? i_in : 0;
    const typename ParSel<T, PF>::type par = ParSel<T, PF>::make(a.c, a.st, i);
    const int64_t k_end = a.k0 + a.n_steps;
    const bool do_stats = SYNTH && a.stats.acc != nullptr;
    const int32_t patience = c.limit_measurement_freq ? (c.upd_per_meas - 1) : 0;
    const bool ra = role == ROLE_A, rb = !ra;

    Nominal<T> s;                          // role B
    T accel[3] = { T(0), T(0), T(0) };     // role B
    int32_t flags = 0, upds = 0;
    Inputs<T, SYNTH> in;
    int64_t k = k_end;
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) { s.r[cc] = T(0); s.v[cc] = T(0); s.q[cc] = T(0); s.ab[cc] = T(0); s.wb[cc] = T(0); }
    s.q[3] = T(1);
    if (live) {
        const int64_t ld = a.st.ld;
        if (ra) {
            constexpr int NP = PS::n * (PS::n + 1) / 2;
#pragma unroll 8
            for (int e = 0; e < NP; ++e) P.el(e) = a.st.P[e * ld + i];
        } else {
            const T *x = a.st.x + i;
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                s.r[cc] = x[(0 + cc) * ld]; s.v[cc] = x[(3 + cc) * ld]; s.q[cc] = x[(6 + cc) * ld];
                s.ab[cc] = x[(10 + cc) * ld]; s.wb[cc] = x[(13 + cc) * ld];
                accel[cc] = a.st.aux[cc * ld + i];
            }
            s.q[3] = x[9 * ld];
        }
        flags = a.st.flags[i];
        upds = a.st.upds[i];
        in.init(a, i);
        k = a.k0;
    }
    gs.sync();

    uint32_t n_pred = 0, n_corr = 0, n_iter = 0, n_sexec = 0;
    int32_t m = a.m0;
    int32_t next_tag_step = (m < a.in.M) ? a.in.tag_step[m] : INT32_MAX;
    int32_t pend_m = -1;
    int32_t held = 0;
    bool at_fence = false;
    bool j_ready = false;                  // the Jacobian pieces of tick k are in the exchange (role B ran ahead)

    // role B: the noisy IMU sample of tick kk, then the kinematics
    auto kinematics = [&](int64_t kk) {
        double raw[6];
        T u[6];
        in.raw_imu(a.in, kk, raw);
        in.imu(a.ns, kk, raw, u);
        kin_publish(s, accel, u, c, X);
    };

    cta_vote_init(vbuf);
    for (uint32_t iter = 0;; ++iter) {
        const bool active = (k < k_end) && !at_fence;

        // ---- AprilTagSubCallback for the arrival scheduled at tick k (node.cpp:153-176) ----
        if (active && k == next_tag_step) {
            if (in.valid(a.in, a.ns, m, (int32_t)k)) {
                pend_m = m;
                flags |= FLAG_READY;
                if (!(flags & FLAG_INIT)) {       // initialize_state (cpp:305-344): nominal part by B, covariance by A
                    if (rb) {
                        T tg[7];
                        in.tag(a.in, a.ns, m, tg);
                        PNull<T, PS::n> pn;
                        initialize_state<T, BIAS>(s, pn, tg, par, false);
                    } else {
                        cov_diag_init(P, c);
                    }
                    flags |= FLAG_INIT;
                    j_ready = false;
                }
            }
            ++m;
            next_tag_step = (m < a.in.M) ? a.in.tag_step[m] : INT32_MAX;
        }

        const bool want = active && (flags & FLAG_INIT) && (flags & FLAG_READY) &&
                          (!c.limit_measurement_freq || (upds + 1) >= c.upd_per_meas);
        // CTA-wide vote (one barrier per iteration): all groups of the CTA serve their corrections in the same
        // iteration and walk the instruction stream together
        const CtaVote v = cta_vote(vbuf, iter, active, want, want && held >= patience, at_fence || k >= k_end, at_fence);
        if (v.active == 0 && v.at_fence == 0) break;
        ++n_iter;
        if (do_stats && v.fenced == v.lanes && v.at_fence != 0) {
            // every filter of the group is at the sampling point (or finished).  Role B never runs ahead across a
            // sampling point, so its nominal state is the one after tick k-1 and the exchange words are free.
            const bool mine = live && at_fence && (flags & FLAG_INIT);
            if (rb && mine) nominal_put(X, s);
            gs.sync();
            if (ra) {
                Nominal<T> n;
                nominal_get(X, n);
                double tb[6] = { 0, 0, 0, 0, 0, 0 };
                if (SYNTH) true_bias(a.ns, in.gid, tb);
                stats_sample<T, BIAS>(a, i, k - 1, n, P, tb, mine);
                if (mine) ++n_sexec;
            }
            gs.sync();
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
            if (pend_m >= 0) {
                in.tag(a.in, a.ns, pend_m, tag);
            } else {
#pragma unroll
                for (int cc = 0; cc < 7; ++cc) tag[cc] = (T)a.st.pend[cc * a.st.ld + i];
            }
            flags &= ~FLAG_READY;
            perform = c.corner_margin_enbl ? corner_gate<T>(tag, c) : true;
            held = 0;
        }

        if (v.active != 0) {
            // ---- catch-up: lanes whose tick-k kinematics have not been run ahead (first tick, after a correction,
            //      a sampling point or an initialisation) ----
            const bool need = exec && !j_ready;
            if (group_any(need)) {
                if (rb && need) kinematics(k);
                gs.sync();
            }
            // ---- phase 1: A: F11 P11 F11^T + Q11 || B: the bias columns ----
            T Bm[9];
            if (exec) {
                PredJac<T> J;
                jac_get(X, par, J, ra);
                if (ra) {
                    pred_cov<T, false>(P, J, par);
#pragma unroll
                    for (int e = 0; e < 9; ++e) Bm[e] = J.B[e];
                } else if (BIAS) {
                    bias_columns(P, J, c.dT);
                }
            }
            gs.sync();
            // ---- phase 2: A: coupling terms || B: the next tick's noise and kinematics ----
            const bool ahead = exec && !perform && (k + 1 < k_end) && !(do_stats && ((k + 1) % a.stats.stride) == 0);
            if (ra) {
                if (exec) {
                    if (BIAS) couple_core(P, Bm, par);
                    ++n_pred;
                }
            } else if (ahead) {
                kinematics(k + 1);
            }
            // ---- single-rate correction (cpp:265-279), by role A on the whole covariance ----
            if (group_any(perform)) {
                if (rb && perform) nominal_put(X, s);
                gs.sync();
                if (ra && perform) {
                    Observation<T> obs;
                    duo_correction<T, BIAS, DIRECT>(P, X, tag, par, &obs);
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) a.st.aux[(3 + cc) * a.st.ld + i] = obs.r_t_vt_obs[cc];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) a.st.aux[(6 + cc) * a.st.ld + i] = obs.q_tv_obs[cc];
                    ++n_corr;
                }
                gs.sync();
                if (rb && perform) nominal_get(X, s);
                gs.sync();      // B has read the corrected state before anybody reuses the exchange words
            } else {
                gs.sync();
            }
            if (exec) {
                if (perform) { upds = 0; flags |= FLAG_CORRECTED; }
                else { upds += 1; flags &= ~FLAG_CORRECTED; }
                flags |= FLAG_ACTIVE;
                j_ready = ahead;
            }
        }
        if (adv) {
            ++k;
            if (do_stats && (k % a.stats.stride) == 0) at_fence = true;
        }
    }
    if (!live) return;

    const int64_t ld = a.st.ld;
    if (ra) {
        constexpr int NP = PS::n * (PS::n + 1) / 2;
#pragma unroll 8
        for (int e = 0; e < NP; ++e) a.st.P[e * ld + i] = P.el(e);
        if (a.st.counts) {
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
    } else {
        if ((flags & FLAG_READY) && pend_m >= 0) {
            double tg[7];
            in.tag_f64(a.in, a.ns, pend_m, tg);
#pragma unroll
            for (int cc = 0; cc < 7; ++cc) a.st.pend[cc * ld + i] = tg[cc];
            a.st.pend[7 * ld + i] = a.in.tag_stamp[pend_m];
        }
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
}

#ifdef __CUDACC__
// G groups of two warps per CTA, one CTA per SM.  Shared memory: covariance [NP][32 G], exchange [X_WORDS][32 G].
template <typename T, bool BIAS, bool DIRECT, bool SYNTH, bool PF, int G>
__global__ void __launch_bounds__(64 * G, 1) run_kernel_duo(const __grid_constant__ RunArgs<T> a)
{
    constexpr int N = BIAS ? 15 : 9, NP = N * (N + 1) / 2, F = 32 * G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = (int)(threadIdx.x & 31);
    const int g = warp >> 1, role = warp & 1;
    const int f = g * 32 + lane;
    const int64_t slot = (int64_t)blockIdx.x * F + f;
    const bool live = slot < a.st.n;
    const int64_t i = (live && a.st.perm) ? (int64_t)a.st.perm[slot] : slot;
    PShared<T, N, F> P{ sm + f };
    const XBuf<T, F> X{ sm + (size_t)NP * F + f };
    int *vbuf = reinterpret_cast<int *>(sm + (size_t)(NP + X_WORDS) * F);
    const GroupSync gs{ g + 1, nullptr, nullptr, 64 };
    run_filter_duo<T, BIAS, DIRECT, SYNTH, PF>(a, i, P, X, role, live, gs, vbuf);
}
template <int N> constexpr size_t duo_smem_bytes(int groups, size_t tsize)
{
    return ((size_t)(N * (N + 1) / 2) + X_WORDS) * 32 * groups * tsize + VOTE_WORDS * sizeof(int);
}
#endif

}  // namespace duo
}  // namespace qekf
