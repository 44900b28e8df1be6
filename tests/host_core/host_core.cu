// host_core.cu -- TEST HARNESS (not product): instantiates the product's per-filter code
// (quadrotor_landing_b200/csrc/ekf_core.cuh, run_filter in ekf_kernels.cuh) for the HOST so that the
// structured arithmetic and the tick sequencer can be checked against the oracle in the CPU-only test
// tier.  The product library never contains or calls this; on a GPU the same templates run inside
// run_kernel.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../quadrotor_landing_b200/csrc/ekf_kernels.cuh"
#include "../../quadrotor_landing_b200/csrc/ekf_params.hpp"

using namespace qekf;

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
    prediction_step<T, BIAS>(s, P, uu, c, a);
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
    correction_step<T, BIAS, DIRECT>(s, P, tg, c, obs);
    pack<T, BIAS>(s, P, xo, Po);
    for (int k = 0; k < 3; ++k) obs7[k] = obs.r_t_vt_obs[k];
    for (int k = 0; k < 4; ++k) obs7[3 + k] = obs.q_tv_obs[k];
}

template <typename T, bool BIAS, bool DIRECT>
void run_t(const qekf_params *p, int64_t N, int64_t k0, int64_t n_steps, const double *imu, int64_t M,
           const int32_t *tag_step, const double *tag_pose, const double *tag_stamp, const uint8_t *tag_valid,
           double t_start, double *x, double *Ppk, double *aux, double *pend, int32_t *flags, int32_t *upds)
{
    constexpr int NS = BIAS ? 15 : 9;
    constexpr int NP = NS * (NS + 1) / 2;
    std::vector<T> xs(16 * N), Ps((size_t)NP * N), as(AUX_DIM * N);
    for (size_t k = 0; k < xs.size(); ++k) xs[k] = (T)x[k];
    for (size_t k = 0; k < Ps.size(); ++k) Ps[k] = (T)Ppk[k];
    for (size_t k = 0; k < as.size(); ++k) as[k] = (T)aux[k];
    RunArgs<T> a;
    a.st.x = xs.data(); a.st.P = Ps.data(); a.st.aux = as.data(); a.st.pend = pend;
    a.st.flags = flags; a.st.upds = upds; a.st.ld = N; a.st.n = N;
    std::memset(&a.in, 0, sizeof a.in);
    a.in.imu = imu; a.in.tag_step = tag_step; a.in.tag_pose = tag_pose; a.in.tag_stamp = tag_stamp;
    a.in.tag_valid = tag_valid; a.in.cs = N; a.in.is = 1; a.in.M = M; a.in.vs = N;
    a.in.t_start = t_start; a.in.update_freq = p->update_freq;
    a.c = make_consts<T>(*p);
    a.k0 = k0; a.n_steps = n_steps;
    int32_t m0 = 0;
    while (m0 < M && tag_step[m0] < k0) ++m0;
    a.m0 = m0;
    for (int64_t i = 0; i < N; ++i) {
        PLocal<T, NS> P;
        run_filter<T, BIAS, DIRECT>(a, i, P);
    }
    for (size_t k = 0; k < xs.size(); ++k) x[k] = xs[k];
    for (size_t k = 0; k < Ps.size(); ++k) Ppk[k] = Ps[k];
    for (size_t k = 0; k < as.size(); ++k) aux[k] = as[k];
}

}  // namespace

#define HC_DISPATCH(prec, p, CALL)                                                                  \
    do {                                                                                            \
        const bool b__ = (p)->est_bias != 0, d__ = (p)->direct_orien_method != 0;                   \
        if ((prec) == 64) {                                                                         \
            if (b__ && d__) { CALL(double, true, true); } else if (b__) { CALL(double, true, false); } \
            else if (d__) { CALL(double, false, true); } else { CALL(double, false, false); }       \
        } else {                                                                                    \
            if (b__ && d__) { CALL(float, true, true); } else if (b__) { CALL(float, true, false); } \
            else if (d__) { CALL(float, false, true); } else { CALL(float, false, false); }         \
        }                                                                                           \
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

}  // extern "C"
