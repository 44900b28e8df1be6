// ekf_kernels.cuh -- sm_100a kernels: one CUDA thread per filter, covariance staged in shared memory
// (element-major, conflict-free), nominal state in registers, many ticks fused per launch.
//
// run_kernel is RelativePoseEKF::filter_update (relative_pose_EKF.cpp:127-303) iterated n_steps times
// with the node's callback sequencing (relative_pose_EKF_node.cpp:144-182) folded in: before tick k
// the tag arrival scheduled for k is latched (and initialises the filter if it is the first one), then
// the IMU sample of tick k is latched, then the tick runs.
#pragma once

#include "ekf_core.cuh"

namespace qekf {

enum : int32_t {
    FLAG_INIT = 1,        // state_initialized
    FLAG_READY = 2,       // measurement_ready
    FLAG_CORRECTED = 4,   // performed_correction
    FLAG_ACTIVE = 8       // filter_active
};

constexpr int AUX_DIM = 11;   // accel_rel(3) r_t_vt_obs(3) q_tv_obs(4) measurement_delay_curr(1)
constexpr int PEND_DIM = 8;   // pending tag pose(7) + capture stamp(1)

// Device-resident state of N filters, structure-of-arrays with leading dimension ld (>= N, multiple
// of 32) so that every per-component access is a coalesced warp transaction.
template <typename T> struct DeviceState {
    T *x;            // [16][ld]   nominal state
    T *P;            // [NP][ld]   packed upper triangle of cov_pert
    T *aux;          // [11][ld]
    double *pend;    // [8][ld]    latched-but-unconsumed tag pose + stamp
    int32_t *flags;  // [ld]
    int32_t *upds;   // [ld]       upds_since_correction
    int64_t ld;
    int64_t n;
};

// Input streams as seen by the kernel.  Element (k, c) of filter i lives at base[(k*6+c)*cs + i*is]:
// explicit per-filter streams use (cs, is) = (N, 1); one stream shared by all filters uses (1, 0).
struct StreamView {
    const double *imu;
    const int32_t *tag_step;
    const double *tag_pose;
    const double *tag_stamp;
    const uint8_t *tag_valid;
    int64_t cs, is;
    int64_t M;
    int64_t vs;           // stride of tag_valid rows (N)
    double t_start;
    double update_freq;
};

template <typename T> struct RunArgs {
    DeviceState<T> st;
    StreamView in;
    Consts<T> c;
    int64_t k0, n_steps;
    int32_t m0;           // first arrival with tag_step >= k0
};

template <typename T, class PS>
QEKF_FN void load_filter(const DeviceState<T> &st, int64_t i, Nominal<T> &s, PS &P)
{
    constexpr int N = PS::n;
    constexpr int NP = N * (N + 1) / 2;
    const T *x = st.x + i;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s.r[k] = x[(0 + k) * st.ld];
        s.v[k] = x[(3 + k) * st.ld];
        s.ab[k] = x[(10 + k) * st.ld];
        s.wb[k] = x[(13 + k) * st.ld];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) s.q[k] = x[(6 + k) * st.ld];
#pragma unroll 8
    for (int e = 0; e < NP; ++e) P.el(e) = st.P[e * st.ld + i];
}

template <typename T, class PS>
QEKF_FN void store_filter(const DeviceState<T> &st, int64_t i, const Nominal<T> &s, const PS &P)
{
    constexpr int N = PS::n;
    constexpr int NP = N * (N + 1) / 2;
    T *x = st.x + i;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        x[(0 + k) * st.ld] = s.r[k];
        x[(3 + k) * st.ld] = s.v[k];
        x[(10 + k) * st.ld] = s.ab[k];
        x[(13 + k) * st.ld] = s.wb[k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) x[(6 + k) * st.ld] = s.q[k];
#pragma unroll 8
    for (int e = 0; e < NP; ++e) st.P[e * st.ld + i] = P.el(e);
}

// ------------------------------------------------------------------------------------------------
// fused multi-tick replay
// ------------------------------------------------------------------------------------------------
// The per-filter replay loop.  Host-callable so that the CPU-side unit tests (tests/host_core) can run
// the very same code against the oracle without a GPU; the product only ever calls it from run_kernel.
template <typename T, bool BIAS, bool DIRECT, class PS>
QEKF_FN void run_filter(const RunArgs<T> &a, const int64_t i, PS &P)
{
    const Consts<T> &c = a.c;
    Nominal<T> s;
    load_filter<T>(a.st, i, s, P);
    int32_t flags = a.st.flags[i];
    int32_t upds = a.st.upds[i];
    T accel[3] = { a.st.aux[0 * a.st.ld + i], a.st.aux[1 * a.st.ld + i], a.st.aux[2 * a.st.ld + i] };

    int32_t m = a.m0;
    int32_t next_tag_step = (m < a.in.M) ? a.in.tag_step[m] : INT32_MAX;
    int32_t pend_m = -1;     // index of the latched arrival; -1 = latched pose lives in st.pend

    const double *imu_i = a.in.imu + i * a.in.is;
    const double *tag_i = a.in.tag_pose + i * a.in.is;

    // software prefetch of the next tick's IMU sample
    double un[6];
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) un[cc] = imu_i[(a.k0 * 6 + cc) * a.in.cs];

    for (int64_t k = a.k0; k < a.k0 + a.n_steps; ++k) {
        T u[6];
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) u[cc] = (T)un[cc];
        if (k + 1 < a.k0 + a.n_steps) {
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) un[cc] = imu_i[((k + 1) * 6 + cc) * a.in.cs];
        }

        // ---- AprilTagSubCallback for the arrival scheduled at this tick (node.cpp:153-176) ----
        if (k == next_tag_step) {
            const bool valid = a.in.tag_valid ? (a.in.tag_valid[m * a.in.vs + i] != 0) : true;
            if (valid) {
                pend_m = m;
                flags |= FLAG_READY;
                if (!(flags & FLAG_INIT)) {
                    T tag[7];
#pragma unroll
                    for (int cc = 0; cc < 7; ++cc) tag[cc] = (T)tag_i[((int64_t)m * 7 + cc) * a.in.cs];
                    initialize_state<T, BIAS>(s, P, tag, c, false);
                    flags |= FLAG_INIT;
                }
            }
            ++m;
            next_tag_step = (m < a.in.M) ? a.in.tag_step[m] : INT32_MAX;
        }
        if (!(flags & FLAG_INIT)) continue;     // filter_update returns early (cpp:129-130)

        // ---- measurement gating (cpp:147-186) ----
        bool perform = false;
        T tag[7];
        if ((flags & FLAG_READY) && (!c.limit_measurement_freq || (upds + 1) >= c.upd_per_meas)) {
            if (pend_m >= 0) {
#pragma unroll
                for (int cc = 0; cc < 7; ++cc) tag[cc] = (T)tag_i[((int64_t)pend_m * 7 + cc) * a.in.cs];
            } else {
#pragma unroll
                for (int cc = 0; cc < 7; ++cc) tag[cc] = (T)a.st.pend[cc * a.st.ld + i];
            }
            flags &= ~FLAG_READY;
            perform = c.corner_margin_enbl ? corner_gate<T>(tag, c) : true;
        }

        // ---- prediction (cpp:240-249), then single-rate correction (cpp:265-279) ----
        prediction_step<T, BIAS>(s, P, u, c, accel);
        if (perform) {
            Observation<T> obs;
            correction_step<T, BIAS, DIRECT>(s, P, tag, c, obs);
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) a.st.aux[(3 + cc) * a.st.ld + i] = obs.r_t_vt_obs[cc];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) a.st.aux[(6 + cc) * a.st.ld + i] = obs.q_tv_obs[cc];
            upds = 0;
            flags |= FLAG_CORRECTED;
        } else {
            upds += 1;
            flags &= ~FLAG_CORRECTED;
        }
        flags |= FLAG_ACTIVE;
    }

    // a latched, still unconsumed measurement survives the launch in st.pend
    if ((flags & FLAG_READY) && pend_m >= 0) {
#pragma unroll
        for (int cc = 0; cc < 7; ++cc) a.st.pend[cc * a.st.ld + i] = tag_i[((int64_t)pend_m * 7 + cc) * a.in.cs];
        a.st.pend[7 * a.st.ld + i] = a.in.tag_stamp[pend_m];
    }
    store_filter<T>(a.st, i, s, P);
    a.st.flags[i] = flags;
    a.st.upds[i] = upds;
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) a.st.aux[cc * a.st.ld + i] = accel[cc];
}

// fused multi-tick replay (single-rate filter: multirate_ekf = false)
template <typename T, bool BIAS, bool DIRECT, int BLOCK>
__global__ void __launch_bounds__(BLOCK) run_kernel(const __grid_constant__ RunArgs<T> a)
{
    constexpr int N = BIAS ? 15 : 9;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= a.st.n) return;
    PShared<T, N, BLOCK> P{ sm + threadIdx.x };
    run_filter<T, BIAS, DIRECT>(a, i, P);
}

// ------------------------------------------------------------------------------------------------
// single-step kernels (stateless step functions and the tag callback)
// ------------------------------------------------------------------------------------------------
// AprilTagSubCallback with one pose shared by all filters (node.cpp:153-176)
template <typename T, bool BIAS, int BLOCK>
__global__ void __launch_bounds__(BLOCK) deliver_tag_kernel(DeviceState<T> st, Consts<T> c, const double *pose8,
                                                            int force_init, int reinit_bias)
{
    constexpr int N = BIAS ? 15 : 9;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= st.n) return;
    int32_t flags = st.flags[i];
    if (!force_init) {
#pragma unroll
        for (int cc = 0; cc < PEND_DIM; ++cc) st.pend[cc * st.ld + i] = pose8[cc];
        flags |= FLAG_READY;
    }
    if (force_init || !(flags & FLAG_INIT)) {
        PShared<T, N, BLOCK> P{ sm + threadIdx.x };
        Nominal<T> s;
        load_filter<T>(st, i, s, P);
        T tag[7];
        // initialize_state reads the latched apriltag_pos/orien members (cpp:310-313)
#pragma unroll
        for (int cc = 0; cc < 7; ++cc) tag[cc] = (T)st.pend[cc * st.ld + i];
        initialize_state<T, BIAS>(s, P, tag, c, reinit_bias != 0);
        store_filter<T>(st, i, s, P);
        flags |= FLAG_INIT;
    }
    st.flags[i] = flags;
}

// prediction_step applied once to every filter with per-filter inputs u [6][ld] (cpp:346-415)
template <typename T, bool BIAS, int BLOCK>
__global__ void __launch_bounds__(BLOCK) predict_kernel(DeviceState<T> st, Consts<T> c, const double *u_in)
{
    constexpr int N = BIAS ? 15 : 9;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= st.n) return;
    PShared<T, N, BLOCK> P{ sm + threadIdx.x };
    Nominal<T> s;
    load_filter<T>(st, i, s, P);
    T u[6], accel[3];
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) u[cc] = (T)u_in[cc * st.ld + i];
    prediction_step<T, BIAS>(s, P, u, c, accel);
    store_filter<T>(st, i, s, P);
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) st.aux[cc * st.ld + i] = accel[cc];
}

// correction_step applied once to every filter with per-filter tag poses [7][ld] (cpp:417-502)
template <typename T, bool BIAS, bool DIRECT, int BLOCK>
__global__ void __launch_bounds__(BLOCK) correct_kernel(DeviceState<T> st, Consts<T> c, const double *tag_in)
{
    constexpr int N = BIAS ? 15 : 9;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= st.n) return;
    PShared<T, N, BLOCK> P{ sm + threadIdx.x };
    Nominal<T> s;
    load_filter<T>(st, i, s, P);
    T tag[7];
#pragma unroll
    for (int cc = 0; cc < 7; ++cc) tag[cc] = (T)tag_in[cc * st.ld + i];
    Observation<T> obs;
    correction_step<T, BIAS, DIRECT>(s, P, tag, c, obs);
    store_filter<T>(st, i, s, P);
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) st.aux[(3 + cc) * st.ld + i] = obs.r_t_vt_obs[cc];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) st.aux[(6 + cc) * st.ld + i] = obs.q_tv_obs[cc];
}

}  // namespace qekf
