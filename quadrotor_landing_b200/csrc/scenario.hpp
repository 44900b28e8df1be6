// scenario.hpp -- host-side synthetic landing scenario: a hover-and-descend truth trajectory above the
// tag, the clean IMU stream that produces it, and clean tag poses with a capture-to-arrival latency.
//
// The reference has no such generator (it was exercised in Gazebo); this is the
// product's Monte-Carlo input.  The measurement geometry is the inverse of the reference's model:
//   q_ct   = conj(q_tv (x) q_vc)                        so that conj(q_vc (x) q_ct) = q_tv   (cpp:310,431)
//   r_c_tc = C_vc^T ( -R(q_tv)^T r_t - r_v_cv )         so that -R(q_tv)(C_vc r_c + r_v_cv) = r_t (cpp:313,438)
//   u_a    = R(q_tv)^T (a - g) + ab_static,  u_w = w_body + wb_static                         (cpp:357-362)
// and the truth is advanced with the filter's own discrete model (explicit Euler on r, v; q (x) exp(dT w)),
// cpp:365-371, so a noise-free replay converges onto it (initialize_state starts from v = 0).
#pragma once

#include <cmath>
#include <cstdint>

#include "../../include/qekf.h"

namespace qekf {
namespace scenario {

inline void q_mul(const double a[4], const double b[4], double o[4])
{
    double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
    double r[4];
    r[3] = aw * bw - ax * bx - ay * by - az * bz;
    r[0] = aw * bx + ax * bw + ay * bz - az * by;
    r[1] = aw * by + ay * bw + az * bx - ax * bz;
    r[2] = aw * bz + az * bw + ax * by - ay * bx;
    for (int i = 0; i < 4; ++i) o[i] = r[i];
}
inline void q_rot(const double q[4], double R[9])
{
    double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
inline void q_normalize(double q[4])
{
    double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= n;
}

// analytic path: lateral sway, yaw oscillation, hover then smooth-step descent
struct Path {
    const qekf_scenario_spec &s;
    explicit Path(const qekf_scenario_spec &spec) : s(spec) {}
    void pos_vel_acc(double t, double p[3], double v[3], double a[3]) const
    {
        p[0] = s.sway_ax * std::sin(s.sway_wx * t);
        v[0] = s.sway_ax * s.sway_wx * std::cos(s.sway_wx * t);
        a[0] = -s.sway_ax * s.sway_wx * s.sway_wx * std::sin(s.sway_wx * t);
        p[1] = s.sway_ay * std::sin(s.sway_wy * t + s.sway_phase_y);
        v[1] = s.sway_ay * s.sway_wy * std::cos(s.sway_wy * t + s.sway_phase_y);
        a[1] = -s.sway_ay * s.sway_wy * s.sway_wy * std::sin(s.sway_wy * t + s.sway_phase_y);
        const double D = s.duration_s - s.hover_s, dz = s.z_end - s.z_start;
        if (t <= s.hover_s || D <= 0) { p[2] = s.z_start; v[2] = 0; a[2] = 0; }
        else {
            double tau = (t - s.hover_s) / D;
            if (tau > 1) { p[2] = s.z_end; v[2] = 0; a[2] = 0; }
            else {
                p[2] = s.z_start + dz * (3 * tau * tau - 2 * tau * tau * tau);
                v[2] = dz * (6 * tau - 6 * tau * tau) / D;
                a[2] = dz * (6 - 12 * tau) / (D * D);
            }
        }
    }
    double yaw(double t) const { return s.yaw_amp * std::sin(s.yaw_w * t); }
    double yaw_rate(double t) const { return s.yaw_amp * s.yaw_w * std::cos(s.yaw_w * t); }
};

inline void defaults(qekf_scenario_spec *s)
{
    s->duration_s = 60.0; s->hover_s = 10.0;
    s->z_start = 3.0; s->z_end = 1.0;
    s->sway_ax = 0.3; s->sway_wx = 0.4; s->sway_ay = 0.2; s->sway_wy = 0.3; s->sway_phase_y = 1.0;
    s->yaw_amp = 0.3; s->yaw_w = 0.2;
    s->tag_rate_hz = 30.0; s->tag_latency_s = 0.0;
    s->t_start = 0.0;
}

inline void sizes(const qekf_params &p, const qekf_scenario_spec &s, int64_t *T, int64_t *M)
{
    *T = (int64_t)std::floor(s.duration_s * p.update_freq + 0.5);
    const double ratio = p.update_freq / s.tag_rate_hz;
    int64_t m = 0;
    while ((int64_t)std::floor(m * ratio) + 1 < *T) ++m;
    *M = m;
}

// truth [T+1][10] (r, v, q_tv after j ticks), imu [T][6], tag_step [M], tag_pose [M][7], tag_stamp [M]
inline void generate(const qekf_params &p, const qekf_scenario_spec &s, double *truth, double *imu,
                     int32_t *tag_step, double *tag_pose, double *tag_stamp)
{
    int64_t T, M;
    sizes(p, s, &T, &M);
    const double f = p.update_freq, dT = 1 / f;
    Path path(s);
    double q_vc[4] = { p.q_vc[0], p.q_vc[1], p.q_vc[2], p.q_vc[3] };
    q_normalize(q_vc);
    if (q_vc[3] < -0.75) for (int i = 0; i < 4; ++i) q_vc[i] = -q_vc[i];
    double C_vc[9];
    q_rot(q_vc, C_vc);

    double r[3], v[3], a0[3], q[4];
    path.pos_vel_acc(0.0, r, v, a0);
    const double psi0 = path.yaw(0.0);
    q[0] = 0; q[1] = 0; q[2] = std::sin(psi0 / 2); q[3] = std::cos(psi0 / 2);
    for (int64_t k = 0; k <= T; ++k) {
        double *X = truth + k * 10;
        for (int i = 0; i < 3; ++i) { X[i] = r[i]; X[3 + i] = v[i]; }
        for (int i = 0; i < 4; ++i) X[6 + i] = q[i];
        if (k == T) break;
        const double t = k * dT;
        double pp[3], vv[3], acc[3];
        path.pos_vel_acc(t, pp, vv, acc);
        const double wz = path.yaw_rate(t);
        double C[9];
        q_rot(q, C);
        double f_sp[3] = { acc[0] - p.g[0], acc[1] - p.g[1], acc[2] - p.g[2] };
        for (int i = 0; i < 3; ++i) {
            imu[k * 6 + i] = C[0 * 3 + i] * f_sp[0] + C[1 * 3 + i] * f_sp[1] + C[2 * 3 + i] * f_sp[2] + p.ab_static[i];
        }
        imu[k * 6 + 3] = 0 + p.wb_static[0];
        imu[k * 6 + 4] = 0 + p.wb_static[1];
        imu[k * 6 + 5] = wz + p.wb_static[2];
        for (int i = 0; i < 3; ++i) r[i] += dT * v[i];
        for (int i = 0; i < 3; ++i) v[i] += dT * acc[i];
        const double half = 0.5 * dT * wz;
        double qe[4] = { 0, 0, std::sin(half), std::cos(half) }, qn[4];
        q_mul(q, qe, qn);
        q_normalize(qn);
        for (int i = 0; i < 4; ++i) q[i] = qn[i];
    }
    const double ratio = f / s.tag_rate_hz;
    const int64_t lat = (int64_t)std::floor(s.tag_latency_s * f + 0.5);
    for (int64_t m = 0; m < M; ++m) {
        const int64_t k = (int64_t)std::floor(m * ratio) + 1;
        // a detection delivered before tick k's update and captured `lat` ticks ago shows the state
        // after tick k-lat, i.e. truth[k+1-lat]
        int64_t c = k + 1 - lat;
        if (c < 0) c = 0;
        const double *X = truth + c * 10;
        const double *qt = X + 6;
        double qq[4], Rt[9];
        q_mul(qt, q_vc, qq);
        tag_pose[m * 7 + 3] = -qq[0]; tag_pose[m * 7 + 4] = -qq[1]; tag_pose[m * 7 + 5] = -qq[2]; tag_pose[m * 7 + 6] = qq[3];
        q_rot(qt, Rt);
        double w[3];
        for (int i = 0; i < 3; ++i) w[i] = -(Rt[0 * 3 + i] * X[0] + Rt[1 * 3 + i] * X[1] + Rt[2 * 3 + i] * X[2]) - p.r_v_cv[i];
        for (int i = 0; i < 3; ++i) tag_pose[m * 7 + i] = C_vc[0 * 3 + i] * w[0] + C_vc[1 * 3 + i] * w[1] + C_vc[2 * 3 + i] * w[2];
        tag_step[m] = (int32_t)k;
        tag_stamp[m] = s.t_start + (double)(c - 1) * dT;
    }
}

}  // namespace scenario
}  // namespace qekf
