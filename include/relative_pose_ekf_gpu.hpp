// relative_pose_ekf_gpu.hpp -- header-only, Eigen-free look-alike of the reference estimator class on top of
// the libqekf C ABI (include/qekf.h), backed by a batch of ONE filter on the GPU.
//
// It replaces `#include "relative_pose_EKF.hpp"` (quad_state_estimation/include/relative_pose_EKF.hpp:20-144)
// for the one caller the class has, RelativePoseEKFNode (src/relative_pose_EKF_node.cpp), which drives the
// estimator by direct public-member access.  The member names, their meaning and the three methods are the
// reference's; the Eigen types are replaced by tiny fixed-size look-alikes that support exactly the
// expressions the node uses on them:
//     v << x, y, z;   v(i)   q.w() q.x() q.y() q.z()   M(i, j)   tag_positions.resize(3, n)
// (An Eigen build can keep its node source unchanged; a ROS-free caller such as tools/replay_driver.cpp uses the
// same spellings.)  What the node writes between ticks -- IMU_accel / IMU_ang_vel, the latched tag pose, its
// stamp and measurement_ready -- is pushed to the device by filter_update() / initialize_state(), and the state
// members the node publishes are refreshed after every tick.
#ifndef RELATIVE_POSE_EKF_GPU_HPP
#define RELATIVE_POSE_EKF_GPU_HPP

#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "qekf.h"

namespace qekf_facade {

// `v << a, b, c;`
template <class V> struct CommaInit {
    V &v;
    int i;
    CommaInit &operator,(double x) { v(i++) = x; return *this; }
};

struct Vec {                                   // Eigen::VectorXd / Vector3d stand-in
    std::vector<double> d;
    Vec() {}
    explicit Vec(int n) : d((size_t)n, 0.0) {}
    explicit Vec(const double *p) : d(p, p + 3) {}         // Eigen::Vector3d(ptr), node.cpp:76-79
    void resize(int n) { d.assign((size_t)n, 0.0); }
    int size() const { return (int)d.size(); }
    double &operator()(int i) { return d.at((size_t)i); }
    double operator()(int i) const { return d.at((size_t)i); }
    const double *data() const { return d.data(); }
    double *data() { return d.data(); }
    CommaInit<Vec> operator<<(double x) { (*this)(0) = x; return CommaInit<Vec>{ *this, 1 }; }
};

struct Quat {                                  // Eigen::Quaterniond stand-in
    double c[4] = { 0, 0, 0, 1 };              // storage x, y, z, w
    Quat() {}
    Quat(double w_, double x_, double y_, double z_) { c[0] = x_; c[1] = y_; c[2] = z_; c[3] = w_; }   // Eigen's (w,x,y,z) ctor
    explicit Quat(const double *xyzw) { for (int i = 0; i < 4; ++i) c[i] = xyzw[i]; }                 // Eigen's array ctor, node.cpp:109
    double &x() { return c[0]; }
    double &y() { return c[1]; }
    double &z() { return c[2]; }
    double &w() { return c[3]; }
    double x() const { return c[0]; }
    double y() const { return c[1]; }
    double z() const { return c[2]; }
    double w() const { return c[3]; }
};

struct Mat {                                   // Eigen::MatrixXd stand-in (row-major storage)
    std::vector<double> d;
    int r = 0, c = 0;
    Mat() {}
    Mat(int rows, int cols) { resize(rows, cols); }
    void resize(int rows, int cols) { r = rows; c = cols; d.assign((size_t)rows * (size_t)cols, 0.0); }
    int rows() const { return r; }
    int cols() const { return c; }
    double &operator()(int i, int j) { return d.at((size_t)i * (size_t)c + (size_t)j); }
    double operator()(int i, int j) const { return d.at((size_t)i * (size_t)c + (size_t)j); }
};

}  // namespace qekf_facade

class RelativePoseEKF {
public:
    typedef qekf_facade::Vec Vec;
    typedef qekf_facade::Quat Quat;
    typedef qekf_facade::Mat Mat;

    // RelativePoseEKF::RelativePoseEKF()  (src/relative_pose_EKF.cpp:8-85): defaults, then initialize_params()
    explicit RelativePoseEKF(int device = 0, int precision = QEKF_FP64) : device_(device), precision_(precision)
    {
        qekf_params p;
        ok(qekf_default_params(&p));
        load_members(p);
        ok(qekf_create(&p, 1, device_, precision_, &h_));
        refresh();
    }
    ~RelativePoseEKF() { qekf_destroy(h_); }
    RelativePoseEKF(const RelativePoseEKF &) = delete;
    RelativePoseEKF &operator=(const RelativePoseEKF &) = delete;

    // Load every parameter member from a node parameter file (what node.cpp:35-136 does through the ROS parameter
    // server); the caller still calls initialize_params() afterwards, as the node does (node.cpp:138).
    void load_parameter_file(const std::string &path)
    {
        qekf_params p;
        ok(qekf_params_from_yaml(path.c_str(), &p));
        load_members(p);
    }

    // ---- the reference's three public methods ----
    // Compute convenience values derived from parameters                        (cpp:87-125)
    void initialize_params()
    {
        qekf_params p = gather_members();
        ok(qekf_set_params(h_, &p));
        dT_nom = 1.0 / update_freq;
        num_states = qekf_num_states(h_);
        upd_per_meas = (int)std::ceil(update_freq / measurement_freq);                       // cpp:91
        refresh();
    }
    // Initialize state to last received AprilTag relative pose                  (cpp:305-344)
    // (reads the latched apriltag_pos / apriltag_orien members; measurement_ready is not its business)
    void initialize_state(bool reinit_bias)
    {
        const bool ready = measurement_ready;
        const double q[4] = { apriltag_orien.x(), apriltag_orien.y(), apriltag_orien.z(), apriltag_orien.w() };
        ok(qekf_latch_tag(h_, apriltag_pos.data(), q, apriltag_time));
        ok(qekf_initialize_state(h_, reinit_bias ? 1 : 0));
        refresh();
        measurement_ready = ready;
    }
    // Perform periodic EKF filter update                                        (cpp:127-303)
    // One launch: the tag callback's effect (if a detection is waiting), the IMU sample, the tick, and every member the
    // node reads afterwards in one record (qekf_tick).
    void filter_update(double t_curr)
    {
        const auto t0 = std::chrono::steady_clock::now();
        double rec[QEKF_TICK_RECORD];
        const double q[4] = { apriltag_orien.x(), apriltag_orien.y(), apriltag_orien.z(), apriltag_orien.w() };
        ok(qekf_tick(h_, IMU_accel.data(), IMU_ang_vel.data(), measurement_ready ? 1 : 0, apriltag_pos.data(), q,
                     apriltag_time, t_curr, 1, rec));
        unpack(rec, rec + 16, rec + 16 + 225, rec + 16 + 225 + 11);
        last_tick_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    double last_tick_seconds = 0;      // host time of the last filter_update() call (launch + synchronisation + unpack)

    // ---- members, named as in relative_pose_EKF.hpp:32-133 ----
    std::mutex mtx_IMU, mtx_apriltag, mtx_state;
    // inputs (written by the sensor callbacks, node.cpp:144-176)
    Vec IMU_accel{ 3 }, IMU_ang_vel{ 3 }, apriltag_pos{ 3 };
    Quat apriltag_orien;
    double apriltag_time = 0;
    // state (read by the timer callback, node.cpp:184-281)
    Vec r_nom{ 3 }, v_nom{ 3 }, accel_rel{ 3 }, ab_nom{ 3 }, wb_nom{ 3 }, r_t_vt_obs{ 3 };
    Quat q_nom, q_tv_obs;
    Mat cov_pert;
    Vec ab_static{ 3 }, wb_static{ 3 };
    // parameters
    double update_freq = 0, dT_nom = 0, measurement_freq = 0, measurement_delay = 0, measurement_delay_max = 0;
    double dyn_measurement_delay_offset = 0;
    bool est_bias = true, limit_measurement_freq = false, corner_margin_enbl = true, direct_orien_method = false;
    bool multirate_ekf = false, dynamic_meas_delay = false;
    int upd_per_meas = 0, num_states = 15;
    double measurement_delay_curr = 0;
    double r_cov_init = 0, v_cov_init = 0, ang_cov_init = 0, ab_cov_init = 0, wb_cov_init = 0;
    Vec Q_a{ 3 }, Q_w{ 3 }, Q_ab{ 3 }, Q_wb{ 3 }, R_r{ 3 }, R_ang{ 3 };
    Vec r_v_cv{ 3 };
    Quat q_vc;
    Mat camera_K{ 3, 3 };
    int camera_width = 0, camera_height = 0;
    int n_tags = 0;
    double tag_in_view_margin = 0;
    Vec tag_widths;
    Mat tag_positions;                          // 3 x n_tags, node.cpp:128-136
    // counters / flags
    bool state_initialized = false, measurement_ready = false, performed_correction = false, filter_active = false;
    int upds_since_correction = 0;
    int history_length = 0;                     // x_hist.size()
    double small_ang_tol = 0;
    Vec g{ 3 };

    qekf_handle *handle() { return h_; }        // for callers that also want the batch interface

private:
    static void ok(int rc)
    {
        if (rc != QEKF_OK) throw std::runtime_error(std::string("libqekf: ") + qekf_last_error_string());
    }
    void load_members(const qekf_params &p)
    {
        update_freq = p.update_freq; measurement_freq = p.measurement_freq; measurement_delay = p.measurement_delay;
        measurement_delay_max = p.measurement_delay_max; dyn_measurement_delay_offset = p.dyn_measurement_delay_offset;
        for (int i = 0; i < 3; ++i) {
            Q_a(i) = p.Q_a[i]; Q_w(i) = p.Q_w[i]; Q_ab(i) = p.Q_ab[i]; Q_wb(i) = p.Q_wb[i];
            R_r(i) = p.R_r[i]; R_ang(i) = p.R_ang[i];
            ab_static(i) = p.ab_static[i]; wb_static(i) = p.wb_static[i]; r_v_cv(i) = p.r_v_cv[i]; g(i) = p.g[i];
        }
        r_cov_init = p.r_cov_init; v_cov_init = p.v_cov_init; ang_cov_init = p.ang_cov_init;
        ab_cov_init = p.ab_cov_init; wb_cov_init = p.wb_cov_init;
        q_vc = Quat(p.q_vc);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) camera_K(i, j) = p.camera_K[3 * i + j];
        camera_width = p.camera_width; camera_height = p.camera_height;
        n_tags = p.n_tags; tag_in_view_margin = p.tag_in_view_margin;
        tag_widths.resize(n_tags);
        tag_positions.resize(3, n_tags);
        for (int i = 0; i < n_tags; ++i) {
            tag_widths(i) = p.tag_widths[i];
            for (int j = 0; j < 3; ++j) tag_positions(j, i) = p.tag_positions[3 * i + j];
        }
        small_ang_tol = p.small_ang_tol;
        est_bias = p.est_bias != 0; limit_measurement_freq = p.limit_measurement_freq != 0;
        corner_margin_enbl = p.corner_margin_enbl != 0; direct_orien_method = p.direct_orien_method != 0;
        multirate_ekf = p.multirate_ekf != 0; dynamic_meas_delay = p.dynamic_meas_delay != 0;
    }
    qekf_params gather_members() const
    {
        qekf_params p;
        ok(qekf_default_params(&p));
        if (n_tags < 0 || n_tags > QEKF_MAX_TAGS) throw std::runtime_error("n_tags out of range");
        if (tag_widths.size() < n_tags || tag_positions.cols() < n_tags || (n_tags > 0 && tag_positions.rows() != 3))
            throw std::runtime_error("tag_widths / tag_positions are smaller than n_tags");
        p.update_freq = update_freq; p.measurement_freq = measurement_freq; p.measurement_delay = measurement_delay;
        p.measurement_delay_max = measurement_delay_max; p.dyn_measurement_delay_offset = dyn_measurement_delay_offset;
        for (int i = 0; i < 3; ++i) {
            p.Q_a[i] = Q_a(i); p.Q_w[i] = Q_w(i); p.Q_ab[i] = Q_ab(i); p.Q_wb[i] = Q_wb(i);
            p.R_r[i] = R_r(i); p.R_ang[i] = R_ang(i);
            p.ab_static[i] = ab_static(i); p.wb_static[i] = wb_static(i); p.r_v_cv[i] = r_v_cv(i); p.g[i] = g(i);
        }
        p.r_cov_init = r_cov_init; p.v_cov_init = v_cov_init; p.ang_cov_init = ang_cov_init;
        p.ab_cov_init = ab_cov_init; p.wb_cov_init = wb_cov_init;
        p.q_vc[0] = q_vc.x(); p.q_vc[1] = q_vc.y(); p.q_vc[2] = q_vc.z(); p.q_vc[3] = q_vc.w();
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) p.camera_K[3 * i + j] = camera_K(i, j);
        p.camera_width = camera_width; p.camera_height = camera_height;
        p.n_tags = n_tags; p.tag_in_view_margin = tag_in_view_margin;
        for (int i = 0; i < n_tags; ++i) {
            p.tag_widths[i] = tag_widths(i);
            for (int j = 0; j < 3; ++j) p.tag_positions[3 * i + j] = tag_positions(j, i);
        }
        p.small_ang_tol = small_ang_tol;
        p.est_bias = est_bias; p.limit_measurement_freq = limit_measurement_freq;
        p.corner_margin_enbl = corner_margin_enbl; p.direct_orien_method = direct_orien_method;
        p.multirate_ekf = multirate_ekf; p.dynamic_meas_delay = dynamic_meas_delay;
        return p;
    }
    // copy what the node reads after a tick (node.cpp:184-281) out of the device (set-up paths; a tick gets the same
    // values in its record)
    void refresh()
    {
        double x[16], aux[11], fl[6];
        int32_t fi[6];
        const int n = qekf_num_states(h_);
        std::vector<double> P((size_t)n * (size_t)n);
        ok(qekf_get_state(h_, 0, 1, x));
        ok(qekf_get_cov(h_, 0, 1, P.data()));
        ok(qekf_get_aux(h_, 0, 1, aux));
        ok(qekf_get_flags(h_, 0, 1, fi));
        for (int i = 0; i < 6; ++i) fl[i] = fi[i];
        unpack(x, P.data(), aux, fl);
    }
    void unpack(const double *x, const double *P, const double *aux, const double *fl)
    {
        const int n = qekf_num_states(h_);
        for (int i = 0; i < 3; ++i) {
            r_nom(i) = x[i]; v_nom(i) = x[3 + i]; ab_nom(i) = x[10 + i]; wb_nom(i) = x[13 + i];
            accel_rel(i) = aux[i]; r_t_vt_obs(i) = aux[3 + i];
        }
        q_nom = Quat(x + 6);
        q_tv_obs = Quat(aux + 6);
        measurement_delay_curr = aux[10];
        if (cov_pert.rows() != n) cov_pert.resize(n, n);
        cov_pert.d.assign(P, P + (size_t)n * (size_t)n);
        num_states = n;
        state_initialized = fl[0] != 0; measurement_ready = fl[1] != 0; performed_correction = fl[2] != 0;
        filter_active = fl[3] != 0; upds_since_correction = (int)fl[4]; history_length = (int)fl[5];
    }

    qekf_handle *h_ = nullptr;
    int device_, precision_;
};

#endif  // RELATIVE_POSE_EKF_GPU_HPP
