// replay_driver.cpp -- ROS-free replay of RelativePoseEKFNode's call sequence against libqekf (N = 1).
//
// The reference's only caller of the estimator is the ROS node (quad_state_estimation/src/
// relative_pose_EKF_node.cpp): the constructor loads the parameter file and calls initialize_params() (:20-138),
// IMUSubCallback and AprilTagSubCallback write the input members (:144-176), FilterUpdateCallback ticks the filter
// and publishes seven topics from the state members (:178-283).  This driver makes the same calls, in the same
// order, through the member-compatible facade include/relative_pose_ekf_gpu.hpp, fed by one noisy realisation of
// the synthetic landing scenario instead of Gazebo / flight sensors, and writes what the node would publish to a
// CSV trace (one row per timer tick) -- the offline-evaluation workflow the reference covers with rosbag record
// (launch/start_EKF_Cpp_rosbag_record.launch).  It also dumps the input streams it used, so that a checker can
// replay exactly the same inputs through another implementation.
//
//   qekf_replay --preset FILE.yaml [--update-freq HZ] [--measurement-freq HZ] [--tag-rate HZ] [--seconds S]
//               [--latency S] [--seed N] [--single-rate] [--fixed-delay] [--out trace.csv] [--dump-streams PREFIX]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "relative_pose_ekf_gpu.hpp"

namespace {

struct Options {
    std::string preset, out = "trace.csv", dump;
    double update_freq = 0, measurement_freq = 0, tag_rate = 30, seconds = 20, latency = 0.03;
    unsigned long long seed = 0x5EED;
    bool single_rate = false, fixed_delay = false;
};

bool parse(int argc, char **argv, Options *o)
{
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&](double *d) { if (i + 1 >= argc) return false; *d = std::atof(argv[++i]); return true; };
        if (a == "--preset" && i + 1 < argc) o->preset = argv[++i];
        else if (a == "--out" && i + 1 < argc) o->out = argv[++i];
        else if (a == "--dump-streams" && i + 1 < argc) o->dump = argv[++i];
        else if (a == "--update-freq") { if (!val(&o->update_freq)) return false; }
        else if (a == "--measurement-freq") { if (!val(&o->measurement_freq)) return false; }
        else if (a == "--tag-rate") { if (!val(&o->tag_rate)) return false; }
        else if (a == "--seconds") { if (!val(&o->seconds)) return false; }
        else if (a == "--latency") { if (!val(&o->latency)) return false; }
        else if (a == "--seed" && i + 1 < argc) o->seed = std::strtoull(argv[++i], nullptr, 0);
        else if (a == "--single-rate") o->single_rate = true;
        else if (a == "--fixed-delay") o->fixed_delay = true;
        else return false;
    }
    return !o->preset.empty();
}

void check(int rc, const char *what)
{
    if (rc != QEKF_OK) {
        std::fprintf(stderr, "qekf_replay: %s failed: %s\n", what, qekf_last_error_string());
        std::exit(2);
    }
}

}  // namespace

int main(int argc, char **argv)
{
    Options opt;
    if (!parse(argc, argv, &opt)) {
        std::fprintf(stderr, "usage: qekf_replay --preset FILE.yaml [--update-freq HZ] [--measurement-freq HZ] [--tag-rate HZ]\n"
                             "       [--seconds S] [--latency S] [--seed N] [--single-rate] [--fixed-delay] [--out trace.csv]\n"
                             "       [--dump-streams PREFIX]\n");
        return 1;
    }
    try {
        // ---- RelativePoseEKFNode::RelativePoseEKFNode (node.cpp:20-138) ----
        RelativePoseEKF rel_pose_ekf;
        rel_pose_ekf.load_parameter_file(opt.preset);
        if (opt.update_freq > 0) rel_pose_ekf.update_freq = opt.update_freq;
        if (opt.measurement_freq > 0) rel_pose_ekf.measurement_freq = opt.measurement_freq;
        if (opt.single_rate) rel_pose_ekf.multirate_ekf = false;
        if (opt.fixed_delay) rel_pose_ekf.dynamic_meas_delay = false;
        rel_pose_ekf.initialize_params();

        // ---- the sensors: one noisy realisation of the synthetic landing ----
        qekf_params p;
        check(qekf_get_params(rel_pose_ekf.handle(), &p), "qekf_get_params");
        qekf_scenario_spec sc;
        check(qekf_scenario_default(&sc), "qekf_scenario_default");
        sc.duration_s = opt.seconds;
        sc.hover_s = opt.seconds * 0.25;
        sc.tag_rate_hz = opt.tag_rate;
        sc.tag_latency_s = rel_pose_ekf.multirate_ekf ? opt.latency : 0.0;
        int64_t T = 0, M = 0;
        check(qekf_scenario_sizes(&p, &sc, &T, &M), "qekf_scenario_sizes");
        std::vector<double> truth((size_t)(T + 1) * 10), imu_clean((size_t)T * 6), pose_clean((size_t)M * 7), stamp((size_t)M);
        std::vector<int32_t> step((size_t)M);
        check(qekf_scenario_generate(&p, &sc, truth.data(), imu_clean.data(), step.data(), pose_clean.data(), stamp.data()),
              "qekf_scenario_generate");
        qekf_noise_spec noise;
        check(qekf_noise_default(&noise), "qekf_noise_default");
        noise.seed = opt.seed;
        noise.dropout_k0 = (int32_t)(T * 0.45);                    // a tag dropout in the middle of the descent
        noise.dropout_k1 = (int32_t)(T * 0.50);
        qekf_shared_streams sh;
        std::memset(&sh, 0, sizeof sh);
        sh.T = T; sh.imu_clean = imu_clean.data(); sh.M = M; sh.tag_step = step.data();
        sh.tag_pose_clean = pose_clean.data(); sh.tag_stamp = stamp.data(); sh.truth = truth.data(); sh.t_start = sc.t_start;
        std::vector<double> imu((size_t)T * 6), pose((size_t)M * 7), bias(6);
        std::vector<uint8_t> valid((size_t)M);
        check(qekf_synthesize_streams(rel_pose_ekf.handle(), &sh, &noise, 0, 1, imu.data(), pose.data(), valid.data(), bias.data()),
              "qekf_synthesize_streams");
        if (!opt.dump.empty()) {
            FILE *f = std::fopen((opt.dump + "_imu.csv").c_str(), "w");
            FILE *g = std::fopen((opt.dump + "_tag.csv").c_str(), "w");
            if (!f || !g) { std::fprintf(stderr, "qekf_replay: cannot write stream dumps\n"); return 2; }
            for (int64_t k = 0; k < T; ++k) {
                for (int c = 0; c < 6; ++c) std::fprintf(f, c ? ",%.17g" : "%.17g", imu[(size_t)k * 6 + c]);
                std::fprintf(f, "\n");
            }
            for (int64_t m = 0; m < M; ++m) {
                std::fprintf(g, "%d,%.17g,%d", step[(size_t)m], stamp[(size_t)m], (int)valid[(size_t)m]);
                for (int c = 0; c < 7; ++c) std::fprintf(g, ",%.17g", pose[(size_t)m * 7 + c]);
                std::fprintf(g, "\n");
            }
            std::fclose(f); std::fclose(g);
        }

        FILE *out = std::fopen(opt.out.c_str(), "w");
        if (!out) { std::fprintf(stderr, "qekf_replay: cannot open %s\n", opt.out.c_str()); return 2; }
        std::fprintf(out, "tick,t,active,px,py,pz,qx,qy,qz,qw");
        for (int i = 0; i < 36; ++i) std::fprintf(out, ",cov%d", i);
        std::fprintf(out, ",bias_ax,bias_ay,bias_az,bias_wx,bias_wy,bias_wz,vx,vy,vz,ax,ay,az,pred_length,corrected,"
                          "obs_px,obs_py,obs_pz,obs_qx,obs_qy,obs_qz,obs_qw,meas_delay\n");

        int64_t m = 0;
        long corrections = 0;
        std::vector<double> tick_us;
        tick_us.reserve((size_t)T);
        for (int64_t k = 0; k < T; ++k) {
            const double t_now = sc.t_start + (double)k / rel_pose_ekf.update_freq;
            // ---- AprilTagSubCallback (node.cpp:153-176) ----
            if (m < M && step[(size_t)m] == k) {
                if (valid[(size_t)m]) {                                        // detections.size() > 0
                    std::lock_guard<std::mutex> lk(rel_pose_ekf.mtx_apriltag);
                    const double *ps = &pose[(size_t)m * 7];
                    rel_pose_ekf.apriltag_pos << ps[0], ps[1], ps[2];
                    rel_pose_ekf.apriltag_orien.w() = ps[6];
                    rel_pose_ekf.apriltag_orien.x() = ps[3];
                    rel_pose_ekf.apriltag_orien.y() = ps[4];
                    rel_pose_ekf.apriltag_orien.z() = ps[5];
                    rel_pose_ekf.apriltag_time = stamp[(size_t)m];
                    rel_pose_ekf.measurement_ready = true;
                    if (!rel_pose_ekf.state_initialized) rel_pose_ekf.initialize_state(false);
                }
                ++m;
            }
            // ---- IMUSubCallback (node.cpp:144-151) ----
            {
                std::lock_guard<std::mutex> lk(rel_pose_ekf.mtx_IMU);
                const double *u = &imu[(size_t)k * 6];
                rel_pose_ekf.IMU_accel << u[0], u[1], u[2];
                rel_pose_ekf.IMU_ang_vel << u[3], u[4], u[5];
            }
            // ---- FilterUpdateCallback (node.cpp:178-283) ----
            rel_pose_ekf.filter_update(t_now);
            tick_us.push_back(rel_pose_ekf.last_tick_seconds * 1e6);
            std::fprintf(out, "%lld,%.17g,%d", (long long)k, t_now, (int)rel_pose_ekf.filter_active);
            if (!rel_pose_ekf.filter_active) { std::fprintf(out, "\n"); continue; }
            std::fprintf(out, ",%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g", rel_pose_ekf.r_nom(0), rel_pose_ekf.r_nom(1),
                         rel_pose_ekf.r_nom(2), rel_pose_ekf.q_nom.x(), rel_pose_ekf.q_nom.y(), rel_pose_ekf.q_nom.z(),
                         rel_pose_ekf.q_nom.w());
            {   // 6x6 pose covariance, row-major: rows / cols {0-2, 6-8} of cov_pert (node.cpp:203-210)
                const int idx[6] = { 0, 1, 2, 6, 7, 8 };
                for (int a = 0; a < 6; ++a)
                    for (int b = 0; b < 6; ++b) std::fprintf(out, ",%.17g", rel_pose_ekf.cov_pert(idx[a], idx[b]));
            }
            for (int i = 0; i < 3; ++i) std::fprintf(out, ",%.17g", rel_pose_ekf.ab_nom(i) + rel_pose_ekf.ab_static(i));
            for (int i = 0; i < 3; ++i) std::fprintf(out, ",%.17g", rel_pose_ekf.wb_nom(i) + rel_pose_ekf.wb_static(i));
            for (int i = 0; i < 3; ++i) std::fprintf(out, ",%.17g", rel_pose_ekf.v_nom(i));
            for (int i = 0; i < 3; ++i) std::fprintf(out, ",%.17g", rel_pose_ekf.accel_rel(i));
            std::fprintf(out, ",%d,%d", rel_pose_ekf.upds_since_correction, (int)rel_pose_ekf.performed_correction);
            if (rel_pose_ekf.performed_correction) {
                ++corrections;
                std::fprintf(out, ",%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g", rel_pose_ekf.r_t_vt_obs(0),
                             rel_pose_ekf.r_t_vt_obs(1), rel_pose_ekf.r_t_vt_obs(2), rel_pose_ekf.q_tv_obs.x(),
                             rel_pose_ekf.q_tv_obs.y(), rel_pose_ekf.q_tv_obs.z(), rel_pose_ekf.q_tv_obs.w(),
                             rel_pose_ekf.measurement_delay_curr);
            }
            std::fprintf(out, "\n");
        }
        std::fclose(out);
        const double *tr = &truth[(size_t)T * 10];
        std::printf("qekf_replay: %lld ticks, %ld corrections; final position error %.4f %.4f %.4f m; trace -> %s\n",
                    (long long)T, corrections, rel_pose_ekf.r_nom(0) - tr[0], rel_pose_ekf.r_nom(1) - tr[1],
                    rel_pose_ekf.r_nom(2) - tr[2], opt.out.c_str());
        // host latency of RelativePoseEKF::filter_update (one launch + one synchronisation), first 200 ticks left out
        if (tick_us.size() > 400) {
            std::vector<double> v(tick_us.begin() + 200, tick_us.end());
            std::sort(v.begin(), v.end());
            std::printf("qekf_replay: {\"tick_latency_us\": {\"median\": %.2f, \"p99\": %.2f, \"max\": %.2f, \"ticks\": %zu, "
                        "\"tick_period_us\": %.1f}}\n", v[v.size() / 2], v[(size_t)((double)v.size() * 0.99)], v.back(), v.size(),
                        1e6 / rel_pose_ekf.update_freq);
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "qekf_replay: %s\n", e.what());
        return 2;
    }
    return 0;
}
