/*
 * ekf_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, dense, double-precision restatement of the relative-pose error-state EKF of
 * mbrymer/quadrotor_landing (quad_state_estimation).  It follows, function by function:
 *
 *   QSE/src/relative_pose_EKF.cpp:8-85     constructor defaults          -> orc_default_params
 *   QSE/src/relative_pose_EKF.cpp:87-125   initialize_params             -> orc_initialize_params
 *   QSE/src/relative_pose_EKF.cpp:127-303  filter_update                 -> orc_filter_update
 *   QSE/src/relative_pose_EKF.cpp:305-344  initialize_state              -> orc_initialize_state
 *   QSE/src/relative_pose_EKF.cpp:346-415  prediction_step               -> orc_prediction_step
 *   QSE/src/relative_pose_EKF.cpp:417-502  correction_step               -> orc_correction_step
 *   QSE/src/quaternion_helper.cpp:9-100    quaternion exp/log/norm/skew  -> orc_quat_*
 *   QSE/src/relative_pose_EKF_node.cpp:144-182  callback sequencing      -> orc_set_imu/orc_set_tag
 *
 * (QSE = /root/reference/quad_state_estimation.)  The third-party dependency the reference leans on
 * is Eigen "3.4+" (un-pinned, un-vendored, absent from this image); the Eigen primitives used at the
 * call sites (Quaterniond::toRotationMatrix, Hamilton product, AngleAxisd::toRotationMatrix,
 * Transform composition, PartialPivLU inverse) are restated from their published definitions.
 *
 * PARITY PINNING: the reference ships no golden vectors / KATs.  This oracle is
 * pinned against outputs of the reference's own Python prototype (QSE/test/rel_pose_EKF_test_class.py),
 * imported in the build container by tests/golden/make_golden.py and committed under tests/golden/ (npz files).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * anything under oracle/.  The product (quadrotor_landing_b200/) never links or calls it.
 */
#ifndef EKF_ORACLE_H
#define EKF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_TAGS 16

/* Same field order as the product's qekf_params (include/qekf.h); a test asserts equal sizeof. */
typedef struct orc_params {
    double update_freq;
    double measurement_freq;
    double measurement_delay;
    double measurement_delay_max;
    double dyn_measurement_delay_offset;
    double Q_a[3], Q_w[3], Q_ab[3], Q_wb[3];
    double R_r[3], R_ang[3];
    double r_cov_init, v_cov_init, ang_cov_init, ab_cov_init, wb_cov_init;
    double ab_static[3], wb_static[3];
    double r_v_cv[3];
    double q_vc[4];                 /* x,y,z,w (node.cpp:108-110) */
    double camera_K[9];             /* row-major (node.cpp:115-117) */
    double tag_in_view_margin;
    double tag_widths[ORC_MAX_TAGS];
    double tag_positions[3 * ORC_MAX_TAGS]; /* 3 per tag (node.cpp:128-136) */
    double small_ang_tol;
    double g[3];
    int32_t camera_width, camera_height;
    int32_t n_tags;
    int32_t est_bias;
    int32_t limit_measurement_freq;
    int32_t corner_margin_enbl;
    int32_t direct_orien_method;
    int32_t multirate_ekf;
    int32_t dynamic_meas_delay;
    int32_t reserved;
} orc_params;

typedef struct orc_filter orc_filter;

/* per-filter parameter override ids (batch sweeps, BASELINE config 5) */
enum {
    ORC_PF_Q = 0,        /* 12 doubles: Q_a,Q_w,Q_ab,Q_wb */
    ORC_PF_R = 1,        /* 6 doubles: R_r,R_ang */
    ORC_PF_R_V_CV = 2,   /* 3 doubles */
    ORC_PF_Q_VC = 3,     /* 4 doubles x,y,z,w */
    ORC_PF_DELAY = 4     /* 2 doubles: measurement_delay, dyn_measurement_delay_offset */
};

void orc_default_params(orc_params *p);
int  orc_sizeof_params(void);

/* ---- quaternion helpers (xyzw arrays) ---- */
void orc_quat_exp(const double v[3], double q[4]);
void orc_quat_log(const double q[4], double v[3]);
void orc_quat_norm(double q[4]);
void orc_quat_mul(const double a[4], const double b[4], double out[4]);
void orc_quat_to_rot(const double q[4], double R[9]); /* row-major */
void orc_skew(const double v[3], double S[9]);

/* ---- single filter with the reference's estimator interface ---- */
orc_filter *orc_create(const orc_params *p);
void orc_destroy(orc_filter *f);
void orc_set_params(orc_filter *f, const orc_params *p);      /* overwrite + initialize_params */
void orc_initialize_params(orc_filter *f);
void orc_initialize_state(orc_filter *f, int reinit_bias);
void orc_set_imu(orc_filter *f, const double accel[3], const double gyro[3]);
/* mirrors AprilTagSubCallback: latch pose+stamp, measurement_ready=true, first call initialises */
void orc_set_tag(orc_filter *f, const double pos[3], const double quat_xyzw[4], double stamp);
void orc_filter_update(orc_filter *f, double t_curr);

void orc_get_state(const orc_filter *f, double x16[16]);
int  orc_get_num_states(const orc_filter *f);
void orc_get_cov(const orc_filter *f, double *P /* n*n row-major */);
void orc_set_state(orc_filter *f, const double x16[16], const double *P /* n*n */);
/* aux: accel_rel(3), r_t_vt_obs(3), q_tv_obs(4 xyzw), measurement_delay_curr(1) */
void orc_get_aux(const orc_filter *f, double aux11[11]);
/* flags: state_initialized, measurement_ready, performed_correction, filter_active,
 *        upds_since_correction, history length */
void orc_get_flags(const orc_filter *f, int32_t flags6[6]);

/* stateless step functions (use f's parameters only; write observation side effects into f) */
void orc_prediction_step(orc_filter *f, const double x[16], const double *P, const double u[6],
                         double x_out[16], double *P_out, double accel[3]);
void orc_correction_step(orc_filter *f, const double x[16], const double *P, const double r_c_tc[3],
                         const double q_ct_xyzw[4], double x_out[16], double *P_out);

/* ---- batch of independent filters replaying explicit streams (OpenMP over filters) ----
 * imu       [T][6][N]   accel xyz, gyro xyz
 * tag_step  [M]         tick index at which arrival m becomes visible (strictly increasing)
 * tag_pose  [M][7][N]   r_c_tc xyz, q_ct xyzw
 * tag_stamp [M]         capture time of arrival m (seconds)
 * tag_valid [M][N]      0 = no detection for that filter (may be NULL = all valid)
 * tick k (absolute index, k in [k0, k0+n_steps)) runs at t_curr = t_start + k/update_freq.
 */
typedef struct orc_batch orc_batch;
orc_batch *orc_batch_create(const orc_params *p, int64_t n_filters);
void orc_batch_destroy(orc_batch *b);
int  orc_batch_set_filter_params(orc_batch *b, int field, const double *values /* [dim][N] */);
void orc_batch_run(orc_batch *b, int64_t k0, int64_t n_steps, int64_t T, const double *imu,
                   int64_t M, const int32_t *tag_step, const double *tag_pose, const double *tag_stamp,
                   const uint8_t *tag_valid, double t_start, int n_threads);
void orc_batch_get_state(const orc_batch *b, double *x /* [16][N] */);
void orc_batch_get_cov(const orc_batch *b, double *P /* [n*n][N] row-major element index */);
void orc_batch_get_aux(const orc_batch *b, double *aux /* [11][N] */);
void orc_batch_get_flags(const orc_batch *b, int32_t *flags /* [6][N] */);
orc_filter *orc_batch_filter(orc_batch *b, int64_t i);

/* instrumentation: number of prediction_step / correction_step calls since creation (all filters) */
void orc_batch_get_counts(const orc_batch *b, int64_t counts2[2]);

#ifdef __cplusplus
}
#endif
#endif
