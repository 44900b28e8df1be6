/*
 * ekf_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See ekf_oracle.h.
 *
 * Dense double-precision restatement of the reference estimator.  Everything is deliberately
 * "boringly literal": full n x n matrices, the reference's formulas in the reference's order
 * (P = F P F^T + W Q W^T; K = P G^T (G P G^T + R_k)^-1 by partial-pivot LU; P^ = (I - K G) P),
 * unbounded history vectors in multirate mode.  No structure is exploited here on purpose: the
 * product's CUDA kernels exploit it, and this file is what they are checked against.
 *
 * Matrices are row-major double arrays with explicit dimensions.  Quaternions are (x,y,z,w) arrays,
 * the storage order of the reference's 16-vector (QSE/src/quaternion_helper.cpp:88-100).
 */
#include "ekf_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NMAX 15
#define NQMAX 12

/* ------------------------------------------------------------------------------------------ */
/* small dense helpers                                                                        */
/* ------------------------------------------------------------------------------------------ */

/* C[m x n] = A[m x k] * B[k x n] */
static void mat_mul(const double *A, const double *B, double *C, int m, int k, int n)
{
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s += A[i * k + l] * B[l * n + j];
            C[i * n + j] = s;
        }
}

static void mat_transpose(const double *A, double *At, int m, int n)
{
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) At[j * m + i] = A[i * n + j];
}

static void mat_identity(double *A, int n)
{
    memset(A, 0, sizeof(double) * (size_t)n * (size_t)n);
    for (int i = 0; i < n; ++i) A[i * n + i] = 1.0;
}

/* write a 3x3 block B (row-major) into A[ld] at (r0,c0) */
static void set_block3(double *A, int ld, int r0, int c0, const double *B)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[(r0 + i) * ld + (c0 + j)] = B[i * 3 + j];
}

/* General inverse by LU with partial pivoting (what Eigen's MatrixXd::inverse() does for a dynamic
 * 6x6: PartialPivLU, then solve against the identity).  relative_pose_EKF.cpp:475. */
static void mat_inverse_lu(const double *A, double *Ainv, int n)
{
    double lu[NMAX * NMAX];
    int perm[NMAX];
    memcpy(lu, A, sizeof(double) * (size_t)n * (size_t)n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = fabs(lu[k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            double v = fabs(lu[i * n + k]);
            if (v > best) { best = v; piv = i; }
        }
        if (piv != k) {
            for (int j = 0; j < n; ++j) {
                double t = lu[k * n + j];
                lu[k * n + j] = lu[piv * n + j];
                lu[piv * n + j] = t;
            }
            int tp = perm[k]; perm[k] = perm[piv]; perm[piv] = tp;
        }
        double d = lu[k * n + k];
        for (int i = k + 1; i < n; ++i) {
            lu[i * n + k] /= d;
            double l = lu[i * n + k];
            for (int j = k + 1; j < n; ++j) lu[i * n + j] -= l * lu[k * n + j];
        }
    }
    /* solve L U X = P I column by column */
    for (int c = 0; c < n; ++c) {
        double y[NMAX];
        for (int i = 0; i < n; ++i) {
            double s = (perm[i] == c) ? 1.0 : 0.0;
            for (int j = 0; j < i; ++j) s -= lu[i * n + j] * y[j];
            y[i] = s;
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < n; ++j) s -= lu[i * n + j] * Ainv[j * n + c];
            Ainv[i * n + c] = s / lu[i * n + i];
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* quaternion helpers: QSE/src/quaternion_helper.cpp:9-100                                    */
/* ------------------------------------------------------------------------------------------ */

/* normalise, then flip sign iff w < -0.75 (quaternion_helper.cpp:61-73) */
void orc_quat_norm(double q[4])
{
    const double q_w_lim = -0.75;
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n > 0.0) { /* Eigen::normalize leaves a zero quaternion alone */
        q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
    }
    if (q[3] < q_w_lim) {
        q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3];
    }
}

/* pure -> unit quaternion (quaternion_helper.cpp:9-33) */
void orc_quat_exp(const double v[3], double q[4])
{
    const double norm_tol = 1e-10;
    double norm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    q[3] = cos(norm / 2);
    if (norm < norm_tol) {
        double s = 1 - pow(norm, 2) / 24;
        for (int i = 0; i < 3; ++i) q[i] = v[i] / 2 * s;
    } else {
        double s = sin(norm / 2);
        for (int i = 0; i < 3; ++i) q[i] = v[i] / norm * s;
    }
    orc_quat_norm(q);
}

/* unit -> pure quaternion (quaternion_helper.cpp:36-58); does not normalise its input */
void orc_quat_log(const double q[4], double v[3])
{
    const double norm_tol = 1e-10;
    double vec_norm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    if (vec_norm < norm_tol) {
        double s = 2 / q[3] * (1 - pow(vec_norm / q[3], 2) / 3);
        for (int i = 0; i < 3; ++i) v[i] = s * q[i];
    } else {
        double phi = 2 * atan2(vec_norm, q[3]);
        double s = phi / vec_norm;
        for (int i = 0; i < 3; ++i) v[i] = s * q[i];
    }
}

/* Hamilton product a (x) b, xyzw storage (Eigen Quaterniond::operator*) */
void orc_quat_mul(const double a[4], const double b[4], double out[4])
{
    double ax = a[0], ay = a[1], az = a[2], aw = a[3];
    double bx = b[0], by = b[1], bz = b[2], bw = b[3];
    double o[4];
    o[3] = aw * bw - ax * bx - ay * by - az * bz;
    o[0] = aw * bx + ax * bw + ay * bz - az * by;
    o[1] = aw * by + ay * bw + az * bx - ax * bz;
    o[2] = aw * bz + az * bw + ax * by - ay * bx;
    memcpy(out, o, sizeof o);
}

static void quat_conj(const double a[4], double out[4])
{
    out[0] = -a[0]; out[1] = -a[1]; out[2] = -a[2]; out[3] = a[3];
}

/* Rotation matrix of a (unit) quaternion, no normalisation (Eigen toRotationMatrix) */
void orc_quat_to_rot(const double q[4], double R[9])
{
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    double twx = tx * w, twy = ty * w, twz = tz * w;
    double txx = tx * x, txy = ty * x, txz = tz * x;
    double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

/* Rodrigues rotation for unit axis and angle (Eigen AngleAxisd::toRotationMatrix) */
static void angle_axis_to_rot(double angle, const double axis[3], double R[9])
{
    double s = sin(angle), c = cos(angle);
    double sa[3] = { s * axis[0], s * axis[1], s * axis[2] };
    double ca[3] = { (1 - c) * axis[0], (1 - c) * axis[1], (1 - c) * axis[2] };
    double t;
    t = ca[0] * axis[1]; R[1] = t - sa[2]; R[3] = t + sa[2];
    t = ca[0] * axis[2]; R[2] = t + sa[1]; R[6] = t - sa[1];
    t = ca[1] * axis[2]; R[5] = t - sa[0]; R[7] = t + sa[0];
    R[0] = ca[0] * axis[0] + c;
    R[4] = ca[1] * axis[1] + c;
    R[8] = ca[2] * axis[2] + c;
}

/* quaternion_helper.cpp:76-85 */
void orc_skew(const double v[3], double S[9])
{
    S[0] = 0;     S[1] = -v[2]; S[2] = v[1];
    S[3] = v[2];  S[4] = 0;     S[5] = -v[0];
    S[6] = -v[1]; S[7] = v[0];  S[8] = 0;
}

static void mat3_vec(const double *R, const double v[3], double out[3])
{
    double o[3];
    for (int i = 0; i < 3; ++i) o[i] = R[i * 3 + 0] * v[0] + R[i * 3 + 1] * v[1] + R[i * 3 + 2] * v[2];
    memcpy(out, o, sizeof o);
}

/* ------------------------------------------------------------------------------------------ */
/* the estimator                                                                              */
/* ------------------------------------------------------------------------------------------ */

typedef struct hist_entry {
    double x[16];
    double u[6];
    double P[NMAX * NMAX];
} hist_entry;

struct orc_filter {
    orc_params p;
    /* derived (initialize_params) */
    double dT_nom;
    int upd_per_meas;
    int num_states;
    int nq;
    int measurement_step_delay;
    double Q[NQMAX * NQMAX];
    double cov_init[NMAX * NMAX];
    double R[36];
    double q_vc[4];
    double C_vc[9];
    /* inputs */
    double IMU_accel[3], IMU_ang_vel[3];
    double apriltag_pos[3], apriltag_orien[4];
    double apriltag_time;
    /* state */
    double r_nom[3], v_nom[3], q_nom[4], ab_nom[3], wb_nom[3];
    double cov_pert[NMAX * NMAX];
    double accel_rel[3];
    double r_t_vt_obs[3], q_tv_obs[4];
    double measurement_delay_curr;
    /* multirate history */
    hist_entry *hist;
    int64_t hist_len, hist_cap;
    /* flags */
    int state_initialized, measurement_ready, performed_correction, filter_active;
    int upds_since_correction;
    /* instrumentation */
    int64_t n_predict, n_correct;
};

/* relative_pose_EKF.cpp:8-85, with the members the constructor leaves uninitialised taken from the
 * node's defaults (relative_pose_EKF_node.cpp:35,64,89-93). */
void orc_default_params(orc_params *p)
{
    memset(p, 0, sizeof *p);
    p->update_freq = 100;
    p->measurement_freq = 10;
    p->measurement_delay = 0.010;
    p->measurement_delay_max = 0.200;
    p->dyn_measurement_delay_offset = 0.0;
    for (int i = 0; i < 3; ++i) {
        p->Q_a[i] = 0.005;
        p->Q_w[i] = 0.0005;
        p->Q_ab[i] = 5e-5;
        p->Q_wb[i] = 5e-6;
    }
    p->R_r[0] = 0.005;  p->R_r[1] = 0.005;  p->R_r[2] = 0.015;
    p->R_ang[0] = 0.0025; p->R_ang[1] = 0.0025; p->R_ang[2] = 0.025;
    p->r_cov_init = 0.1; p->v_cov_init = 0.1; p->ang_cov_init = 0.15;
    p->ab_cov_init = 0.5; p->wb_cov_init = 0.1;
    p->r_v_cv[2] = -0.073;
    p->q_vc[0] = 0.70711; p->q_vc[1] = -0.70711; p->q_vc[2] = 0; p->q_vc[3] = 0;
    p->camera_K[0] = 241.4268; p->camera_K[2] = 376.5;
    p->camera_K[4] = 241.4268; p->camera_K[5] = 240.5;
    p->camera_K[8] = 1;
    p->camera_width = 752; p->camera_height = 480;
    p->n_tags = 1;
    p->tag_in_view_margin = 0.02;
    p->tag_widths[0] = 0.8;
    p->small_ang_tol = 1e-10;
    p->g[2] = -9.8;
    p->est_bias = 1;
    p->limit_measurement_freq = 0;
    p->corner_margin_enbl = 1;
    p->direct_orien_method = 0;
    p->multirate_ekf = 0;
    p->dynamic_meas_delay = 0;
}

int orc_sizeof_params(void) { return (int)sizeof(orc_params); }

/* relative_pose_EKF.cpp:87-125 */
void orc_initialize_params(orc_filter *f)
{
    const orc_params *p = &f->p;
    f->dT_nom = 1 / p->update_freq;
    f->upd_per_meas = (int)ceil(p->update_freq / p->measurement_freq);
    f->num_states = p->est_bias ? 15 : 9;
    {
        int s = (int)(p->measurement_delay / f->dT_nom + 0.5);
        f->measurement_step_delay = s > 1 ? s : 1;
    }
    int n = f->num_states;
    double q_stack[NQMAX], c_stack[NMAX];
    if (p->est_bias) {
        f->nq = 12;
        for (int i = 0; i < 3; ++i) {
            q_stack[i] = p->Q_a[i]; q_stack[3 + i] = p->Q_w[i];
            q_stack[6 + i] = p->Q_ab[i]; q_stack[9 + i] = p->Q_wb[i];
            c_stack[i] = p->r_cov_init; c_stack[3 + i] = p->v_cov_init; c_stack[6 + i] = p->ang_cov_init;
            c_stack[9 + i] = p->ab_cov_init; c_stack[12 + i] = p->wb_cov_init;
        }
    } else {
        f->nq = 6;
        for (int i = 0; i < 3; ++i) {
            q_stack[i] = p->Q_a[i]; q_stack[3 + i] = p->Q_w[i];
            c_stack[i] = p->r_cov_init; c_stack[3 + i] = p->v_cov_init; c_stack[6 + i] = p->ang_cov_init;
        }
    }
    memset(f->Q, 0, sizeof f->Q);
    for (int i = 0; i < f->nq; ++i) f->Q[i * f->nq + i] = q_stack[i];
    memset(f->cov_init, 0, sizeof f->cov_init);
    for (int i = 0; i < n; ++i) f->cov_init[i * n + i] = c_stack[i];
    memcpy(f->cov_pert, f->cov_init, sizeof f->cov_pert);

    memset(f->R, 0, sizeof f->R);
    for (int i = 0; i < 3; ++i) {
        f->R[i * 6 + i] = p->R_r[i];
        f->R[(3 + i) * 6 + (3 + i)] = p->R_ang[i];
    }
    /* camera calibration: quaternion_norm(q_vc); C_vc = R(q_vc); T_vc = [C_vc | r_v_cv] */
    memcpy(f->q_vc, p->q_vc, sizeof f->q_vc);
    orc_quat_norm(f->q_vc);
    orc_quat_to_rot(f->q_vc, f->C_vc);
}

orc_filter *orc_create(const orc_params *p)
{
    orc_filter *f = (orc_filter *)calloc(1, sizeof *f);
    if (!f) return NULL;
    f->p = *p;
    f->apriltag_orien[3] = 1.0;
    f->q_nom[3] = 1.0;
    f->q_tv_obs[3] = 1.0;
    orc_initialize_params(f);
    return f;
}

void orc_destroy(orc_filter *f)
{
    if (!f) return;
    free(f->hist);
    free(f);
}

void orc_set_params(orc_filter *f, const orc_params *p)
{
    f->p = *p;
    orc_initialize_params(f);
}

static void hist_reserve(orc_filter *f, int64_t n)
{
    if (n <= f->hist_cap) return;
    int64_t cap = f->hist_cap ? f->hist_cap : 64;
    while (cap < n) cap *= 2;
    f->hist = (hist_entry *)realloc(f->hist, sizeof(hist_entry) * (size_t)cap);
    f->hist_cap = cap;
}

static void pack_state(const orc_filter *f, double x[16])
{
    memcpy(x + 0, f->r_nom, 3 * sizeof(double));
    memcpy(x + 3, f->v_nom, 3 * sizeof(double));
    memcpy(x + 6, f->q_nom, 4 * sizeof(double));
    memcpy(x + 10, f->ab_nom, 3 * sizeof(double));
    memcpy(x + 13, f->wb_nom, 3 * sizeof(double));
}

static void unpack_state(orc_filter *f, const double x[16])
{
    memcpy(f->r_nom, x + 0, 3 * sizeof(double));
    memcpy(f->v_nom, x + 3, 3 * sizeof(double));
    memcpy(f->q_nom, x + 6, 4 * sizeof(double));
    memcpy(f->ab_nom, x + 10, 3 * sizeof(double));
    memcpy(f->wb_nom, x + 13, 3 * sizeof(double));
}

/* -R(q) (C_vc p + r_v_cv): the reading of  -(q * T_vc * p.homogeneous())  fixed by SURVEY A-8 and
 * by the prototype's explicit form (rel_pose_EKF_test_class.py:355,436). */
static void neg_rot_cam_point(const orc_filter *f, const double q[4], const double p[3], double out[3])
{
    double Rq[9], t[3], o[3];
    orc_quat_to_rot(q, Rq);
    mat3_vec(f->C_vc, p, t);
    for (int i = 0; i < 3; ++i) t[i] += f->p.r_v_cv[i];
    mat3_vec(Rq, t, o);
    for (int i = 0; i < 3; ++i) out[i] = -o[i];
}

/* relative_pose_EKF.cpp:305-344 */
void orc_initialize_state(orc_filter *f, int reinit_bias)
{
    double qq[4];
    orc_quat_mul(f->q_vc, f->apriltag_orien, qq);
    quat_conj(qq, f->q_nom);
    orc_quat_norm(f->q_nom);
    neg_rot_cam_point(f, f->q_nom, f->apriltag_pos, f->r_nom);
    memset(f->v_nom, 0, sizeof f->v_nom);
    if (reinit_bias) {
        memset(f->ab_nom, 0, sizeof f->ab_nom);
        memset(f->wb_nom, 0, sizeof f->wb_nom);
    }
    memcpy(f->cov_pert, f->cov_init, sizeof f->cov_pert);

    /* history <- single entry (x, u = 0, cov_init); biases enter the stack only when estimated */
    hist_reserve(f, 1);
    hist_entry *h = &f->hist[0];
    memset(h, 0, sizeof *h);
    memcpy(h->x + 0, f->r_nom, 3 * sizeof(double));
    memcpy(h->x + 3, f->v_nom, 3 * sizeof(double));
    memcpy(h->x + 6, f->q_nom, 4 * sizeof(double));
    if (f->p.est_bias) {
        memcpy(h->x + 10, f->ab_nom, 3 * sizeof(double));
        memcpy(h->x + 13, f->wb_nom, 3 * sizeof(double));
    }
    memcpy(h->P, f->cov_init, sizeof h->P);
    f->hist_len = 1;
    f->state_initialized = 1;
}

/* relative_pose_EKF.cpp:346-415 */
void orc_prediction_step(orc_filter *f, const double x[16], const double *P, const double u[6],
                         double x_out[16], double *P_out, double accel[3])
{
    const orc_params *p = &f->p;
    const int n = f->num_states, nq = f->nq;
    const double dT = f->dT_nom;
    const double *r = x, *v = x + 3, *q = x + 6, *ab = x + 10, *wb = x + 13;
    double a_nom[3], w_nom[3], C[9];
    f->n_predict++;

    for (int i = 0; i < 3; ++i) {
        a_nom[i] = u[i] - ab[i] - p->ab_static[i];
        w_nom[i] = u[3 + i] - wb[i] - p->wb_static[i];
    }
    orc_quat_to_rot(q, C);

    double acc[3];
    mat3_vec(C, a_nom, acc);
    for (int i = 0; i < 3; ++i) acc[i] += p->g[i];

    /* nominal state */
    double xo[16];
    for (int i = 0; i < 3; ++i) {
        xo[i] = r[i] + dT * v[i];
        xo[3 + i] = v[i] + dT * acc[i];
    }
    double dth[3] = { dT * w_nom[0], dT * w_nom[1], dT * w_nom[2] };
    double qe[4], qn[4];
    orc_quat_exp(dth, qe);
    orc_quat_mul(q, qe, qn);
    orc_quat_norm(qn);
    memcpy(xo + 6, qn, sizeof qn);
    for (int i = 0; i < 3; ++i) { xo[10 + i] = ab[i]; xo[13 + i] = wb[i]; }

    /* Jacobians */
    double F[NMAX * NMAX], W[NMAX * NQMAX];
    mat_identity(F, n);
    memset(W, 0, sizeof W);
    double blk[9], S[9], mdC[9];
    memset(blk, 0, sizeof blk);
    blk[0] = blk[4] = blk[8] = dT;
    set_block3(F, n, 0, 3, blk);
    for (int i = 0; i < 9; ++i) mdC[i] = -dT * C[i];
    orc_skew(a_nom, S);
    mat_mul(mdC, S, blk, 3, 3, 3);
    set_block3(F, n, 3, 6, blk);

    double w_int_angle = sqrt(dth[0] * dth[0] + dth[1] * dth[1] + dth[2] * dth[2]);
    if (w_int_angle < p->small_ang_tol) {
        orc_skew(dth, S);
        for (int i = 0; i < 9; ++i) blk[i] = ((i % 4 == 0) ? 1.0 : 0.0) - S[i];
    } else {
        double axis[3] = { dth[0] / w_int_angle, dth[1] / w_int_angle, dth[2] / w_int_angle };
        angle_axis_to_rot(-w_int_angle, axis, blk);
    }
    set_block3(F, n, 6, 6, blk);

    double mC[9];
    for (int i = 0; i < 9; ++i) mC[i] = -C[i];
    if (p->est_bias) {
        set_block3(F, n, 3, 9, mdC);
        memset(blk, 0, sizeof blk);
        blk[0] = blk[4] = blk[8] = -dT;
        set_block3(F, n, 6, 12, blk);
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) W[(3 + i) * nq + j] = mC[i * 3 + j];
    for (int i = 6; i < n; ++i) W[i * nq + (i - 3)] = 1.0;

    /* P_check = F P F^T + W Q W^T */
    double Ft[NMAX * NMAX], Wt[NQMAX * NMAX], FP[NMAX * NMAX], FPFt[NMAX * NMAX];
    double WQ[NMAX * NQMAX], WQWt[NMAX * NMAX];
    mat_transpose(F, Ft, n, n);
    mat_transpose(W, Wt, n, nq);
    mat_mul(F, P, FP, n, n, n);
    mat_mul(FP, Ft, FPFt, n, n, n);
    mat_mul(W, f->Q, WQ, n, nq, nq);
    mat_mul(WQ, Wt, WQWt, n, nq, n);
    for (int i = 0; i < n * n; ++i) P_out[i] = FPFt[i] + WQWt[i];

    memcpy(x_out, xo, sizeof xo);
    memcpy(accel, acc, sizeof acc);
}

/* relative_pose_EKF.cpp:417-502 */
void orc_correction_step(orc_filter *f, const double x[16], const double *P, const double r_c_tc[3],
                         const double q_ct[4], double x_out[16], double *P_out)
{
    const orc_params *p = &f->p;
    const int n = f->num_states;
    const double *r = x, *v = x + 3, *q = x + 6, *ab = x + 10, *wb = x + 13;
    double C[9], Ct[9];
    f->n_correct++;

    orc_quat_to_rot(q, C);
    mat_transpose(C, Ct, 3, 3);
    {
        double qq[4];
        orc_quat_mul(f->q_vc, q_ct, qq);
        quat_conj(qq, f->q_tv_obs);
        orc_quat_norm(f->q_tv_obs);
    }
    if (p->direct_orien_method)
        neg_rot_cam_point(f, f->q_tv_obs, r_c_tc, f->r_t_vt_obs);
    else
        neg_rot_cam_point(f, q, r_c_tc, f->r_t_vt_obs);

    double dy[6];
    for (int i = 0; i < 3; ++i) dy[i] = f->r_t_vt_obs[i] - r[i];
    {
        double qc[4], dq[4];
        quat_conj(q, qc);
        orc_quat_mul(qc, f->q_tv_obs, dq);
        orc_quat_norm(dq);
        orc_quat_log(dq, dy + 3);
    }

    /* G (6 x n) */
    double G[6 * NMAX], Gt[NMAX * 6];
    memset(G, 0, sizeof G);
    for (int i = 0; i < 3; ++i) G[i * n + i] = 1.0;
    if (!p->direct_orien_method) {
        double Ctr[3], S[9], blk[9];
        mat3_vec(Ct, r, Ctr);
        orc_skew(Ctr, S);
        mat_mul(C, S, blk, 3, 3, 3);
        set_block3(G, n, 0, 6, blk);
    }
    for (int i = 0; i < 3; ++i) G[(3 + i) * n + (6 + i)] = 1.0;
    mat_transpose(G, Gt, 6, n);

    /* N (6 x 6) = [[-C C_vc, direct ? skew(r) : 0], [0, C_vc]] */
    double Nk[36], Nt[36];
    memset(Nk, 0, sizeof Nk);
    {
        double mC[9], blk[9];
        for (int i = 0; i < 9; ++i) mC[i] = -C[i];
        mat_mul(mC, f->C_vc, blk, 3, 3, 3);
        set_block3(Nk, 6, 0, 0, blk);
        set_block3(Nk, 6, 3, 3, f->C_vc);
        if (p->direct_orien_method) {
            orc_skew(r, blk);
            set_block3(Nk, 6, 0, 3, blk);
        }
    }
    mat_transpose(Nk, Nt, 6, 6);
    double NR[36], Rk[36];
    mat_mul(Nk, f->R, NR, 6, 6, 6);
    mat_mul(NR, Nt, Rk, 6, 6, 6);

    /* K = P G^T (G P G^T + R_k)^-1 */
    double GP[6 * NMAX], S6[36], Sinv[36], PGt[NMAX * 6], K[NMAX * 6];
    mat_mul(G, P, GP, 6, n, n);
    mat_mul(GP, Gt, S6, 6, n, 6);
    for (int i = 0; i < 36; ++i) S6[i] += Rk[i];
    mat_inverse_lu(S6, Sinv, 6);
    mat_mul(P, Gt, PGt, n, n, 6);
    mat_mul(PGt, Sinv, K, n, 6, 6);

    /* P_hat = (I - K G) P ;  delta_x = K delta_y */
    double KG[NMAX * NMAX], IKG[NMAX * NMAX], Ph[NMAX * NMAX], dx[NMAX];
    mat_mul(K, G, KG, n, 6, n);
    mat_identity(IKG, n);
    for (int i = 0; i < n * n; ++i) IKG[i] -= KG[i];
    mat_mul(IKG, P, Ph, n, n, n);
    mat_mul(K, dy, dx, n, 6, 1);

    /* inject */
    double xo[16];
    for (int i = 0; i < 3; ++i) {
        xo[i] = r[i] + dx[i];
        xo[3 + i] = v[i] + dx[3 + i];
    }
    {
        double qe[4], qn[4];
        orc_quat_exp(dx + 6, qe);
        orc_quat_mul(q, qe, qn);
        orc_quat_norm(qn);
        memcpy(xo + 6, qn, sizeof qn);
    }
    for (int i = 0; i < 3; ++i) {
        xo[10 + i] = p->est_bias ? ab[i] + dx[9 + i] : 0.0;
        xo[13 + i] = p->est_bias ? wb[i] + dx[12 + i] : 0.0;
    }
    memcpy(x_out, xo, sizeof xo);
    memcpy(P_out, Ph, sizeof(double) * (size_t)n * (size_t)n);
}

/* corner-margin gate, relative_pose_EKF.cpp:156-186 */
static int corner_gate(const orc_filter *f, const double r_c_tc[3], const double q_ct[4])
{
    const orc_params *p = &f->p;
    double Rct[9];
    orc_quat_to_rot(q_ct, Rct);
    int ok = 0;
    for (int i = 0; i < p->n_tags; ++i) {
        double hw = p->tag_widths[i] / 2;
        double px0 = p->tag_positions[3 * i + 0], py0 = p->tag_positions[3 * i + 1];
        double cx[4] = { hw + px0, -hw + px0, -hw + px0, hw + px0 };
        double cy[4] = { hw + py0, hw + py0, -hw + py0, -hw + py0 };
        double min_u = 0, min_v = 0, max_u = 0, max_v = 0;
        for (int c = 0; c < 4; ++c) {
            double pt[3] = { cx[c], cy[c], 0.0 }, pc[3];
            mat3_vec(Rct, pt, pc);
            for (int k = 0; k < 3; ++k) pc[k] += r_c_tc[k];
            double inv_z = 1.0 / pc[2];
            double xn = pc[0] * inv_z, yn = pc[1] * inv_z, zn = pc[2] * inv_z;
            double uu = p->camera_K[0] * xn + p->camera_K[1] * yn + p->camera_K[2] * zn;
            double vv = p->camera_K[3] * xn + p->camera_K[4] * yn + p->camera_K[5] * zn;
            if (c == 0) { min_u = max_u = uu; min_v = max_v = vv; }
            else {
                if (uu < min_u) min_u = uu;
                if (uu > max_u) max_u = uu;
                if (vv < min_v) min_v = vv;
                if (vv > max_v) max_v = vv;
            }
        }
        ok = (min_u > p->camera_width * p->tag_in_view_margin &&
              min_v > p->camera_height * p->tag_in_view_margin &&
              max_u < p->camera_width * (1 - p->tag_in_view_margin) &&
              max_v < p->camera_height * (1 - p->tag_in_view_margin));
        if (ok) break;
    }
    return ok;
}

/* relative_pose_EKF.cpp:127-303 */
void orc_filter_update(orc_filter *f, double t_curr)
{
    const orc_params *p = &f->p;
    const int n = f->num_states;
    if (!f->state_initialized) return;

    double imu[6];
    memcpy(imu, f->IMU_accel, 3 * sizeof(double));
    memcpy(imu + 3, f->IMU_ang_vel, 3 * sizeof(double));

    double r_c_tc[3] = { 0, 0, 0 }, q_ct[4] = { 0, 0, 0, 1 };
    int perform_correction = 0;

    if (f->measurement_ready &&
        (!p->limit_measurement_freq || (f->upds_since_correction + 1) >= f->upd_per_meas)) {
        memcpy(r_c_tc, f->apriltag_pos, sizeof r_c_tc);
        memcpy(q_ct, f->apriltag_orien, sizeof q_ct);
        f->measurement_ready = 0;
        if (p->corner_margin_enbl)
            perform_correction = corner_gate(f, r_c_tc, q_ct);
        else
            perform_correction = 1;
    }

    if (p->multirate_ekf && perform_correction) {
        f->measurement_delay_curr = p->dynamic_meas_delay
            ? fmin(t_curr - f->apriltag_time + p->dyn_measurement_delay_offset, p->measurement_delay_max)
            : p->measurement_delay;
        int step = (int)(f->measurement_delay_curr / f->dT_nom + 0.5);
        if (step < 1) step = 1;
        int64_t ind = f->hist_len - step;
        if (ind < 0) ind = 0;

        double x_hat[16], P_hat[NMAX * NMAX];
        orc_correction_step(f, f->hist[ind].x, f->hist[ind].P, r_c_tc, q_ct, x_hat, P_hat);
        memcpy(f->hist[ind].x, x_hat, sizeof x_hat);
        memcpy(f->hist[ind].P, P_hat, sizeof(double) * (size_t)n * (size_t)n);

        if (ind > 0) {
            memmove(f->hist, f->hist + ind, sizeof(hist_entry) * (size_t)(f->hist_len - ind));
            f->hist_len -= ind;
        }
        for (int64_t i = 1; i < f->hist_len; ++i) {
            double xo[16], Po[NMAX * NMAX], foo[3];
            orc_prediction_step(f, f->hist[i - 1].x, f->hist[i - 1].P, f->hist[i].u, xo, Po, foo);
            memcpy(f->hist[i].x, xo, sizeof xo);
            memcpy(f->hist[i].P, Po, sizeof(double) * (size_t)n * (size_t)n);
        }
        unpack_state(f, f->hist[f->hist_len - 1].x);
        memcpy(f->cov_pert, f->hist[f->hist_len - 1].P, sizeof(double) * (size_t)n * (size_t)n);
    }

    double x_km1[16], x_check[16], P_check[NMAX * NMAX];
    pack_state(f, x_km1);
    orc_prediction_step(f, x_km1, f->cov_pert, imu, x_check, P_check, f->accel_rel);

    if (p->multirate_ekf) {
        hist_reserve(f, f->hist_len + 1);
        hist_entry *h = &f->hist[f->hist_len++];
        memcpy(h->x, x_check, sizeof x_check);
        memcpy(h->u, imu, sizeof imu);
        memcpy(h->P, P_check, sizeof(double) * (size_t)n * (size_t)n);
        unpack_state(f, x_check);
        memcpy(f->cov_pert, P_check, sizeof(double) * (size_t)n * (size_t)n);
    } else if (perform_correction) {
        double x_hat[16], P_hat[NMAX * NMAX];
        orc_correction_step(f, x_check, P_check, r_c_tc, q_ct, x_hat, P_hat);
        unpack_state(f, x_hat);
        memcpy(f->cov_pert, P_hat, sizeof(double) * (size_t)n * (size_t)n);
    } else {
        unpack_state(f, x_check);
        memcpy(f->cov_pert, P_check, sizeof(double) * (size_t)n * (size_t)n);
    }

    if (perform_correction) f->upds_since_correction = 0;
    else f->upds_since_correction += 1;
    f->performed_correction = perform_correction;
    f->filter_active = 1;
}

/* relative_pose_EKF_node.cpp:144-151 */
void orc_set_imu(orc_filter *f, const double accel[3], const double gyro[3])
{
    memcpy(f->IMU_accel, accel, 3 * sizeof(double));
    memcpy(f->IMU_ang_vel, gyro, 3 * sizeof(double));
}

/* relative_pose_EKF_node.cpp:153-176 */
void orc_set_tag(orc_filter *f, const double pos[3], const double quat_xyzw[4], double stamp)
{
    memcpy(f->apriltag_pos, pos, 3 * sizeof(double));
    memcpy(f->apriltag_orien, quat_xyzw, 4 * sizeof(double));
    f->apriltag_time = stamp;
    f->measurement_ready = 1;
    if (!f->state_initialized) orc_initialize_state(f, 0);
}

void orc_get_state(const orc_filter *f, double x16[16]) { pack_state(f, x16); }
int orc_get_num_states(const orc_filter *f) { return f->num_states; }

void orc_get_cov(const orc_filter *f, double *P)
{
    memcpy(P, f->cov_pert, sizeof(double) * (size_t)f->num_states * (size_t)f->num_states);
}

void orc_set_state(orc_filter *f, const double x16[16], const double *P)
{
    unpack_state(f, x16);
    memcpy(f->cov_pert, P, sizeof(double) * (size_t)f->num_states * (size_t)f->num_states);
    f->state_initialized = 1;
    hist_reserve(f, 1);
    memset(&f->hist[0], 0, sizeof(hist_entry));
    memcpy(f->hist[0].x, x16, 16 * sizeof(double));
    memcpy(f->hist[0].P, P, sizeof(double) * (size_t)f->num_states * (size_t)f->num_states);
    f->hist_len = 1;
}

void orc_get_aux(const orc_filter *f, double aux[11])
{
    memcpy(aux, f->accel_rel, 3 * sizeof(double));
    memcpy(aux + 3, f->r_t_vt_obs, 3 * sizeof(double));
    memcpy(aux + 6, f->q_tv_obs, 4 * sizeof(double));
    aux[10] = f->measurement_delay_curr;
}

void orc_get_flags(const orc_filter *f, int32_t fl[6])
{
    fl[0] = f->state_initialized;
    fl[1] = f->measurement_ready;
    fl[2] = f->performed_correction;
    fl[3] = f->filter_active;
    fl[4] = f->upds_since_correction;
    fl[5] = (int32_t)f->hist_len;
}

/* ------------------------------------------------------------------------------------------ */
/* batch replay                                                                               */
/* ------------------------------------------------------------------------------------------ */

struct orc_batch {
    int64_t n;
    orc_filter **f;
};

orc_batch *orc_batch_create(const orc_params *p, int64_t n_filters)
{
    orc_batch *b = (orc_batch *)calloc(1, sizeof *b);
    if (!b) return NULL;
    b->n = n_filters;
    b->f = (orc_filter **)calloc((size_t)n_filters, sizeof(orc_filter *));
    for (int64_t i = 0; i < n_filters; ++i) b->f[i] = orc_create(p);
    return b;
}

void orc_batch_destroy(orc_batch *b)
{
    if (!b) return;
    for (int64_t i = 0; i < b->n; ++i) orc_destroy(b->f[i]);
    free(b->f);
    free(b);
}

orc_filter *orc_batch_filter(orc_batch *b, int64_t i) { return b->f[i]; }

int orc_batch_set_filter_params(orc_batch *b, int field, const double *v)
{
    const int64_t N = b->n;
    for (int64_t i = 0; i < N; ++i) {
        orc_params *p = &b->f[i]->p;
        switch (field) {
        case ORC_PF_Q:
            for (int k = 0; k < 3; ++k) {
                p->Q_a[k] = v[(0 + k) * N + i];
                p->Q_w[k] = v[(3 + k) * N + i];
                p->Q_ab[k] = v[(6 + k) * N + i];
                p->Q_wb[k] = v[(9 + k) * N + i];
            }
            break;
        case ORC_PF_R:
            for (int k = 0; k < 3; ++k) {
                p->R_r[k] = v[(0 + k) * N + i];
                p->R_ang[k] = v[(3 + k) * N + i];
            }
            break;
        case ORC_PF_R_V_CV:
            for (int k = 0; k < 3; ++k) p->r_v_cv[k] = v[k * N + i];
            break;
        case ORC_PF_Q_VC:
            for (int k = 0; k < 4; ++k) p->q_vc[k] = v[k * N + i];
            break;
        case ORC_PF_DELAY:
            p->measurement_delay = v[0 * N + i];
            p->dyn_measurement_delay_offset = v[1 * N + i];
            break;
        default:
            return -1;
        }
        orc_initialize_params(b->f[i]);
    }
    return 0;
}

void orc_batch_run(orc_batch *b, int64_t k0, int64_t n_steps, int64_t T, const double *imu,
                   int64_t M, const int32_t *tag_step, const double *tag_pose, const double *tag_stamp,
                   const uint8_t *tag_valid, double t_start, int n_threads)
{
    const int64_t N = b->n;
    (void)T;
    int64_t m0 = 0;
    while (m0 < M && tag_step[m0] < k0) ++m0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        orc_filter *f = b->f[i];
        int64_t m = m0;
        for (int64_t k = k0; k < k0 + n_steps; ++k) {
            if (m < M && tag_step[m] == k) {
                if (!tag_valid || tag_valid[m * N + i]) {
                    double pos[3], quat[4];
                    for (int c = 0; c < 3; ++c) pos[c] = tag_pose[(m * 7 + c) * N + i];
                    for (int c = 0; c < 4; ++c) quat[c] = tag_pose[(m * 7 + 3 + c) * N + i];
                    orc_set_tag(f, pos, quat, tag_stamp[m]);
                }
                ++m;
            }
            double a[3], w[3];
            for (int c = 0; c < 3; ++c) {
                a[c] = imu[(k * 6 + c) * N + i];
                w[c] = imu[(k * 6 + 3 + c) * N + i];
            }
            orc_set_imu(f, a, w);
            orc_filter_update(f, t_start + (double)k / f->p.update_freq);
        }
    }
}

void orc_batch_get_state(const orc_batch *b, double *x)
{
    const int64_t N = b->n;
    for (int64_t i = 0; i < N; ++i) {
        double xi[16];
        orc_get_state(b->f[i], xi);
        for (int c = 0; c < 16; ++c) x[c * N + i] = xi[c];
    }
}

void orc_batch_get_cov(const orc_batch *b, double *P)
{
    const int64_t N = b->n;
    for (int64_t i = 0; i < N; ++i) {
        const orc_filter *f = b->f[i];
        int nn = f->num_states * f->num_states;
        for (int c = 0; c < nn; ++c) P[c * N + i] = f->cov_pert[c];
    }
}

void orc_batch_get_aux(const orc_batch *b, double *aux)
{
    const int64_t N = b->n;
    for (int64_t i = 0; i < N; ++i) {
        double a[11];
        orc_get_aux(b->f[i], a);
        for (int c = 0; c < 11; ++c) aux[c * N + i] = a[c];
    }
}

void orc_batch_get_flags(const orc_batch *b, int32_t *flags)
{
    const int64_t N = b->n;
    for (int64_t i = 0; i < N; ++i) {
        int32_t fl[6];
        orc_get_flags(b->f[i], fl);
        for (int c = 0; c < 6; ++c) flags[c * N + i] = fl[c];
    }
}

void orc_batch_get_counts(const orc_batch *b, int64_t counts[2])
{
    counts[0] = counts[1] = 0;
    for (int64_t i = 0; i < b->n; ++i) {
        counts[0] += b->f[i]->n_predict;
        counts[1] += b->f[i]->n_correct;
    }
}
