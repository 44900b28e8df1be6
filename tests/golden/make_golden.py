#!/usr/bin/env python
"""Generate golden vectors from the REFERENCE's own Python prototype.

The reference ships no golden vectors or known-answer tests and its C++ path
cannot be compiled in this image (Eigen >= 3.4 absent).  What can run here is the reference's
numpy prototype of the same filter:

    /root/reference/quad_state_estimation/test/rel_pose_EKF_test_class.py   (RelativePoseEKF)
    /root/reference/quad_state_estimation/test/quaternion_helper.py

It imports rospy / tf / ROS message packages, none of which exist here, so this script injects
attribute-bag stub modules, restates the five `tf.transformations` functions the prototype uses
(xyzw convention, from the published ROS tf library), and shims `numpy.math` (removed in NumPy 2).
The prototype's own arithmetic (prediction_step, correction_step, filter_update) runs UNMODIFIED.

Run in the build container only (it reads /root/reference):

    python tests/golden/make_golden.py            # rewrites tests/golden/prototype_vectors.npz

The output is committed; tests never read /root/reference.
"""
from __future__ import annotations

import math
import os
import sys
import types

import numpy as np

REF_TEST_DIR = "/root/reference/quad_state_estimation/test"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "prototype_vectors.npz")


# ------------------------------------------------------------------------------------------------
# stubs
# ------------------------------------------------------------------------------------------------
class _Bag(object):
    """Attribute bag standing in for a ROS message: kwargs become attributes, anything else is a Bag."""

    def __init__(self, *args, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __getattr__(self, name):  # only called for missing attributes
        if name.startswith("__"):
            raise AttributeError(name)
        b = _Bag()
        object.__setattr__(self, name, b)
        return b


def _tf_transformations():
    m = types.ModuleType("tf.transformations")
    _EPS = np.finfo(float).eps * 4.0

    def quaternion_matrix(quaternion):
        q = np.array(quaternion[:4], dtype=np.float64, copy=True)
        nq = np.dot(q, q)
        if nq < _EPS:
            return np.identity(4)
        q *= math.sqrt(2.0 / nq)
        q = np.outer(q, q)
        return np.array((
            (1.0 - q[1, 1] - q[2, 2], q[0, 1] - q[2, 3], q[0, 2] + q[1, 3], 0.0),
            (q[0, 1] + q[2, 3], 1.0 - q[0, 0] - q[2, 2], q[1, 2] - q[0, 3], 0.0),
            (q[0, 2] - q[1, 3], q[1, 2] + q[0, 3], 1.0 - q[0, 0] - q[1, 1], 0.0),
            (0.0, 0.0, 0.0, 1.0)), dtype=np.float64)

    def quaternion_multiply(quaternion1, quaternion0):
        x0, y0, z0, w0 = quaternion0
        x1, y1, z1, w1 = quaternion1
        return np.array((
            x1 * w0 + y1 * z0 - z1 * y0 + w1 * x0,
            -x1 * z0 + y1 * w0 + z1 * x0 + w1 * y0,
            x1 * y0 - y1 * x0 + z1 * w0 + w1 * z0,
            -x1 * x0 - y1 * y0 - z1 * z0 + w1 * w0), dtype=np.float64)

    def quaternion_conjugate(quaternion):
        return np.array((-quaternion[0], -quaternion[1], -quaternion[2], quaternion[3]), dtype=np.float64)

    def quaternion_about_axis(angle, axis):
        q = np.zeros(4)
        ax = np.array(axis[:3], dtype=np.float64)
        qlen = np.linalg.norm(ax)
        if qlen > _EPS:
            q[:3] = ax * math.sin(angle / 2.0) / qlen
        q[3] = math.cos(angle / 2.0)
        return q

    def rotation_matrix(angle, direction, point=None):
        sina = math.sin(angle)
        cosa = math.cos(angle)
        d = np.array(direction, dtype=np.float64).flatten()[:3]
        d = d / math.sqrt(np.dot(d, d))
        R = np.array(((cosa, 0.0, 0.0), (0.0, cosa, 0.0), (0.0, 0.0, cosa)), dtype=np.float64)
        R += np.outer(d, d) * (1.0 - cosa)
        d = d * sina
        R += np.array(((0.0, -d[2], d[1]), (d[2], 0.0, -d[0]), (-d[1], d[0], 0.0)), dtype=np.float64)
        M = np.identity(4)
        M[:3, :3] = R
        return M

    for f in (quaternion_matrix, quaternion_multiply, quaternion_conjugate, quaternion_about_axis,
              rotation_matrix):
        setattr(m, f.__name__, f)
    return m


def load_prototype():
    np.math = math  # numpy.math was removed in NumPy 2; quaternion_helper.py:19,26,42 uses it
    rospy = types.ModuleType("rospy")
    rospy.loginfo = lambda *a, **k: None
    rospy.get_rostime = lambda: 0.0
    tf = types.ModuleType("tf")

    class _NoOp(object):
        def __init__(self, *a, **k):
            pass

        def sendTransform(self, *a, **k):
            pass

    tf.TransformBroadcaster = _NoOp
    tf.TransformListener = _NoOp
    tf.transformations = _tf_transformations()
    sys.modules["rospy"] = rospy
    sys.modules["tf"] = tf
    sys.modules["tf.transformations"] = tf.transformations
    for pkg, names in {
        "geometry_msgs": ["Point", "PointStamped", "Vector3", "Vector3Stamped", "Quaternion",
                          "PoseWithCovariance", "PoseWithCovarianceStamped", "Pose", "PoseStamped"],
        "sensor_msgs": ["Imu"],
        "apriltag_ros": ["AprilTagDetection", "AprilTagDetectionArray"],
        "std_msgs": ["Float64", "Header"],
    }.items():
        top = types.ModuleType(pkg)
        msg = types.ModuleType(pkg + ".msg")
        for n in names:
            setattr(msg, n, type(n, (_Bag,), {}))
        top.msg = msg
        sys.modules[pkg] = top
        sys.modules[pkg + ".msg"] = msg
    sys.path.insert(0, REF_TEST_DIR)
    import quaternion_helper as qh  # noqa: E402  (the reference's file)
    import rel_pose_EKF_test_class as proto  # noqa: E402  (the reference's file)
    # rel_pose_EKF_test_class.py:455 calls skew_symm() on a (3,1) column; NumPy < 1.25 silently
    # converted the size-1 arrays to scalars, NumPy 2 refuses.  Flatten the argument and hand it to
    # the reference's own skew_symm -- a compatibility shim like numpy.math, no arithmetic changed.
    proto.skew_symm = lambda v: qh.skew_symm(np.asarray(v, dtype=np.float64).flatten())
    return qh, proto


# ------------------------------------------------------------------------------------------------
# input generators (plain numpy; nothing here comes from the product)
# ------------------------------------------------------------------------------------------------
def rand_unit_quat(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def rand_spd(rng, n, scale=0.1):
    A = rng.normal(size=(n, n))
    return scale * (A @ A.T / n + 0.5 * np.eye(n))


def q_mul(a, b):  # Hamilton, xyzw
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def q_conj(a):
    return np.array([-a[0], -a[1], -a[2], a[3]])


def q_rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def q_exp(v):
    n = np.linalg.norm(v)
    if n < 1e-12:
        return np.array([v[0] / 2, v[1] / 2, v[2] / 2, 1.0])
    return np.concatenate([v / n * math.sin(n / 2), [math.cos(n / 2)]])


def make_sequence(rng, q_vc, r_v_cv, T, dT, meas_every, latency_steps, dropout=None, edge_window=None):
    """Noisy hover/translate trajectory above the tag; returns IMU [T,6], arrivals and tag poses.

    Truth recursion is the filter's own discrete model (explicit Euler), so the clean data are
    exactly consistent with it.  The camera looks down (q_vc of the prototype), the vehicle stays
    1.5-3 m above a 0.8 m tag so that the corner gate passes, except inside `edge_window` where the
    reported tag position is pushed towards the image edge to make the gate fail.
    """
    C_vc = q_rot(q_vc)
    g = np.array([0, 0, -9.8])
    # bounded sway: r(t) = r0 + [0.3 sin .7t, 0.2 (1 - cos .5t), 0.15 sin .3t]  (stays inside the image)
    r = np.array([0.2, -0.1, 2.5]); v = np.array([0.3 * 0.7, 0.0, 0.15 * 0.3]); q = q_exp(np.array([0.02, -0.01, 0.3]))
    ab = np.array([0.05, -0.03, 0.02]); wb = np.array([0.002, -0.001, 0.0015])
    truth = [(r.copy(), q.copy())]
    imu = np.zeros((T, 6))
    for k in range(T):
        t = k * dT
        acc = np.array([-0.3 * 0.49 * math.sin(0.7 * t), 0.2 * 0.25 * math.cos(0.5 * t),
                        -0.15 * 0.09 * math.sin(0.3 * t)])
        w = np.array([0.05 * math.sin(0.9 * t), 0.04 * math.cos(0.6 * t), 0.2 * math.sin(0.4 * t)])
        C = q_rot(q)
        imu[k, 0:3] = C.T @ (acc - g) + ab + rng.normal(scale=0.02, size=3)
        imu[k, 3:6] = w + wb + rng.normal(scale=0.005, size=3)
        r = r + dT * v
        v = v + dT * acc
        q = q_mul(q, q_exp(dT * w)); q /= np.linalg.norm(q)
        truth.append((r.copy(), q.copy()))
    steps, poses = [], []
    for k in range(3, T, meas_every):
        if dropout and dropout[0] <= k < dropout[1]:
            continue
        c = max(k + 1 - latency_steps, 0)
        rt, qt = truth[c]
        qn = q_mul(qt, q_exp(rng.normal(scale=0.01, size=3)))
        q_ct = q_conj(q_mul(qn, q_vc))
        r_c = C_vc.T @ (-q_rot(qt).T @ rt - r_v_cv) + rng.normal(scale=0.02, size=3)
        if edge_window and edge_window[0] <= k < edge_window[1]:
            r_c = r_c + np.array([5.0, 0.0, 0.0])  # far off-axis: projected corners leave the image
        steps.append(k)
        poses.append(np.concatenate([r_c, q_ct]))
    return imu, np.array(steps, dtype=np.int32), np.array(poses)


def run_prototype_sequence(proto, multirate, est_bias, imu, steps, poses, update_freq, meas_freq):
    ekf = proto.RelativePoseEKF(update_freq, meas_freq)
    ekf.multirate_EKF = multirate
    if not est_bias:  # re-derive what __init__ derives from est_bias (rel_pose_EKF_test_class.py:67-106)
        from scipy.linalg import block_diag
        ekf.est_bias = False
        ekf.num_states = 9
        ekf.cov_pert = np.zeros((9, 9))
        ekf.cov_init = np.diag(np.hstack((ekf.r_cov_init * np.ones(3), ekf.v_cov_init * np.ones(3),
                                          ekf.ang_cov_init * np.ones(3))))
        ekf.Q = block_diag(ekf.Q_a, ekf.Q_w)
    T = imu.shape[0]
    n = ekf.num_states
    xs = np.zeros((T, 16)); Ps = np.zeros((T, n, n)); upds = np.zeros(T, dtype=np.int32)
    active = np.zeros(T, dtype=np.int32)
    m = 0
    for k in range(T):
        if m < len(steps) and steps[m] == k:
            det = _Bag()
            det.pose.pose.pose.position.x, det.pose.pose.pose.position.y, det.pose.pose.pose.position.z = poses[m, 0:3]
            (det.pose.pose.pose.orientation.x, det.pose.pose.pose.orientation.y,
             det.pose.pose.pose.orientation.z, det.pose.pose.pose.orientation.w) = poses[m, 3:7]
            msg = _Bag(); msg.detections = [det]
            ekf.apriltag_msg = msg
            ekf.measurement_ready = True
            if not ekf.state_initialized:
                ekf.initialize_state(False)
            m += 1
        im = _Bag()
        im.linear_acceleration.x, im.linear_acceleration.y, im.linear_acceleration.z = imu[k, 0:3]
        im.angular_velocity.x, im.angular_velocity.y, im.angular_velocity.z = imu[k, 3:6]
        ekf.IMU_msg = im
        ekf.filter_update()
        if ekf.state_initialized:
            active[k] = 1
            xs[k] = np.concatenate([np.ravel(ekf.r_nom), np.ravel(ekf.v_nom), np.ravel(ekf.q_nom),
                                    np.ravel(ekf.ab_nom), np.ravel(ekf.wb_nom)])
            Ps[k] = ekf.cov_pert
            upds[k] = ekf.upds_since_correction
    return xs, Ps, upds, active


def main():
    qh, proto = load_prototype()
    rng = np.random.default_rng(20221207)
    out = {}

    # ---- 1. helpers ---------------------------------------------------------------------------
    vs = np.concatenate([rng.normal(scale=s, size=(8, 3)) for s in (1.0, 1e-3, 1e-8, 3e-11, 0.0)])
    vs[-1] = 0.0
    out["exp_in"] = vs
    out["exp_out_raw"] = np.array([qh.quaternion_exp(v) for v in vs])  # prototype does NOT normalise
    qs = np.array([rand_unit_quat(rng) for _ in range(24)])
    qs[:4, 3] = -np.abs(qs[:4, 3]) - 0.8          # force the w < -0.75 clip
    small = np.array([np.concatenate([rng.normal(scale=3e-11, size=3), [1.0]]) for _ in range(4)])
    qs = np.concatenate([qs, small])
    out["norm_in"] = qs
    out["norm_out"] = np.array([qh.quaternion_norm(q) for q in qs])
    out["log_in"] = out["norm_out"]
    out["log_out"] = np.array([qh.quaternion_log(q) for q in out["norm_out"]])
    out["skew_in"] = vs[:8]
    out["skew_out"] = np.array([qh.skew_symm(v) for v in vs[:8]])

    # ---- 2. step-level vectors ----------------------------------------------------------------
    for est_bias in (True, False):
        tagb = "b" if est_bias else "nb"
        ekf = proto.RelativePoseEKF(100.0, 10.0)
        if not est_bias:
            from scipy.linalg import block_diag
            ekf.est_bias = False
            ekf.num_states = 9
            ekf.Q = block_diag(ekf.Q_a, ekf.Q_w)
        n = ekf.num_states
        K = 12
        X = np.zeros((K, 16)); P = np.zeros((K, n, n)); U = np.zeros((K, 6))
        Xo = np.zeros((K, 16)); Po = np.zeros((K, n, n)); Ao = np.zeros((K, 3))
        TAG = np.zeros((K, 7)); Xc = np.zeros((K, 16)); Pc = np.zeros((K, n, n))
        for i in range(K):
            x = np.zeros(16)
            x[0:3] = rng.normal(scale=1.0, size=3) + np.array([0, 0, 2.0])
            x[3:6] = rng.normal(scale=0.5, size=3)
            x[6:10] = qh.quaternion_norm(rand_unit_quat(rng))
            if est_bias:
                x[10:13] = rng.normal(scale=0.05, size=3)
                x[13:16] = rng.normal(scale=0.005, size=3)
            u = np.concatenate([rng.normal(scale=1.0, size=3) + np.array([0, 0, 9.8]),
                                rng.normal(scale=0.3, size=3)])
            if i == 0:
                u[3:6] = x[13:16]  # exactly zero rate -> small-angle branch of F_theta_theta
            Pm = rand_spd(rng, n)
            xo, Pout, acc = ekf.prediction_step(x.reshape(16, 1), u.reshape(6, 1), Pm)
            X[i], P[i], U[i] = x, Pm, u
            Xo[i], Po[i], Ao[i] = xo.flatten(), Pout, acc.flatten()
            r_c = rng.normal(scale=0.5, size=3) + np.array([0, 0, 2.0])
            q_ct = rand_unit_quat(rng)
            if i < 6:  # measurement close to the prediction (small innovation)
                qt = q_mul(x[6:10], q_exp(rng.normal(scale=0.02, size=3)))
                q_ct = q_conj(q_mul(qt, ekf.q_vc))
            xc, Pcor = ekf.correction_step(x.reshape(16, 1), Pm, r_c.reshape(3, 1), q_ct)
            TAG[i] = np.concatenate([r_c, q_ct]); Xc[i] = xc.flatten(); Pc[i] = Pcor
        for k, v in dict(x=X, P=P, u=U, x_pred=Xo, P_pred=Po, accel=Ao, tag=TAG, x_corr=Xc, P_corr=Pc).items():
            out["step_%s_%s" % (tagb, k)] = v

    # ---- 3. initialize_state -------------------------------------------------------------------
    ekf = proto.RelativePoseEKF(100.0, 10.0)
    init_in = np.zeros((6, 7)); init_out = np.zeros((6, 16))
    for i in range(6):
        r_c = rng.normal(scale=0.5, size=3) + np.array([0, 0, 2.0]); q_ct = rand_unit_quat(rng)
        det = _Bag()
        det.pose.pose.pose.position.x, det.pose.pose.pose.position.y, det.pose.pose.pose.position.z = r_c
        (det.pose.pose.pose.orientation.x, det.pose.pose.pose.orientation.y,
         det.pose.pose.pose.orientation.z, det.pose.pose.pose.orientation.w) = q_ct
        msg = _Bag(); msg.detections = [det]
        ekf.apriltag_msg = msg
        ekf.initialize_state(True)
        init_in[i] = np.concatenate([r_c, q_ct])
        init_out[i] = np.concatenate([np.ravel(ekf.r_nom), np.ravel(ekf.v_nom), np.ravel(ekf.q_nom),
                                      np.ravel(ekf.ab_nom), np.ravel(ekf.wb_nom)])
    out["init_in"] = init_in; out["init_out"] = init_out

    # ---- 4. full filter_update sequences: 100 Hz update, measurement_freq 12.5 (upd_per_meas = 8),
    #         tag arrivals every 10 ticks, 5-tick capture latency in the multirate cases ------------
    ekf0 = proto.RelativePoseEKF(100.0, 12.5)
    T = 800
    for name, multirate, est_bias, lat in (("seq_mr", True, True, 5), ("seq_sr", False, True, 0),
                                           ("seq_mr_nb", True, False, 5)):
        imu, steps, poses = make_sequence(rng, ekf0.q_vc, ekf0.r_v_cv.flatten(), T, 0.01, 10, lat,
                                          dropout=(250, 330), edge_window=(400, 440))
        xs, Ps, upds, active = run_prototype_sequence(proto, multirate, est_bias, imu, steps, poses, 100.0, 12.5)
        out[name + "_imu"] = imu; out[name + "_tag_step"] = steps; out[name + "_tag_pose"] = poses
        out[name + "_x"] = xs; out[name + "_P"] = Ps[::10].copy(); out[name + "_P_last"] = Ps[-1]
        out[name + "_upds"] = upds; out[name + "_active"] = active
    out["seq_measurement_freq"] = np.array(12.5)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "%.1f KB" % (os.path.getsize(OUT) / 1024.0))


if __name__ == "__main__":
    main()
