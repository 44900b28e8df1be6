"""Test-side helpers: explicit per-filter noisy streams built with numpy on top of the product's clean
scenario, and the parameter presets used across the parity tests."""
import numpy as np


def q_mul(a, b):
    """Hamilton product, xyzw, broadcasting over trailing axes: a, b of shape (4, ...)."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def q_exp(v):
    """v of shape (3, ...) -> unit quaternion (4, ...)."""
    n = np.sqrt((v * v).sum(axis=0))
    half = 0.5 * n
    f = np.where(n < 1e-12, 0.5, np.sin(half) / np.where(n < 1e-12, 1.0, n))
    return np.concatenate([v * f, np.cos(half)[None]], axis=0)


def noisy_streams(scn, N, seed, sigma_a=0.02, sigma_w=0.007, sigma_ba=0.05, sigma_bw=0.002,
                  sigma_p=0.02, sigma_th=0.01, dropout=None, random_dropout_ticks=0, T=None):
    """Per-filter explicit streams in the C-ABI layout: imu [T][6][N], tag_pose [M][7][N], tag_valid [M][N].

    Noise model (camera-frame measurement noise, matching the filter's R_k = N R N^T structure):
    imu += bias_i + N(0, sigma); r_c += N(0, sigma_p); q_ct <- exp(N(0, sigma_th)) (x) q_ct.
    """
    rng = np.random.default_rng(seed)
    T = scn.T if T is None else T
    sel = scn.tag_step < T
    steps = scn.tag_step[sel]
    M = len(steps)
    imu = np.repeat(scn.imu_clean[:T, :, None], N, axis=2)
    bias = np.concatenate([rng.normal(scale=sigma_ba, size=(3, N)), rng.normal(scale=sigma_bw, size=(3, N))])
    imu += bias[None]
    imu[:, 0:3] += rng.normal(scale=sigma_a, size=(T, 3, N))
    imu[:, 3:6] += rng.normal(scale=sigma_w, size=(T, 3, N))
    pose = np.repeat(scn.tag_pose_clean[sel][:, :, None], N, axis=2)
    pose[:, 0:3] += rng.normal(scale=sigma_p, size=(M, 3, N))
    dq = q_exp(rng.normal(scale=sigma_th, size=(M, 3, N)).transpose(1, 0, 2))       # (4, M, N)
    q = q_mul(dq, pose[:, 3:7].transpose(1, 0, 2))
    pose[:, 3:7] = q.transpose(1, 0, 2)
    valid = np.ones((M, N), dtype=np.uint8)
    if dropout is not None:
        valid[(steps >= dropout[0]) & (steps < dropout[1])] = 0
    if random_dropout_ticks > 0:
        start = rng.integers(0, max(T - random_dropout_ticks, 1), size=N)
        for i in range(N):
            valid[(steps >= start[i]) & (steps < start[i] + random_dropout_ticks), i] = 0
    return dict(imu=np.ascontiguousarray(imu), tag_step=steps.astype(np.int32), tag_pose=np.ascontiguousarray(pose),
                tag_stamp=scn.tag_stamp[sel].copy(), tag_valid=valid, bias=bias)


def rotors_params(P, update_freq=200.0, measurement_freq=30.0, multirate=False, direct=True, est_bias=True,
                  dynamic_delay=False):
    """quad_state_estimation/config/relative_pose_EKF_rotors.yaml:3-45 with the benchmark rates
    (200 Hz update, 30 Hz tag).  P is a params struct (oracle or product)."""
    P.update_freq = update_freq
    P.measurement_freq = measurement_freq
    P.measurement_delay = 0.030
    P.measurement_delay_max = 0.200
    P.dyn_measurement_delay_offset = 0.005
    for i in range(3):
        P.Q_a[i] = 0.0005; P.Q_w[i] = 0.00005; P.Q_ab[i] = 5.0e-5; P.Q_wb[i] = 5.0e-6
    P.R_r[0], P.R_r[1], P.R_r[2] = 0.015, 0.015, 0.020
    P.R_ang[0], P.R_ang[1], P.R_ang[2] = 0.0015, 0.0015, 0.04
    P.limit_measurement_freq = 1
    P.corner_margin_enbl = 1
    P.est_bias = int(est_bias)
    P.direct_orien_method = int(direct)
    P.multirate_ekf = int(multirate)
    P.dynamic_meas_delay = int(dynamic_delay)
    return P


def norm_rel(a, b):
    """max|a-b| / max|b| (norm-relative; element-wise relative error is meaningless for near-zero
    cross-covariances)."""
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
