"""CPU ORACLE for the product's in-kernel noise generator (TEST INFRASTRUCTURE, not product code).

An independent numpy restatement of the published Philox4x32-10 algorithm (Salmon et al., "Parallel
random numbers: as easy as 1, 2, 3", SC'11; constants 0xD2511F53 / 0xCD9E8D57, key increments
0x9E3779B9 / 0xBB67AE85), of Box-Muller, and of the noise model documented in include/qekf.h
(qekf_noise_spec).  The product's device generator (csrc/ekf_synth.cuh) is checked against this.
Normals are formed in float32 on the device with the device's logf/sincospif, so agreement is to a few
float ulps, not bit-exact; parity runs therefore replay the streams the device dumps.
"""
import numpy as np

STREAM_IMU, STREAM_TAG, STREAM_BIAS, STREAM_DROPOUT = 0, 2, 4, 6
M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0); k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def uniforms21x6(w):
    """Six 21-bit integers from the 128 bits of one Philox block (w: four uint32 arrays)."""
    w0, w1, w2, w3 = (x.astype(np.uint32) for x in w)
    u32 = np.uint32
    return [w0 & u32(0x1FFFFF),
            (w0 >> u32(21)) | ((w1 & u32(0x3FF)) << u32(11)),
            (w1 >> u32(10)) & u32(0x1FFFFF),
            (w1 >> u32(31)) | ((w2 & u32(0xFFFFF)) << u32(1)),
            (w2 >> u32(20)) | ((w3 & u32(0x1FF)) << u32(12)),
            (w3 >> u32(9)) & u32(0x1FFFFF)]


def box_muller(a, b):
    u1 = (a.astype(np.float64) + 0.5) / 2097152.0
    u2 = (b.astype(np.float64) + 0.5) / 2097152.0
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)


def normals6(seed, gid, stream, index):
    """gid, index broadcastable integer arrays -> array [..., 6] of standard normals (one Philox block each)."""
    gid = np.asarray(gid, dtype=np.uint64)
    g0, g1 = (gid & MASK).astype(np.uint32), (gid >> np.uint64(32)).astype(np.uint32)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    idx = np.asarray(index, dtype=np.uint32)
    u = uniforms21x6(philox4x32_10(idx, np.uint32(stream), g0, g1, k0, k1))
    z0, z1 = box_muller(u[0], u[1])
    z2, z3 = box_muller(u[2], u[3])
    z4, z5 = box_muller(u[4], u[5])
    return np.stack([z0, z1, z2, z3, z4, z5], axis=-1)


def q_mul(a, b):
    ax, ay, az, aw = np.moveaxis(a, -1, 0)
    bx, by, bz, bw = np.moveaxis(b, -1, 0)
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def q_rot(q):
    """[..., 4] unit quaternions (x, y, z, w) -> [..., 3, 3] rotation matrices."""
    x, y, z, w = np.moveaxis(q, -1, 0)
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def bundle_in_image(tag_pose_clean, params):
    """Detection front-end geometry (include/qekf.h, qekf_noise_spec.edge_loss): [M] bool, true when at least one
    tag of the bundle projects with all four corners strictly inside the image (clean pose, pinhole camera_K)."""
    K = np.array(list(params.camera_K)).reshape(3, 3)
    R = q_rot(tag_pose_clean[:, 3:7])
    t = tag_pose_clean[:, 0:3]
    ok = np.zeros(tag_pose_clean.shape[0], dtype=bool)
    for i in range(params.n_tags):
        hw = params.tag_widths[i] / 2
        c0 = np.array([params.tag_positions[3 * i], params.tag_positions[3 * i + 1]])
        inside = np.ones_like(ok)
        for sx, sy in ((1, 1), (-1, 1), (-1, -1), (1, -1)):
            corner = np.array([sx * hw + c0[0], sy * hw + c0[1], 0.0])
            pc = R @ corner + t
            u = K[0, 0] * pc[:, 0] / pc[:, 2] + K[0, 1] * pc[:, 1] / pc[:, 2] + K[0, 2]
            v = K[1, 0] * pc[:, 0] / pc[:, 2] + K[1, 1] * pc[:, 1] / pc[:, 2] + K[1, 2]
            inside &= (pc[:, 2] > 0) & (u > 0) & (u < params.camera_width) & (v > 0) & (v < params.camera_height)
        ok |= inside
    return ok


def synthesize(noise, imu_clean, tag_step, tag_pose_clean, gids, params=None):
    """Explicit streams for the given global filter ids, layout of qekf_streams:
    imu [T][6][n], tag_pose [M][7][n], tag_valid [M][n], bias [6][n].  `noise` has the qekf_noise_spec fields;
    `params` (camera + bundle geometry) is needed when the detection front-end is on."""
    gids = np.asarray(gids, dtype=np.int64)
    n = gids.shape[0]
    T, M = imu_clean.shape[0], tag_step.shape[0]
    seed = int(noise.seed)
    zb = normals6(seed, gids, STREAM_BIAS, 0)                                  # [n, 6]
    bias = np.concatenate([noise.sigma_bias_accel * zb[:, 0:3], noise.sigma_bias_gyro * zb[:, 3:6]], axis=1).T
    z = normals6(seed, gids[None, :], STREAM_IMU, np.arange(T)[:, None])       # [T, n, 6]
    sig = np.array([noise.sigma_accel] * 3 + [noise.sigma_gyro] * 3)
    imu = imu_clean[:, :, None] + bias[None] + (z * sig).transpose(0, 2, 1)
    zt = normals6(seed, gids[None, :], STREAM_TAG, np.arange(M)[:, None])      # [M, n, 6]
    pose = np.repeat(tag_pose_clean[:, :, None], n, axis=2)
    sp = np.full(M, noise.sigma_tag_pos)
    sth = np.full(M, noise.sigma_tag_ang)
    if getattr(noise, "range_ref", 0.0) > 0:
        rel = np.linalg.norm(tag_pose_clean[:, 0:3], axis=1) / noise.range_ref
        if noise.range_exp_pos != 0:
            sp = sp * rel ** noise.range_exp_pos
        if noise.range_exp_ang != 0:
            sth = sth * rel ** noise.range_exp_ang
    pose[:, 0:3] += sp[:, None, None] * zt[:, :, 0:3].transpose(0, 2, 1)
    v = sth[:, None, None] * zt[:, :, 3:6]
    nn = np.linalg.norm(v, axis=-1)
    f = np.where(nn < 1e-10, 0.5, np.sin(0.5 * nn) / np.where(nn < 1e-10, 1.0, nn))
    dq = np.concatenate([v * f[..., None], np.cos(0.5 * nn)[..., None]], axis=-1)       # [M, n, 4]
    q = q_mul(dq, np.broadcast_to(tag_pose_clean[:, None, 3:7], dq.shape))
    pose[:, 3:7] = q.transpose(0, 2, 1)
    valid = np.ones((M, n), dtype=np.uint8)
    st = tag_step.astype(np.int64)
    valid[(st >= noise.dropout_k0) & (st < noise.dropout_k1)] = 0
    if noise.rand_dropout_len > 0 and noise.rand_dropout_hi > noise.rand_dropout_lo:
        g = gids.astype(np.uint64)
        w = philox4x32_10(np.uint32(0), np.uint32(STREAM_DROPOUT), (g & MASK).astype(np.uint32),
                          (g >> np.uint64(32)).astype(np.uint32), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        span = np.uint64(noise.rand_dropout_hi - noise.rand_dropout_lo)
        start = noise.rand_dropout_lo + ((w[0].astype(np.uint64) * span) >> np.uint64(32)).astype(np.int64)
        valid[(st[:, None] >= start[None]) & (st[:, None] < start[None] + noise.rand_dropout_len)] = 0
    if getattr(noise, "edge_loss", 0):
        valid[~bundle_in_image(tag_pose_clean, params)] = 0
    return dict(imu=imu, tag_pose=pose, tag_valid=valid, bias=bias)


def error_stats(x, P, truth_row, bias, n_states):
    """RMSE/NEES ingredients from final states, restating the definition in include/qekf.h:
    e = (r_t - r, v_t - v, log(conj(q) (x) q_t), b_a - ab, b_w - wb), NEES = e^T P^-1 e."""
    N = x.shape[1]
    e = np.zeros((n_states, N))
    e[0:3] = truth_row[0:3, None] - x[0:3]
    e[3:6] = truth_row[3:6, None] - x[3:6]
    qc = x[6:10].T * np.array([-1, -1, -1, 1.0])
    dq = q_mul(qc, np.broadcast_to(truth_row[6:10], qc.shape))
    dq /= np.linalg.norm(dq, axis=1, keepdims=True)
    dq[dq[:, 3] < -0.75] *= -1
    vn = np.linalg.norm(dq[:, 0:3], axis=1)
    f = np.where(vn < 1e-10, 2.0 / dq[:, 3], 2 * np.arctan2(vn, dq[:, 3]) / np.where(vn < 1e-10, 1.0, vn))
    e[6:9] = (dq[:, 0:3] * f[:, None]).T
    if n_states == 15:
        e[9:12] = bias[0:3] - x[10:13]
        e[12:15] = bias[3:6] - x[13:16]
    nees = np.array([e[:, i] @ np.linalg.solve(P[:, :, i], e[:, i]) for i in range(N)])
    return e, nees
