"""CPU tier: delayed-measurement fusion (multirate_ekf, relative_pose_EKF.cpp:196-264) and per-filter
parameter overrides, product code instantiated for the host (tests/host_core) against the dense oracle,
which keeps the reference's full x_hist / u_hist / P_hist vectors.  The product evaluates the history
lazily (lagged checkpoint + IMU ring); the results must be the same to rounding."""
import numpy as np
import pytest

import host_core as hc
import quadrotor_landing_b200 as q
from oracle import ekf_oracle as orc
from quadrotor_landing_b200 import scenario
from streams_np import noisy_streams, norm_rel, rotors_params

TOL = 1e-9


def delayed_scenario(p, latency_s, seconds=15.0):
    spec = scenario.default_spec()
    spec.duration_s, spec.hover_s = seconds, 3.0
    spec.tag_latency_s = latency_s
    return scenario.generate(p, spec)


def compare(hb, ob, check_hist=True):
    assert norm_rel(hb.state(), ob.state()) < TOL
    assert norm_rel(hb.cov(), ob.cov()) < TOL
    fl = ob.flags()
    assert np.array_equal(hb.upds, fl[4])
    assert np.array_equal(hb.flags & 1, fl[0]) and np.array_equal((hb.flags >> 1) & 1, fl[1])
    assert np.array_equal((hb.flags >> 2) & 1, fl[2])
    if check_hist:
        assert np.array_equal(hb.history_length(), fl[5])
    init = fl[0] != 0
    if init.any():
        assert norm_rel(hb.aux[:, init], ob.aux()[:, init]) < TOL


@pytest.mark.parametrize("est_bias,direct,dynamic", [(1, 1, 0), (1, 1, 1), (1, 0, 0), (0, 1, 1), (0, 0, 0)])
def test_multirate_replay_matches_oracle(est_bias, direct, dynamic):
    """Fixed (30 ms = 6 ticks) and dynamic (stamp-derived) measurement delay; chunked replay with chunk
    borders that cut through pending measurements and through the history."""
    p = rotors_params(q.default_params(), est_bias=est_bias, direct=direct, multirate=True, dynamic_delay=bool(dynamic))
    scn = delayed_scenario(p, 0.042 if dynamic else 0.030)
    N, T = 5, 3000
    st = noisy_streams(scn, N, seed=23, T=T, dropout=(900, 1250), random_dropout_ticks=180)
    ob = orc.Batch(orc.params_from(p), N)
    hb = hc.HostBatch(p, N)
    for k0, n in ((0, 1), (1, 700), (701, 3), (704, 1297), (2001, 999)):
        ob.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        hb.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        compare(hb, ob)
    assert ob.counts()[1] > 150 * N
    err = ob.state()[0:3] - scn.truth[T][0:3, None]
    assert np.max(np.abs(err)) < 0.1


def test_multirate_without_frequency_gate_and_with_long_delays():
    """limit_measurement_freq = 0 (a correction on every arrival), delays up to the 0.2 s cap (40 ticks),
    capture-to-arrival latency 150 ms, tick by tick for the first stretch (the per-tick interface path)."""
    p = rotors_params(q.default_params(), multirate=True, dynamic_delay=True)
    p.limit_measurement_freq = 0
    p.dyn_measurement_delay_offset = 0.06       # 150 ms + 60 ms > delay_max: the clamp is exercised
    scn = delayed_scenario(p, 0.150, seconds=8.0)
    N, T = 3, 1500
    st = noisy_streams(scn, N, seed=5, T=T, dropout=(400, 520))
    ob = orc.Batch(orc.params_from(p), N)
    hb = hc.HostBatch(p, N)
    for k in range(0, 120):
        ob.run(k, 1, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        hb.run(k, 1, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        compare(hb, ob)
    ob.run(120, T - 120, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    hb.run(120, T - 120, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    compare(hb, ob)
    assert np.max(ob.aux()[10]) == pytest.approx(p.measurement_delay_max)


def sweep_values(rng, p, N, with_delay):
    """BASELINE config 5: per-filter Q and R within x[0.1, 10] of the preset (log-uniform), camera extrinsic
    +-2 cm / +-1 deg, measurement delay 20..60 ms."""
    base_q = np.array(list(p.Q_a) + list(p.Q_w) + list(p.Q_ab) + list(p.Q_wb))
    base_r = np.array(list(p.R_r) + list(p.R_ang))
    Q = base_q[:, None] * 10 ** rng.uniform(-1, 1, size=(12, N))
    R = base_r[:, None] * 10 ** rng.uniform(-1, 1, size=(6, N))
    rv = np.array(list(p.r_v_cv))[:, None] + rng.uniform(-0.02, 0.02, size=(3, N))
    qv = np.zeros((4, N))
    for i in range(N):
        dq = orc.quat_exp(rng.normal(scale=np.deg2rad(1.0), size=3))
        qv[:, i] = orc.quat_mul(np.array(list(p.q_vc)), dq) * 1.0003     # deliberately not unit length
    out = {orc.PF_Q: Q, orc.PF_R: R, orc.PF_R_V_CV: rv, orc.PF_Q_VC: qv}
    if with_delay:
        out[orc.PF_DELAY] = np.stack([rng.uniform(0.020, 0.060, size=N), rng.uniform(0.0, 0.01, size=N)])
    return out


@pytest.mark.parametrize("multirate,dynamic,direct", [(0, 0, 1), (1, 0, 1), (1, 1, 0)])
def test_per_filter_parameter_sweep_matches_oracle(multirate, dynamic, direct):
    p = rotors_params(q.default_params(), multirate=bool(multirate), dynamic_delay=bool(dynamic), direct=direct)
    scn = delayed_scenario(p, 0.035 if multirate else 0.0, seconds=10.0)
    N, T = 6, 2000
    st = noisy_streams(scn, N, seed=77, T=T, dropout=(700, 900))
    ob = orc.Batch(orc.params_from(p), N)
    hb = hc.HostBatch(p, N)
    for field, v in sweep_values(np.random.default_rng(9), p, N, bool(multirate)).items():
        ob.set_filter_params(field, v)
        hb.set_filter_params(field, v)
    for k0, n in ((0, 801), (801, 1199)):
        ob.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        hb.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        compare(hb, ob)
    # the sweep really produced different filters
    assert norm_rel(ob.cov()[:, :, 0], ob.cov()[:, :, 1]) > 1e-2


@pytest.mark.parametrize("dynamic,est_bias,direct,sweep", [(0, 1, 1, 0), (1, 1, 1, 0), (1, 0, 0, 0), (1, 1, 1, 1)])
def test_monte_carlo_delayed_fusion_without_ring_matches_ring_and_oracle(dynamic, est_bias, direct, sweep):
    """The Monte-Carlo delayed-fusion loop that re-synthesises the history inputs and jumps from decision point to
    decision point (run_filter_mrs, the product code instantiated for the host) against (i) the ring-based loop, bit
    for bit, including the on-chip statistics, over chunked launches, and (ii) the dense oracle -- which keeps the
    reference's x_hist / u_hist / P_hist vectors (relative_pose_EKF.cpp:196-264) -- replaying the dumped realisation."""
    from quadrotor_landing_b200 import _native as nat
    p = rotors_params(nat.default_params(), multirate=True, dynamic_delay=bool(dynamic), est_bias=est_bias, direct=direct)
    scn = delayed_scenario(p, 0.042 if dynamic else 0.030, seconds=6.0)
    noise = nat.default_noise()
    noise.seed = 99
    noise.first_global_id = 5_000_000_000
    noise.dropout_k0, noise.dropout_k1 = 400, 520
    noise.rand_dropout_len, noise.rand_dropout_lo, noise.rand_dropout_hi = 120, 100, 900
    N, stride = 6, 200
    nb = scn.T // stride
    runs = []
    for lazy in (False, True):
        hb = hc.HostBatch(p, N)
        if sweep:
            for field, v in sweep_values(np.random.default_rng(4), p, N, True).items():
                hb.set_filter_params(field, v)
        acc = np.zeros((32, nb, 20))
        for k0, n in ((0, 431), (431, 3), (434, scn.T - 434)):
            hb.run_mc(scn, noise, k0, n, acc, stride, lazy=lazy)
        runs.append((hb, acc.sum(axis=0)))
    (ring, ring_stats), (lazy_b, lazy_stats) = runs
    assert np.array_equal(ring.state(), lazy_b.state()) and np.array_equal(ring.cov(), lazy_b.cov())
    assert np.array_equal(ring.aux, lazy_b.aux) and np.array_equal(ring.flags, lazy_b.flags) and np.array_equal(ring.upds, lazy_b.upds)
    assert np.array_equal(ring.history_length(), lazy_b.history_length())
    assert np.array_equal(ring_stats[:, 16:19], lazy_stats[:, 16:19]) and np.allclose(ring_stats, lazy_stats, rtol=1e-12, atol=0)
    assert ring_stats[:, 16].sum() > 0
    st = hc.synthesize(scn, noise, 0, N, params=p)
    ob = orc.Batch(orc.params_from(p), N)
    if sweep:
        for field, v in sweep_values(np.random.default_rng(4), p, N, True).items():
            ob.set_filter_params(field, v)
    ob.run(0, scn.T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    compare(lazy_b, ob)
    assert ob.counts()[1] > 50 * N
