"""GPU tier: delayed-measurement fusion (multirate_ekf, relative_pose_EKF.cpp:196-264) and per-filter
parameter sweeps (BASELINE config 5) through the C ABI against the dense oracle, which keeps the
reference's full x_hist / u_hist / P_hist vectors.  FP64 tolerance 1e-9 norm-relative."""
import os

import numpy as np
import pytest

import quadrotor_landing_b200 as q
from oracle import ekf_oracle as orc
from quadrotor_landing_b200 import scenario
from streams_np import noisy_streams, norm_rel, rotors_params
from test_multirate_host import delayed_scenario, sweep_values

pytestmark = pytest.mark.gpu
TOL = 1e-9
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "prototype_vectors.npz"))


def compare(b, ob):
    assert norm_rel(b.state(), ob.state()) < TOL
    assert norm_rel(b.cov(), ob.cov()) < TOL
    fl = ob.flags()
    assert np.array_equal(b.flags(), fl)           # incl. x_hist.size()
    init = fl[0] != 0
    if init.any():
        assert norm_rel(b.aux()[:, init], ob.aux()[:, init]) < TOL


@pytest.mark.parametrize("est_bias,direct,dynamic", [(1, 1, 0), (1, 1, 1), (1, 0, 0), (0, 1, 1), (0, 0, 0)])
def test_multirate_replay_matches_oracle(est_bias, direct, dynamic):
    """300 filters (a ragged last CTA) with private dropouts, so that lanes of one CTA correct at different
    ticks, catch their checkpoints up inside each other's correction events and hit the ring-full path."""
    p = rotors_params(q.default_params(), est_bias=est_bias, direct=direct, multirate=True, dynamic_delay=bool(dynamic))
    scn = delayed_scenario(p, 0.042 if dynamic else 0.030)
    N, T = 300, 3000
    st = noisy_streams(scn, N, seed=23, T=T, dropout=(900, 1250), random_dropout_ticks=180)
    ob = orc.Batch(orc.params_from(p), N)
    b = q.BatchEKF(p, N)
    for k0, n in ((0, 1), (1, 700), (701, 3), (704, 1297), (2001, 999)):
        ob.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        b.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        compare(b, ob)
    assert ob.counts()[1] > 150 * N
    # lazily evaluated history: about one prediction per tick, where the eager reference runs ~ two
    npred, ncorr = b.step_counts()
    assert ncorr == ob.counts()[1]
    assert npred < 0.75 * ob.counts()[0]
    b.close()


def test_multirate_long_delays_no_frequency_gate():
    p = rotors_params(q.default_params(), multirate=True, dynamic_delay=True)
    p.limit_measurement_freq = 0
    p.dyn_measurement_delay_offset = 0.06
    scn = delayed_scenario(p, 0.150, seconds=8.0)
    N, T = 64, 1500
    st = noisy_streams(scn, N, seed=5, T=T, dropout=(400, 520), random_dropout_ticks=100)
    ob = orc.Batch(orc.params_from(p), N)
    b = q.BatchEKF(p, N)
    for k0, n in ((0, 333), (333, 1167)):
        ob.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        b.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        compare(b, ob)
    b.close()


@pytest.mark.parametrize("name,est_bias", [("seq_mr", 1), ("seq_mr_nb", 0)])
def test_multirate_per_tick_interface_matches_prototype_golden(name, est_bias):
    """The reference prototype's own multirate sequences (fixed 50 ms delay, rel_pose_EKF_test_class.py:59)
    through the N = 1 per-tick interface: set_imu / set_tag / filter_update, as
    relative_pose_EKF_node.cpp:144-182 calls them."""
    ekf = q.RelativePoseEKF()
    ekf.update_freq, ekf.measurement_freq = 100.0, float(G["seq_measurement_freq"])
    ekf.limit_measurement_freq = ekf.corner_margin_enbl = ekf.direct_orien_method = 1
    ekf.multirate_ekf, ekf.dynamic_meas_delay, ekf.est_bias = 1, 0, est_bias
    ekf.measurement_delay = 0.050
    ekf.initialize_params()
    imu, steps, poses = G[name + "_imu"], G[name + "_tag_step"], G[name + "_tag_pose"]
    arrivals = {int(s): m for m, s in enumerate(steps)}
    worst = 0.0
    for k in range(imu.shape[0]):
        if k in arrivals:
            m = arrivals[k]
            ekf.set_tag(poses[m, 0:3], poses[m, 3:7], 0.0)
        ekf.set_imu(imu[k, 0:3], imu[k, 3:6])
        ekf.filter_update(k * 0.01)
        if not G[name + "_active"][k]:
            continue
        assert ekf.upds_since_correction == G[name + "_upds"][k]
        x = np.concatenate([ekf.r_nom, ekf.v_nom, ekf.q_nom, ekf.ab_nom, ekf.wb_nom])
        worst = max(worst, norm_rel(x, G[name + "_x"][k]))
        if k % 10 == 0:
            worst = max(worst, norm_rel(ekf.cov_pert, G[name + "_P"][k // 10]))
    assert worst < TOL


@pytest.mark.parametrize("multirate,dynamic,direct", [(0, 0, 1), (1, 0, 1), (1, 1, 0)])
def test_per_filter_parameter_sweep_matches_oracle(multirate, dynamic, direct):
    p = rotors_params(q.default_params(), multirate=bool(multirate), dynamic_delay=bool(dynamic), direct=direct)
    scn = delayed_scenario(p, 0.035 if multirate else 0.0, seconds=10.0)
    N, T = 257, 2000
    st = noisy_streams(scn, N, seed=77, T=T, dropout=(700, 900))
    ob = orc.Batch(orc.params_from(p), N)
    b = q.BatchEKF(p, N)
    for field, v in sweep_values(np.random.default_rng(9), p, N, bool(multirate)).items():
        ob.set_filter_params(field, v)
        b.set_filter_params(field, v)
    for k0, n in ((0, 801), (801, 1199)):
        ob.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        b.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        compare(b, ob)
    b.close()


def test_multirate_monte_carlo_matches_explicit_path_and_statistics():
    """In-kernel noise + delayed fusion: bit-identical to the explicit-stream kernel fed the dumped
    realisation, and the on-chip statistics (sampled from the materialised head) match the oracle's states."""
    from oracle import noise_np
    p = rotors_params(q.default_params(), multirate=True)
    scn = delayed_scenario(p, 0.030, seconds=8.0)
    noise = q.default_noise()
    noise.first_global_id = 7000
    noise.dropout_k0, noise.dropout_k1 = 600, 800
    noise.rand_dropout_len, noise.rand_dropout_lo, noise.rand_dropout_hi = 150, 100, 1200
    N, stride = 300, 400
    nb = scn.T // stride
    b = q.BatchEKF(p, N)
    b.stats_configure(nb, stride)
    b.run_monte_carlo(scn, noise, 0, 777)
    b.run_monte_carlo(scn, noise, 777, scn.T - 777)
    st = b.synthesize_streams(scn, noise, 0, N)
    b2 = q.BatchEKF(p, N)
    b2.run(0, scn.T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    assert np.array_equal(b.state(), b2.state()) and np.array_equal(b.cov(), b2.cov())
    ob = orc.Batch(orc.params_from(p), N)
    ref = np.zeros((nb, 20))
    for k in range(nb):
        ob.run(k * stride, stride, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        e, nees = noise_np.error_stats(ob.state(), ob.cov(), scn.truth[(k + 1) * stride], st["bias"], 15)
        ref[k, 0:15] = (e ** 2).sum(axis=1)
        ref[k, 15], ref[k, 16] = nees.sum(), N
        ref[k, 19] = (e[0:3] ** 2).sum()
    compare(b, ob)
    stats = b.stats()
    assert np.array_equal(stats[:, 16], ref[:, 16])
    assert norm_rel(stats[:, 0:16], ref[:, 0:16]) < TOL and norm_rel(stats[:, 19], ref[:, 19]) < TOL
    b.close(); b2.close()


def _mc_noise(first=0):
    n = q.default_noise()
    n.seed = 4711
    n.first_global_id = first
    n.dropout_k0, n.dropout_k1 = 500, 640
    n.rand_dropout_len, n.rand_dropout_lo, n.rand_dropout_hi = 160, 100, 1300
    return n


@pytest.mark.parametrize("dynamic,sweep,precision,est_bias,direct", [(0, 0, 64, 1, 1), (1, 0, 64, 1, 1), (1, 1, 64, 1, 1),
                                                                     (1, 0, 32, 1, 1), (1, 0, 64, 0, 0), (0, 1, 64, 1, 0)])
def test_resynthesised_history_equals_ring_history(monkeypatch, dynamic, sweep, precision, est_bias, direct):
    """Monte-Carlo launches with delayed fusion re-synthesise the IMU inputs of the history entries (no ring, no light
    ticks: run_filter_mrs); QEKF_NO_LAZY_MR=1 keeps the ring-based loop.  Same arithmetic on the same inputs: every
    accessor bit-identical, over chunked launches whose cuts fall between corrections."""
    p = rotors_params(q.default_params(), multirate=True, dynamic_delay=bool(dynamic), est_bias=est_bias, direct=direct)
    scn = delayed_scenario(p, 0.042 if dynamic else 0.030, seconds=8.0)
    N = 1500
    out = []
    for env in ({"QEKF_NO_LAZY_MR": "1"}, {}):
        monkeypatch.delenv("QEKF_NO_LAZY_MR", raising=False)
        for k_, v_ in env.items():
            monkeypatch.setenv(k_, v_)
        b = q.BatchEKF(p, N, precision=precision)
        if sweep:
            for field, v in sweep_values(np.random.default_rng(3), p, N, True).items():
                b.set_filter_params(field, v)
        b.stats_configure(16, 100)
        for k0, n in ((0, 611), (611, 2), (613, 987)):
            b.run_monte_carlo(scn, _mc_noise(), k0, n)
        out.append(([b.state(), b.cov(), b.aux(), b.flags()], b.stats(), b.step_counts()))
        b.close()
    (ring, ring_stats, ring_counts), (lazy, lazy_stats, lazy_counts) = out
    for x, y in zip(ring, lazy):
        assert np.array_equal(x, y, equal_nan=True)
    assert lazy_counts[1] == ring_counts[1] and lazy_counts[1] > 100 * N
    assert np.array_equal(lazy_stats[:, 16:19], ring_stats[:, 16:19])
    assert np.allclose(lazy_stats, ring_stats, rtol=1e-11, atol=0)


def test_monte_carlo_history_hands_over_to_the_other_entry_points():
    """A Monte-Carlo launch leaves a history the ring-based paths can continue from (its epilogue writes the pending
    entries' inputs to the ring), and a handle whose history came from explicit streams is not re-synthesised: explicit
    run -> Monte-Carlo launch -> explicit run, against the oracle replaying the same inputs end to end."""
    p = rotors_params(q.default_params(), multirate=True, dynamic_delay=True)
    scn = delayed_scenario(p, 0.042, seconds=8.0)
    N, T = 96, 1500
    noise = _mc_noise(first=1000)
    b = q.BatchEKF(p, N)
    st = b.synthesize_streams(scn, noise, 0, N)       # the realisation as explicit streams: what the oracle replays
    args = (st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    ob = orc.Batch(orc.params_from(p), N)
    ob.run(0, T, *args)
    b.run(0, 403, *args)                              # explicit streams (history entries live in the ring)
    b.run_monte_carlo(scn, noise, 403, 300)           # must read them from the ring
    b.run_monte_carlo(scn, noise, 703, 300)           # (still the ring: entries from before the explicit run may be pending)
    b.run(1003, T - 1003, *args)                      # continues from the ring the Monte-Carlo launch left behind
    compare(b, ob)
    b.close()


@pytest.mark.parametrize("multirate", [0, 1])
def test_monte_carlo_without_frequency_gate_equals_explicit_replay(multirate):
    """limit_measurement_freq = 0: every valid detection is fused at its arrival tick (cpp:147), so the event-driven
    loops stop at every arrival instead of every upd_per_meas ticks.  The Monte-Carlo launch (delayed fusion: no ring)
    must equal, bit for bit, the explicit-stream replay of its own dumped realisation, and the oracle to 1e-9."""
    p = rotors_params(q.default_params(), multirate=bool(multirate), dynamic_delay=bool(multirate))
    p.limit_measurement_freq = 0
    p.corner_margin_enbl = 0
    scn = delayed_scenario(p, 0.060 if multirate else 0.0, seconds=6.0)
    N = 200
    noise = _mc_noise(first=77)
    b = q.BatchEKF(p, N)
    b.run_monte_carlo(scn, noise, 0, 500)
    b.run_monte_carlo(scn, noise, 500, scn.T - 500)
    st = b.synthesize_streams(scn, noise, 0, N)
    args = (st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
    b2 = q.BatchEKF(p, N)
    b2.run(0, scn.T, *args)
    for x, y in ((b.state(), b2.state()), (b.cov(), b2.cov()), (b.aux(), b2.aux()), (b.flags(), b2.flags())):
        assert np.array_equal(x, y, equal_nan=True)
    ob = orc.Batch(orc.params_from(p), N)
    ob.run(0, scn.T, *args)
    compare(b, ob)
    assert ob.counts()[1] > 100 * N
    b.close(); b2.close()
