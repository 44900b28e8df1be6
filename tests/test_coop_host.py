"""The cooperative (three lanes per filter) step functions of ekf_coop.cuh, run on the host lane by lane, against the
thread-per-filter step functions (which tests/test_core_host.py pins to the oracle) -- and, through the access
tracer, free of races between the lanes whatever the order in which they run a phase."""
import numpy as np
import pytest

import host_core as hc
import quadrotor_landing_b200 as q
from streams_np import norm_rel, rotors_params


def _case(rng, n):
    x = np.zeros(16)
    x[0:3] = rng.normal(0, 1.0, 3) + [0, 0, 2.5]
    x[3:6] = rng.normal(0, 0.5, 3)
    qq = rng.normal(0, 1, 4); qq /= np.linalg.norm(qq)
    if qq[3] < 0:
        qq = -qq
    x[6:10] = qq
    x[10:13] = rng.normal(0, 0.05, 3)
    x[13:16] = rng.normal(0, 0.01, 3)
    A = rng.normal(0, 1, (n, n))
    sc = np.diag(10.0 ** rng.uniform(-2.5, -0.5, n))
    P = sc @ (A @ A.T + n * np.eye(n)) @ sc
    return x, 0.5 * (P + P.T)


@pytest.mark.parametrize("est_bias", [1, 0])
def test_cooperative_prediction_matches_thread_per_filter(est_bias):
    p = rotors_params(q.default_params(), est_bias=est_bias)
    n = 15 if est_bias else 9
    rng = np.random.default_rng(5)
    for trial in range(12):
        x, P = _case(rng, n)
        if not est_bias:
            x[10:16] = 0
        u = np.concatenate([rng.normal(0, 2.0, 3) + [0, 0, 9.8], rng.normal(0, 0.8 if trial % 3 else 60.0, 3)])
        xr, Pr, ar = hc.prediction_step(p, x, P, u)
        for order in range(6):
            xc, Pc, ac, diag = hc.coop_prediction_step(p, x, P, u, order)
            assert diag[0] == 0, "race between lanes (order %d)" % order
            assert diag[1] < 1e-14
            assert norm_rel(xc, xr) < 1e-14 and norm_rel(Pc, Pr) < 1e-13 and norm_rel(ac, ar) < 1e-14


@pytest.mark.parametrize("est_bias,direct", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_cooperative_correction_matches_thread_per_filter(est_bias, direct):
    p = rotors_params(q.default_params(), est_bias=est_bias, direct=direct)
    n = 15 if est_bias else 9
    rng = np.random.default_rng(11)
    for trial in range(12):
        x, P = _case(rng, n)
        if not est_bias:
            x[10:16] = 0
        tq = rng.normal(0, 1, 4); tq /= np.linalg.norm(tq)
        tag = np.concatenate([rng.normal(0, 0.3, 3) + [0, 0, 2.5], tq])
        xr, Pr, obr = hc.correction_step(p, x, P, tag)
        for order in range(6):
            xc, Pc, obc, diag = hc.coop_correction_step(p, x, P, tag, order)
            assert diag[0] == 0, "race between lanes (order %d)" % order
            assert norm_rel(xc, xr) < 1e-12 and norm_rel(Pc, Pr) < 1e-11 and norm_rel(obc, obr) < 1e-13


# ---- the whole replay loop: three threads per filter, barriers as on the device --------------------------------
from oracle import ekf_oracle as orc, noise_np          # noqa: E402
from quadrotor_landing_b200 import scenario             # noqa: E402
from streams_np import noisy_streams                    # noqa: E402
from test_monte_carlo_host import make_noise, short_scenario   # noqa: E402

TOL = 1e-9


@pytest.mark.parametrize("est_bias,direct", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_cooperative_replay_matches_oracle(est_bias, direct):
    """1500 ticks, 3 filters with independent noise, a common and a per-filter dropout, replayed in two chunks (the
    latched measurement, the lanes' private covariance entries and the split nominal state must survive the cut)."""
    p = rotors_params(q.default_params(), est_bias=est_bias, direct=direct)
    scn = scenario.generate(p)
    N, T = 3, 1500
    st = noisy_streams(scn, N, seed=11, T=T, dropout=(600, 800), random_dropout_ticks=150)
    ob = orc.Batch(orc.params_from(p), N)
    hb = hc.CoopHostBatch(p, N)
    for k0, n in ((0, 703), (703, 797)):
        ob.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        hb.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        assert hb.races == 0
        assert norm_rel(hb.state(), ob.state()) < TOL
        assert norm_rel(hb.cov(), ob.cov()) < TOL
        fl = ob.flags()
        assert np.array_equal(hb.upds, fl[4])
        assert np.array_equal(hb.flags & 1, fl[0]) and np.array_equal((hb.flags >> 1) & 1, fl[1])
        oa = ob.aux()
        assert norm_rel(hb.aux[0:3], oa[0:3]) < TOL                       # accel_rel
        assert norm_rel(hb.aux[3:6], oa[3:6]) < TOL and norm_rel(hb.aux[6:10], oa[6:10]) < TOL   # r_t_vt_obs, q_tv_obs
    assert ob.counts()[1] > 100 * N // 2


def test_cooperative_monte_carlo_matches_oracle_with_statistics():
    p = rotors_params(q.default_params())
    scn = short_scenario(p)
    noise = make_noise(first=1000)
    N = 4
    hb = hc.CoopHostBatch(p, N)
    stride = 400
    nb = scn.T // stride
    acc = np.zeros((32, nb, 20))
    hb.run_mc(scn, noise, 0, 777, acc, stride)
    assert hb.races == 0
    hb.run_mc(scn, noise, 777, scn.T - 777, acc, stride)
    assert hb.races == 0
    st = hc.synthesize(scn, noise, 0, N)
    ob = orc.Batch(orc.params_from(p), N)
    n = ob.n
    stats_ref = np.zeros((nb, 20))
    for b in range(nb):
        ob.run(b * stride, stride, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        e, nees = noise_np.error_stats(ob.state(), ob.cov(), scn.truth[(b + 1) * stride], st["bias"], n)
        stats_ref[b, 0:n] = (e ** 2).sum(axis=1)
        stats_ref[b, 15] = nees.sum(); stats_ref[b, 16] = N
        stats_ref[b, 17] = np.sum((nees >= 6.262137795043251) & (nees <= 27.488392863442982))
        stats_ref[b, 19] = (e[0:3] ** 2).sum()
    assert norm_rel(hb.state(), ob.state()) < TOL and norm_rel(hb.cov(), ob.cov()) < TOL
    stats = acc.sum(axis=0)
    assert np.array_equal(stats[:, 16:19], stats_ref[:, 16:19])
    assert norm_rel(stats[:, 0:16], stats_ref[:, 0:16]) < TOL
    assert norm_rel(stats[:, 19], stats_ref[:, 19]) < TOL
