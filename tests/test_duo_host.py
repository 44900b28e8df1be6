"""The two-role mapping (ekf_duo.cuh: role A = covariance core + corrections, role B = nominal state + bias columns),
run on the host as two threads per filter meeting at the barriers the device uses, against the oracle."""
import numpy as np
import pytest

import host_core as hc
import quadrotor_landing_b200 as q
from oracle import ekf_oracle as orc, noise_np
from quadrotor_landing_b200 import scenario
from streams_np import noisy_streams, norm_rel, rotors_params
from test_monte_carlo_host import make_noise, short_scenario

TOL = 1e-9


@pytest.mark.parametrize("est_bias,direct", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_two_role_replay_matches_oracle(est_bias, direct):
    """1500 ticks, 3 filters with independent noise, a common and a per-filter dropout, replayed in two chunks (role B
    runs the kinematics one tick ahead: the cut, the dropouts and every correction interrupt that)."""
    p = rotors_params(q.default_params(), est_bias=est_bias, direct=direct)
    scn = scenario.generate(p)
    N, T = 3, 1500
    st = noisy_streams(scn, N, seed=11, T=T, dropout=(600, 800), random_dropout_ticks=150)
    ob = orc.Batch(orc.params_from(p), N)
    hb = hc.DuoHostBatch(p, N)
    for k0, n in ((0, 703), (703, 797)):
        ob.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        hb.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        assert norm_rel(hb.state(), ob.state()) < TOL
        assert norm_rel(hb.cov(), ob.cov()) < TOL
        fl = ob.flags()
        assert np.array_equal(hb.upds, fl[4])
        assert np.array_equal(hb.flags & 1, fl[0]) and np.array_equal((hb.flags >> 1) & 1, fl[1])
        oa = ob.aux()
        assert norm_rel(hb.aux[0:3], oa[0:3]) < TOL
        assert norm_rel(hb.aux[3:6], oa[3:6]) < TOL and norm_rel(hb.aux[6:10], oa[6:10]) < TOL
    assert ob.counts()[1] > 100 * N // 2


def test_two_role_monte_carlo_matches_oracle_with_statistics():
    p = rotors_params(q.default_params())
    scn = short_scenario(p)
    noise = make_noise(first=1000)
    N = 4
    hb = hc.DuoHostBatch(p, N)
    stride = 400
    nb = scn.T // stride
    acc = np.zeros((32, nb, 20))
    hb.run_mc(scn, noise, 0, 777, acc, stride)
    hb.run_mc(scn, noise, 777, scn.T - 777, acc, stride)
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
