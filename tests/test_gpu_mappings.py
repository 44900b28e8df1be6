"""GPU tier: the two cooperative mappings of the fused replay (qekf_set_mapping: 2 = two role-specialised warps per 32
filters, ekf_duo.cuh; 3 = three lanes per filter, ekf_coop.cuh) through the C ABI against the CPU oracle, 1e-9 norm-relative
on state and covariance -- explicit streams in three launches with a ragged last CTA, and the Monte-Carlo path with
statistics against the thread-per-filter mapping (itself oracle-pinned by tests/test_gpu_parity.py)."""
import numpy as np
import pytest

import quadrotor_landing_b200 as q
from oracle import ekf_oracle as orc
from quadrotor_landing_b200 import scenario
from streams_np import noisy_streams, norm_rel, rotors_params
from test_monte_carlo_host import make_noise, short_scenario

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("lanes", [2, 3])
@pytest.mark.parametrize("est_bias,direct", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_mapping_replay_matches_oracle(lanes, est_bias, direct):
    p = rotors_params(q.default_params(), est_bias=est_bias, direct=direct)
    scn = scenario.generate(p)
    N, T = 200, 2400       # 200 filters: a ragged last CTA whatever the CTA size
    st = noisy_streams(scn, N, seed=21, T=T, dropout=(1000, 1300), random_dropout_ticks=200)
    ob = orc.Batch(orc.params_from(p), N)
    b = q.BatchEKF(p, N)
    b.set_mapping(lanes)
    for k0, n in ((0, 1003), (1003, 698), (1701, 699)):
        ob.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        b.run(k0, n, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        assert norm_rel(b.state(), ob.state()) < TOL
        assert norm_rel(b.cov(), ob.cov()) < TOL
        fl, gf = ob.flags(), b.flags()
        assert np.array_equal(gf[4], fl[4]) and np.array_equal(gf[0], fl[0]) and np.array_equal(gf[1], fl[1])
        assert norm_rel(b.aux()[0:10], ob.aux()[0:10]) < TOL
    assert b.step_counts() == ob.counts()
    b.close()


@pytest.mark.parametrize("lanes", [2, 3])
def test_mapping_monte_carlo_and_statistics_match_thread_per_filter(lanes):
    p = rotors_params(q.default_params())
    scn = short_scenario(p)
    noise = make_noise(first=1000)
    N, stride = 1000, 400
    out = []
    for m in (1, lanes):
        b = q.BatchEKF(p, N)
        b.set_mapping(m)
        b.stats_configure(scn.T // stride, stride)
        b.run_monte_carlo(scn, noise, 0, 777)
        b.run_monte_carlo(scn, noise, 777, scn.T - 777)
        out.append((b.state(), b.cov(), b.stats(), b.step_counts(), b.flags()))
        b.close()
    ref, got = out
    assert norm_rel(got[0], ref[0]) < TOL and norm_rel(got[1], ref[1]) < TOL
    assert np.array_equal(got[2][:, 16:19], ref[2][:, 16:19]) and norm_rel(got[2], ref[2]) < TOL
    assert got[3] == ref[3] and np.array_equal(got[4], ref[4])


def test_mapping_per_filter_parameters_match_thread_per_filter():
    """Config-5 style per-filter Q / R / extrinsic overrides through the cooperative kernels."""
    p = rotors_params(q.default_params())
    scn = scenario.generate(p)
    N, T = 96, 1200
    st = noisy_streams(scn, N, seed=5, T=T)
    rng = np.random.default_rng(3)
    Q = np.array(list(p.Q_a) + list(p.Q_w) + list(p.Q_ab) + list(p.Q_wb))[:, None] * rng.uniform(0.5, 2.0, size=(12, N))
    R = np.array(list(p.R_r) + list(p.R_ang))[:, None] * rng.uniform(0.5, 2.0, size=(6, N))
    out = []
    for m in (1, 2, 3):
        b = q.BatchEKF(p, N)
        b.set_filter_params(q.PF_Q, Q)
        b.set_filter_params(q.PF_R, R)
        b.set_mapping(m)
        b.run(0, T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        out.append((b.state(), b.cov()))
        b.close()
    for got in out[1:]:
        assert norm_rel(got[0], out[0][0]) < TOL and norm_rel(got[1], out[0][1]) < TOL
