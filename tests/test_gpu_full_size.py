"""GPU tier, BASELINE.json's full sizes: 4,096 (config 2) and 1,048,576 (config 3) Monte-Carlo filters over the
whole 60 s landing (12,000 ticks, common + per-filter tag dropouts).  The oracle cannot replay a million
filters, so parity is checked on windows of the global id range (the device dumps the realisation of exactly
those filters and the oracle replays it), and the whole batch through size-independent properties: every state
finite, every sampled covariance symmetric positive-definite, one statistics sample per filter and second,
no divergence, position RMSE at the level the scenario's noise implies, NEES/RMSE sums of the windows equal to
the oracle's."""
import importlib.util
import os

import numpy as np
import pytest

import quadrotor_landing_b200 as q
from oracle import ekf_oracle as orc
from quadrotor_landing_b200 import scenario
from streams_np import norm_rel

pytestmark = pytest.mark.gpu
TOL = 1e-9

_spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
bench = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(bench)


@pytest.mark.parametrize("N,windows", [(4096, [(0, 96), (2000, 64), (4032, 64)]),
                                       (1 << 20, [(0, 48), (524288 - 16, 32), ((1 << 20) - 48, 48)])])
def test_full_size_monte_carlo(N, windows):
    p = bench.bench_params(q)
    scn = scenario.generate(p)
    noise = bench.bench_noise(q)
    stride = 200
    nb = scn.T // stride
    b = q.BatchEKF(p, N)
    b.stats_configure(nb, stride)
    b.run_monte_carlo(scn, noise, 0, 5003)                     # two launches: state and skew survive the cut
    b.run_monte_carlo(scn, noise, 5003, scn.T - 5003)
    stats = b.stats()
    assert np.array_equal(stats[:, 16], np.full(nb, float(N)))   # every filter sampled once per second
    assert stats[:, 18].sum() == 0                               # nothing diverged
    rmse_pos = np.sqrt(stats[:, 19] / stats[:, 16] / 3)
    assert rmse_pos[-1] < 0.02 and rmse_pos.max() < 0.2
    for first, count in windows:
        x, P = b.state(first, count), b.cov(first, count)
        assert np.isfinite(x).all() and np.isfinite(P).all()
        assert np.all(np.linalg.eigvalsh(P.transpose(2, 0, 1)) > 0)
        st = b.synthesize_streams(scn, noise, first, count)
        ob = orc.Batch(orc.params_from(p), count)
        ob.run(0, scn.T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        assert norm_rel(x, ob.state()) < TOL and norm_rel(P, ob.cov()) < TOL
        assert np.array_equal(b.flags(first, count)[0:5], ob.flags()[0:5])
    # spot check of the whole batch: finite everywhere (state is 134 MB at 1M filters)
    assert np.isfinite(b.state()).all()
    b.close()


def test_full_size_fp32_mode():
    """BASELINE config 3, FP32 mode: 1,048,576 filters x 12,000 ticks in FP32 (Joseph-form update, two CTAs per SM).
    Every filter sampled every second, nothing diverged, windows of the id range within 1e-4 of FP64 handles that
    own exactly those global ids, their covariances symmetric positive-definite."""
    p = bench.bench_params(q)
    scn = scenario.generate(p)
    noise = bench.bench_noise(q)
    N, stride = 1 << 20, 200
    nb = scn.T // stride
    b = q.BatchEKF(p, N, precision=q.QEKF_FP32)
    b.stats_configure(nb, stride)
    b.run_monte_carlo(scn, noise)
    stats = b.stats()
    assert np.array_equal(stats[:, 16], np.full(nb, float(N))) and stats[:, 18].sum() == 0
    for first, count in [(0, 64), (777777, 64), (N - 64, 64)]:
        ref = q.BatchEKF(p, count)
        nz = bench.bench_noise(q, first_global_id=first)
        ref.run_monte_carlo(scn, nz)
        x32, P32 = b.state(first, count), b.cov(first, count)
        assert norm_rel(x32, ref.state()) < 1e-4 and norm_rel(P32, ref.cov()) < 1e-4
        assert np.all(np.linalg.eigvalsh(P32.transpose(2, 0, 1)) > 0)
        ref.close()
    b.close()


def test_full_size_delayed_fusion():
    """Delayed-measurement fusion at scale: 262,144 filters x 12,000 ticks, dynamic delay, 30 ms tag latency.
    Windows of the id range against the oracle (which keeps the reference's full history vectors), the whole batch
    through the statistics."""
    p = bench.bench_params(q, multirate=True, dynamic=True)
    scn = bench.bench_scenario(q, p)
    noise = bench.bench_noise(q)
    N, stride = 1 << 18, 200
    nb = scn.T // stride
    b = q.BatchEKF(p, N)
    b.stats_configure(nb, stride)
    b.run_monte_carlo(scn, noise, 0, 7001)
    b.run_monte_carlo(scn, noise, 7001, scn.T - 7001)
    stats = b.stats()
    assert np.array_equal(stats[:, 16], np.full(nb, float(N))) and stats[:, 18].sum() == 0
    for first, count in [(0, 48), (N // 2 + 5, 32), (N - 48, 48)]:
        st = b.synthesize_streams(scn, noise, first, count)
        ob = orc.Batch(orc.params_from(p), count)
        ob.run(0, scn.T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        assert norm_rel(b.state(first, count), ob.state()) < TOL and norm_rel(b.cov(first, count), ob.cov()) < TOL
        assert np.array_equal(b.flags(first, count), ob.flags())          # incl. x_hist.size()
    npred, ncorr = b.step_counts()
    assert npred < 1.15 * N * scn.T                                      # lazily evaluated history: ~1.08 predictions per tick
    b.close()


def test_full_size_config5_sweep_delayed_dynamic():
    """BASELINE config 5 at full size: 1,048,576 filters x 12,000 ticks, every filter with its own Q, R and camera
    extrinsic (qekf_set_filter_params), delayed-measurement fusion with the stamp-derived (dynamic) delay -- the
    workload of bench.py --sweep --multirate --dynamic-delay.  Windows of the id range against the oracle carrying
    the same per-filter parameter sets and the reference's full history vectors (relative_pose_EKF.cpp:196-264); the
    whole batch through the statistics."""
    p = bench.bench_params(q, multirate=True, dynamic=True)
    scn = bench.bench_scenario(q, p)
    noise = bench.bench_noise(q)
    N, stride = 1 << 20, 200
    nb = scn.T // stride
    b = q.BatchEKF(p, N)
    vals = bench.apply_sweep(q, b, p, N, seed=1234)
    b.stats_configure(nb, stride)
    b.run_monte_carlo(scn, noise, 0, 6007)
    b.run_monte_carlo(scn, noise, 6007, scn.T - 6007)
    stats = b.stats()
    assert np.array_equal(stats[:, 16], np.full(nb, float(N)))          # every filter sampled once per second
    # (a sweep over two decades of Q and R is allowed a few inconsistent filters, but no non-finite state)
    assert stats[:, 18].sum() <= 1e-5 * stats[:, 16].sum()
    for first, count in [(0, 40), (N // 2 - 13, 32), (N - 40, 40)]:
        st = b.synthesize_streams(scn, noise, first, count)
        ob = orc.Batch(orc.params_from(p), count)
        for field, v in vals.items():
            ob.set_filter_params(field, np.ascontiguousarray(v[:, first:first + count]))
        ob.run(0, scn.T, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
        assert norm_rel(b.state(first, count), ob.state()) < TOL and norm_rel(b.cov(first, count), ob.cov()) < TOL
        assert np.array_equal(b.flags(first, count), ob.flags())          # incl. x_hist.size()
        assert norm_rel(b.aux(first, count), ob.aux()) < TOL              # incl. measurement_delay_curr
    assert np.isfinite(b.state()).all()
    b.close()
