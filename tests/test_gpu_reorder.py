"""GPU tier: Monte-Carlo launches on reordered arrays.

qekf_run_monte_carlo sorts the filters by the start of their private tag dropout and, for long replays, gathers every
per-filter array into that order before the launch and scatters it back after (filters that share a CTA then lose
and regain their measurements together).  Filters are independent and the noise is keyed by the filter id, so nothing a
caller can observe may depend on it: every accessor must return, bit for bit, what the un-reordered launch returns --
single-rate, delayed fusion (checkpoint, IMU ring and history counters travel too), per-filter parameter tables
(BASELINE config 5), FP32, ragged sizes, chunked launches that mix reordered and plain launches."""
import numpy as np
import pytest

import quadrotor_landing_b200 as q
from quadrotor_landing_b200 import scenario
from streams_np import rotors_params
from test_multirate_host import delayed_scenario, sweep_values

pytestmark = pytest.mark.gpu


def _noise(first=0):
    n = q.default_noise()
    n.seed = 2024
    n.first_global_id = first
    n.dropout_k0, n.dropout_k1 = 500, 640
    n.rand_dropout_len, n.rand_dropout_lo, n.rand_dropout_hi = 160, 100, 1300
    return n


def _run(monkeypatch, env, p, scn, N, precision, chunks, sweep, stats=(15, 100)):
    for k in ("QEKF_NO_PERM", "QEKF_PERM_MIN_STEPS"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)           # read by qekf_create
    b = q.BatchEKF(p, N, precision=precision)
    if sweep:
        for field, v in sweep_values(np.random.default_rng(9), p, N, bool(p.multirate_ekf)).items():
            b.set_filter_params(field, v)
    b.stats_configure(*stats)
    noise = _noise(first=10_000_000_000)
    for k0, n in chunks:
        b.run_monte_carlo(scn, noise, k0, n)
    out = [b.state(), b.cov(), b.aux(), b.flags()], b.stats(), b.step_counts(), b.launch_count
    b.close()
    return out


@pytest.mark.parametrize("multirate,dynamic,sweep,precision", [(0, 0, 0, 64), (1, 0, 0, 64), (1, 1, 1, 64), (0, 0, 1, 64), (1, 1, 0, 32)])
def test_reordered_launch_is_invisible(monkeypatch, multirate, dynamic, sweep, precision):
    p = rotors_params(q.default_params(), multirate=bool(multirate), dynamic_delay=bool(dynamic))
    scn = delayed_scenario(p, 0.035, seconds=8.0) if multirate else scenario.generate(p)
    N = 4999                                   # ragged last CTA, padding columns
    chunks = ((0, 700), (700, 3), (703, 797))  # the 3-tick launch stays below the reordering threshold
    plain, plain_stats, plain_counts, plain_launches = _run(monkeypatch, {"QEKF_NO_PERM": "1"}, p, scn, N, precision, chunks, sweep)
    perm, perm_stats, perm_counts, perm_launches = _run(monkeypatch, {"QEKF_PERM_MIN_STEPS": "100"}, p, scn, N, precision, chunks, sweep)
    assert perm_launches > plain_launches      # the reordering passes really ran
    for a, b in zip(plain, perm):
        assert np.array_equal(a, b, equal_nan=True)
    # corrections are per-filter semantics; with delayed fusion the number of prediction_step EXECUTIONS is not (a lane
    # without a correction catches its checkpoint up inside its CTA-mates' correction events, so it depends on who they are)
    assert perm_counts[1] == plain_counts[1]
    if not multirate:
        assert perm_counts == plain_counts
    assert np.array_equal(perm_stats[:, 16:19], plain_stats[:, 16:19])        # sample / in-interval / diverged counts
    assert np.allclose(perm_stats, plain_stats, rtol=1e-11, atol=0)           # sums: atomics in another order
    assert plain_stats[:, 16].sum() > 0
