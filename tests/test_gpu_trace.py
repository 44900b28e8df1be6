"""GPU tier: trace export of a sampled subset of a Monte-Carlo batch (quadrotor_landing_b200/trace.py).  Cutting
the fused replay into launches must not change anything (bit-identical end state), and every traced row must
equal what the oracle holds at that tick when it replays the dumped realisation of that filter."""
import numpy as np
import pytest

import quadrotor_landing_b200 as q
from oracle import ekf_oracle as orc
from quadrotor_landing_b200.trace import COLUMNS, trace_monte_carlo
from streams_np import norm_rel, rotors_params
from test_multirate_host import delayed_scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("multirate", [0, 1])
def test_trace_of_a_subset(tmp_path, multirate):
    p = rotors_params(q.default_params(), multirate=bool(multirate))
    scn = delayed_scenario(p, 0.030 if multirate else 0.0, seconds=6.0)
    noise = q.default_noise()
    noise.first_global_id = 4242
    noise.dropout_k0, noise.dropout_k1 = 500, 620
    noise.rand_dropout_len, noise.rand_dropout_lo, noise.rand_dropout_hi = 100, 100, 900
    N, every, picks = 500, 37, [3, 255, 499]
    b = q.BatchEKF(p, N)
    path = str(tmp_path / "trace.csv")
    rows = trace_monte_carlo(b, scn, noise, picks, every, path)
    whole = q.BatchEKF(p, N)
    whole.run_monte_carlo(scn, noise)
    assert np.array_equal(b.state(), whole.state()) and np.array_equal(b.cov(), whole.cov())
    lines = [ln.strip().split(",") for ln in open(path)]
    assert lines[0] == COLUMNS and len(lines) - 1 == rows == len(picks) * -(-scn.T // every)
    idx = [0, 1, 2, 6, 7, 8]
    worst = 0.0
    for f in picks:
        st = b.synthesize_streams(scn, noise, f, 1)
        ob = orc.Batch(orc.params_from(p), 1)
        k = 0
        for r in [ln for ln in lines[1:] if int(ln[0]) == f]:
            tick = int(r[1])
            ob.run(k, tick + 1 - k, st["imu"], st["tag_step"], st["tag_pose"], st["tag_stamp"], st["tag_valid"])
            k = tick + 1
            fl = ob.flags()[:, 0]
            assert int(r[3]) == fl[3]
            if not fl[3]:
                continue
            v = np.array([float(z) for z in r[4:]])
            x, P, aux = ob.state()[:, 0], ob.cov()[:, :, 0], ob.aux()[:, 0]
            worst = max(worst, norm_rel(v[0:3], x[0:3]), norm_rel(v[3:7], x[6:10]),
                        norm_rel(v[7:43], P[np.ix_(idx, idx)].reshape(-1)), norm_rel(v[49:52], x[3:6]))
            assert int(v[55]) == fl[4] and int(v[56]) == fl[2]
            worst = max(worst, norm_rel(v[57:64], aux[3:10]))
    assert worst < 1e-9
    b.close(); whole.close()
