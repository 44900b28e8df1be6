"""The per-call flop counts bench.py's roofline uses are the ones the product's structured code executes
(instrumented scalar type, tests/host_core hc_count_flops), and they undercut the dense formulas."""
import importlib.util
import os

import host_core as hc
import quadrotor_landing_b200 as q

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_flop_table_matches_instrumented_count():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for (b, d), (fp, fc) in bench.FLOPS.items():
        p = q.default_params()
        p.est_bias, p.direct_orien_method = b, d
        assert hc.count_flops(p) == (fp, fc)
    # dense reference formulas: F P F^T + W Q W^T ~ 2*(15^3*2 + 15*12*12 + 15*12*15) = 23.2 kflop
    assert bench.FLOPS[(1, 1)][0] < 23200 / 10
