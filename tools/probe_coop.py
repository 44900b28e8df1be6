"""A/B probe of the two mappings of the fused replay on one GPU: parity of the three-lane kernel against the
thread-per-filter kernel on the benchmark workload (small batch), then CUDA-event timings of both on ticks
2000..2000+n of the benchmark scenario (private dropouts and statistics on), for every CTA size of the build.
Usage: probe_coop.py [N] [n=<ticks>] [nostats] [nodrop]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
import numpy as np
import torch
import quadrotor_landing_b200 as q
from quadrotor_landing_b200 import scenario
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)

args = sys.argv[1:]
N = int(args[0]) if args and args[0].isdigit() else 524288
n2 = 300
for a in args:
    if a.startswith("n="):
        n2 = int(a[2:])
p = bench.bench_params(q)
scn = scenario.generate(p, scenario.default_spec())
noise = bench.bench_noise(q)
if "nodrop" in args:
    noise.rand_dropout_len = 0


def nrel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def run_small(lanes, groups, Ns=4096, T=3000):
    b = q.BatchEKF(p, Ns)
    b.set_mapping(lanes, groups)
    b.stats_configure(scn.T // 200, 200)
    b.run_monte_carlo(scn, noise, 0, 1777)
    b.run_monte_carlo(scn, noise, 1777, T - 1777)
    out = b.state(), b.cov(), b.stats(), b.aux(), b.flags(), b.step_counts()
    b.close()
    return out


ref = run_small(1, 0)
for lanes, g in ((2, 6), (2, 5), (2, 4), (3, 4)):
    try:
        got = run_small(lanes, g)
    except Exception as e:   # noqa: BLE001
        print("groups=%d: %s" % (g, e), flush=True)
        continue
    print("parity lanes=%d groups=%d: state %.2e cov %.2e stats %.2e aux %.2e flags_equal=%s counts %s vs %s" % (
        lanes, g, nrel(got[0], ref[0]), nrel(got[1], ref[1]), nrel(got[2], ref[2]), nrel(got[3], ref[3]),
        bool(np.array_equal(got[4], ref[4])), got[5], ref[5]), flush=True)


def timed(lanes, groups):
    b = q.BatchEKF(p, N)
    b.set_mapping(lanes, groups)
    if "nostats" not in args:
        b.stats_configure(scn.T // 200, 200)
    b.run_monte_carlo(scn, noise, 0, 2000)
    b.step_counts(reset=True)
    s = torch.cuda.current_stream(); b.set_stream(s.cuda_stream)
    best = 1e30
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        b.run_monte_carlo(scn, noise, 2000 + rep * n2, n2, sync=False)
        e1.record(s); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    npred, ncorr = b.step_counts()
    b.close()
    print("N=%d lanes=%d groups=%d: %.2f ms -> %.3e filter-steps/s (pred/tick %.3f corr/tick %.4f)" % (
        N, lanes, groups, best, N * n2 / (best * 1e-3), npred / (2 * N * n2), ncorr / (2 * N * n2)), flush=True)


timed(1, 0)
for lanes, g in ((2, 6), (2, 5), (2, 4), (3, 4)):
    try:
        timed(lanes, g)
    except Exception as e:   # noqa: BLE001
        print("groups=%d: %s" % (g, e), flush=True)
