"""Two Monte-Carlo launches at benchmark shape: ticks [0,2000) then [2000,2300) (the second one is the one to profile)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
import torch
import quadrotor_landing_b200 as q
from quadrotor_landing_b200 import scenario
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
prec = q.QEKF_FP32 if "fp32" in sys.argv else q.QEKF_FP64
p = bench.bench_params(q)
scn = scenario.generate(p)
noise = bench.bench_noise(q)
b = q.BatchEKF(p, N, precision=prec)
b.stats_configure(scn.T // 200, 200)
b.run_monte_carlo(scn, noise, 0, 2000)
s = torch.cuda.current_stream(); b.set_stream(s.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
b.run_monte_carlo(scn, noise, 2000, 300, sync=False)
e1.record(s); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("N=%d ticks 2000..2300 fp%d: %.2f ms -> %.3e filter-steps/s" % (N, prec, ms, N * 300 / (ms * 1e-3)))
