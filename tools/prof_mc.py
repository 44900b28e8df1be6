"""Two Monte-Carlo launches at benchmark shape: ticks [0,2000) then [2000,2000+n) (the second one is the one to
profile).  Flags: fp32, mr (delayed fusion, 30 ms latency), dyn (dynamic delay), nostats, n=<ticks>."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
import torch
import quadrotor_landing_b200 as q
from quadrotor_landing_b200 import scenario
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
args = sys.argv[1:]
N = int(args[0]) if args and args[0].isdigit() else 262144
prec = q.QEKF_FP32 if "fp32" in args else q.QEKF_FP64
n2 = 300
for a in args:
    if a.startswith("n="):
        n2 = int(a[2:])
p = bench.bench_params(q)
sp = scenario.default_spec()
if "mr" in args:
    p.multirate_ekf = 1
    p.dynamic_meas_delay = 1 if "dyn" in args else 0
    sp.tag_latency_s = 0.030
scn = scenario.generate(p, sp)
noise = bench.bench_noise(q)
if "nodrop" in args:
    noise.rand_dropout_len = 0
if "droplate" in args:          # private dropouts exist, but none starts before tick 5000
    noise.rand_dropout_lo = 5000
if "nocommon" in args:
    noise.dropout_k0 = noise.dropout_k1 = 0
b = q.BatchEKF(p, N, precision=prec)
if "nostats" not in args:
    b.stats_configure(scn.T // 200, 200)
b.run_monte_carlo(scn, noise, 0, 2000)
b.step_counts(reset=True)
s = torch.cuda.current_stream(); b.set_stream(s.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
b.run_monte_carlo(scn, noise, 2000, n2, sync=False)
e1.record(s); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
npred, ncorr = b.step_counts()
print("N=%d ticks 2000..%d fp%d %s: %.2f ms -> %.3e filter-steps/s  (predictions/tick %.3f, corrections/tick %.4f)" % (
    N, 2000 + n2, prec, " ".join(a for a in args if not a.isdigit()), ms, N * n2 / (ms * 1e-3), npred / (N * n2), ncorr / (N * n2)))
