"""Quick device-side throughput probe (explicit device-resident streams built with torch)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import quadrotor_landing_b200 as q
from quadrotor_landing_b200 import scenario
from streams_np import rotors_params


def probe(N, T, precision, est_bias=1, direct=1, reps=3):
    p = rotors_params(q.default_params(), est_bias=est_bias, direct=direct)
    scn = scenario.generate(p)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(1)
    sel = scn.tag_step < T
    steps = torch.tensor(scn.tag_step[sel], dtype=torch.int32, device=dev)
    M = int(sel.sum())
    imu = torch.tensor(scn.imu_clean[:T], device=dev)[:, :, None] + 0.02 * torch.randn((T, 6, N), dtype=torch.float64, device=dev, generator=g)
    pose = torch.tensor(scn.tag_pose_clean[sel], device=dev)[:, :, None].repeat(1, 1, N)
    pose[:, 0:3] += 0.02 * torch.randn((M, 3, N), dtype=torch.float64, device=dev, generator=g)
    stamp = torch.tensor(scn.tag_stamp[sel], device=dev)
    b = q.BatchEKF(p, N, precision=precision)
    s = torch.cuda.current_stream()
    b.set_stream(s.cuda_stream)
    times = []
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        b.run_device(0, T, T, imu.data_ptr(), M, steps.data_ptr(), pose.data_ptr(), stamp.data_ptr())
        e1.record(s)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    t = min(times[1:])
    x = b.state()
    print("N=%d T=%d fp%d bias=%d direct=%d: %.2f ms  -> %.3e filter-steps/s  (finite=%s)" % (
        N, T, precision, est_bias, direct, t, N * T / (t * 1e-3), bool(np.isfinite(x).all())), flush=True)
    b.close()
    del imu, pose
    torch.cuda.empty_cache()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for N in (4096, 65536, 1 << 20):
        T = 240 if N > 100000 else 2000
        probe(N, T, q.QEKF_FP64)
        probe(N, T, q.QEKF_FP32)
    probe(1 << 20, 240, q.QEKF_FP64, est_bias=0)
    probe(1 << 20, 240, q.QEKF_FP64, direct=0)
