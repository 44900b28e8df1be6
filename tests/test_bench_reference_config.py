"""`bench.py --impl reference` must time the SAME workload as the GPU arm without loading the product library: the
oracle-side preset, scenario generator (oracle/libscenario_ref.so) and noise specification equal the product's."""
import importlib.util
import os
import subprocess
import sys

import numpy as np

import quadrotor_landing_b200 as q
from oracle import bench_ref, ekf_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_reference_arm_workload_equals_product_arm_workload():
    bench = _bench()
    for mr, dyn in ((False, False), (True, False), (True, True)):
        p = bench.bench_params(q, mr, dyn)
        assert bytes(orc.params_from(p)) == bytes(bench_ref.params(mr, dyn))
        a, b = bench.bench_scenario(q, p), bench_ref.scenario(bench_ref.params(mr, dyn), 0.030 if mr else 0.0)
        assert (a.T, a.M) == (b.T, b.M)
        for f in ("truth", "imu_clean", "tag_step", "tag_pose_clean", "tag_stamp"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), f
    n, r = bench.bench_noise(q), bench_ref.noise()
    for f, _ in n._fields_:
        assert getattr(n, f) == getattr(r, f), f


def test_reference_arm_does_not_load_the_product_library():
    code = ("import sys, bench; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0'];"
            "import oracle.bench_ref as br; p=br.params(); s=br.scenario(p);"
            "st=bench.cpu_streams(br.noise(), s, 2); bench.cpu_replay(p, s, st, 1);"
            "maps=open('/proc/self/maps').read();"
            "assert 'libqekf' not in maps and 'quadrotor_landing_b200' not in sys.modules, 'product loaded';"
            "assert 'libekf_oracle' in maps and 'libscenario_ref' in maps; print('ok')")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
