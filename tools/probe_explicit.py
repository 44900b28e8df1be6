import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools"); sys.path.insert(0, "/root/repo/tests")
import perf_probe as pp
import quadrotor_landing_b200 as q
pp.probe(262144, 600, q.QEKF_FP64)
pp.probe(262144, 600, q.QEKF_FP32)
