"""Smallest program that launches the hot kernel at benchmark shape (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from perf_probe import probe
import quadrotor_landing_b200 as q
if __name__ == "__main__":
    prec = q.QEKF_FP32 if "fp32" in sys.argv else q.QEKF_FP64
    probe(1 << 20, 120, prec, reps=1)
