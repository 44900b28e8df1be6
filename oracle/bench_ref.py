"""The benchmark workload for the CPU arm of bench.py (`--impl reference`) WITHOUT the product package: parameter preset,
clean scenario, noise specification.  BASELINE INFRASTRUCTURE: only bench.py's reference / cpu_baseline legs and tests/
import this.  tests/test_bench_reference_config.py checks that every piece equals what the product's own arm uses.

Loads oracle/libekf_oracle.so and oracle/libscenario_ref.so only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import numpy as np

from . import ekf_oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libscenario_ref.so")
_SRC = [os.path.join(_HERE, "scenario_ref.cpp"), os.path.join(_HERE, "..", "quadrotor_landing_b200", "csrc", "scenario.hpp"),
        os.path.join(_HERE, "..", "include", "qekf.h")]


class ScenarioSpec(C.Structure):      # qekf_scenario_spec (include/qekf.h)
    _fields_ = [(n, C.c_double) for n in (
        "duration_s", "hover_s", "z_start", "z_end", "sway_ax", "sway_wx", "sway_ay", "sway_wy", "sway_phase_y",
        "yaw_amp", "yaw_w", "tag_rate_hz", "tag_latency_s", "t_start")]


def build(force=False):
    orc.build(force)
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in _SRC):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-s", "-C", _HERE, "libscenario_ref.so"] + (["-B"] if force else []), check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return _LIB


_lib = None


def _L():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
    return _lib


def params(multirate=False, dynamic=False) -> orc.OrcParams:
    """bench.py's bench_params() on the oracle's parameter struct: rotors.yaml noises
    (quad_state_estimation/config/relative_pose_EKF_rotors.yaml:13-19), 200 Hz update, 30 Hz tag, gated, direct model."""
    p = orc.OrcParams()
    orc.lib().orc_default_params(C.byref(p))
    p.update_freq, p.measurement_freq = 200.0, 30.0
    p.measurement_delay, p.measurement_delay_max, p.dyn_measurement_delay_offset = 0.030, 0.200, 0.005
    for i in range(3):
        p.Q_a[i], p.Q_w[i], p.Q_ab[i], p.Q_wb[i] = 0.0005, 0.00005, 5.0e-5, 5.0e-6
    p.R_r[0], p.R_r[1], p.R_r[2] = 0.015, 0.015, 0.020
    p.R_ang[0], p.R_ang[1], p.R_ang[2] = 0.0015, 0.0015, 0.04
    p.limit_measurement_freq = p.corner_margin_enbl = p.est_bias = p.direct_orien_method = 1
    p.multirate_ekf, p.dynamic_meas_delay = int(multirate), int(dynamic)
    return p


def scenario(p: orc.OrcParams, tag_latency_s=0.0):
    L = _L()
    s = ScenarioSpec()
    L.scn_defaults(C.byref(s))
    s.tag_latency_s = tag_latency_s
    T, M = C.c_int64(), C.c_int64()
    L.scn_sizes(C.byref(p), C.byref(s), C.byref(T), C.byref(M))
    T, M = T.value, M.value
    truth = np.zeros((T + 1, 10)); imu = np.zeros((T, 6))
    step = np.zeros(M, dtype=np.int32); pose = np.zeros((M, 7)); stamp = np.zeros(M)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.scn_generate(C.byref(p), C.byref(s), truth.ctypes.data_as(dp), imu.ctypes.data_as(dp), step.ctypes.data_as(ip),
                   pose.ctypes.data_as(dp), stamp.ctypes.data_as(dp))
    return SimpleNamespace(T=T, M=M, truth=truth, imu_clean=imu, tag_step=step, tag_pose_clean=pose, tag_stamp=stamp, spec=s)


def noise(first_global_id=0):
    """bench.py's bench_noise(): the library's default sigmas (qekf_noise_default) + the benchmark's dropout windows."""
    return SimpleNamespace(seed=0x5EED, first_global_id=first_global_id, sigma_accel=0.02, sigma_gyro=0.007,
                           sigma_bias_accel=0.05, sigma_bias_gyro=0.002, sigma_tag_pos=0.02, sigma_tag_ang=0.01,
                           dropout_k0=5000, dropout_k1=5400, rand_dropout_len=200, rand_dropout_lo=400,
                           rand_dropout_hi=11600, edge_loss=0, range_ref=0.0, range_exp_pos=0.0, range_exp_ang=0.0)
