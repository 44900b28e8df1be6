"""Synthetic landing scenario (thin wrapper over the C++ generator in csrc/scenario.hpp).

Returns the shared clean trajectory: truth states, clean IMU samples and clean tag poses with their
arrival ticks and capture stamps.  Per-filter noise realisations are added on the device.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as nat
from ._native import QekfParams, QekfScenarioSpec, check


@dataclass
class Scenario:
    T: int
    M: int
    truth: np.ndarray        # [T+1, 10]  r, v, q_tv after j ticks (state after tick k is truth[k+1])
    imu_clean: np.ndarray    # [T, 6]
    tag_step: np.ndarray     # [M] int32
    tag_pose_clean: np.ndarray  # [M, 7]
    tag_stamp: np.ndarray    # [M]
    spec: QekfScenarioSpec


def default_spec() -> QekfScenarioSpec:
    s = QekfScenarioSpec()
    check(nat.lib().qekf_scenario_default(C.byref(s)))
    return s


def generate(params: QekfParams, spec: QekfScenarioSpec | None = None) -> Scenario:
    L = nat.lib()
    spec = spec if spec is not None else default_spec()
    T, M = C.c_int64(), C.c_int64()
    check(L.qekf_scenario_sizes(C.byref(params), C.byref(spec), C.byref(T), C.byref(M)))
    T, M = T.value, M.value
    truth = np.zeros((T + 1, 10)); imu = np.zeros((T, 6))
    step = np.zeros(M, dtype=np.int32); pose = np.zeros((M, 7)); stamp = np.zeros(M)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    check(L.qekf_scenario_generate(C.byref(params), C.byref(spec), truth.ctypes.data_as(dp), imu.ctypes.data_as(dp),
                                   step.ctypes.data_as(ip), pose.ctypes.data_as(dp), stamp.ctypes.data_as(dp)))
    return Scenario(T, M, truth, imu, step, pose, stamp, spec)
