"""quadrotor_landing_b200 -- B200-native batched relative-pose error-state EKF.

The product is the CUDA library `lib/libqekf.so` (C ABI in include/qekf.h); this package is the thin
Python host side: the reference's estimator interface (`RelativePoseEKF`) and the batch interface
(`BatchEKF`).  Importing the package does not load the library; the first use does, and fails loudly
if it is not built.
"""
from ._native import (QEKF_FP32, QEKF_FP64, PF_DELAY, PF_Q, PF_Q_VC, PF_R, PF_R_V_CV, STAT_DIM, QekfError,
                      QekfNoiseSpec, QekfParams, QekfSharedStreams, default_noise, default_params,
                      params_from_yaml, params_from_yaml_text)
from .ekf import BatchEKF, RelativePoseEKF

__all__ = ["BatchEKF", "RelativePoseEKF", "QekfParams", "QekfError", "default_params", "default_noise",
           "QekfNoiseSpec", "QekfSharedStreams", "params_from_yaml", "params_from_yaml_text", "STAT_DIM", "QEKF_FP64", "QEKF_FP32",
           "PF_Q", "PF_R", "PF_R_V_CV", "PF_Q_VC", "PF_DELAY"]
