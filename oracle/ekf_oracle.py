"""ctypes loader for the CPU ORACLE (oracle/ekf_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (quadrotor_landing_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libekf_oracle.so")
MAX_TAGS = 16

PF_Q, PF_R, PF_R_V_CV, PF_Q_VC, PF_DELAY = range(5)


class OrcParams(C.Structure):
    """Mirror of orc_params (ekf_oracle.h); field meaning follows relative_pose_EKF.hpp:66-133."""

    _fields_ = [
        ("update_freq", C.c_double),
        ("measurement_freq", C.c_double),
        ("measurement_delay", C.c_double),
        ("measurement_delay_max", C.c_double),
        ("dyn_measurement_delay_offset", C.c_double),
        ("Q_a", C.c_double * 3),
        ("Q_w", C.c_double * 3),
        ("Q_ab", C.c_double * 3),
        ("Q_wb", C.c_double * 3),
        ("R_r", C.c_double * 3),
        ("R_ang", C.c_double * 3),
        ("r_cov_init", C.c_double),
        ("v_cov_init", C.c_double),
        ("ang_cov_init", C.c_double),
        ("ab_cov_init", C.c_double),
        ("wb_cov_init", C.c_double),
        ("ab_static", C.c_double * 3),
        ("wb_static", C.c_double * 3),
        ("r_v_cv", C.c_double * 3),
        ("q_vc", C.c_double * 4),
        ("camera_K", C.c_double * 9),
        ("tag_in_view_margin", C.c_double),
        ("tag_widths", C.c_double * MAX_TAGS),
        ("tag_positions", C.c_double * (3 * MAX_TAGS)),
        ("small_ang_tol", C.c_double),
        ("g", C.c_double * 3),
        ("camera_width", C.c_int32),
        ("camera_height", C.c_int32),
        ("n_tags", C.c_int32),
        ("est_bias", C.c_int32),
        ("limit_measurement_freq", C.c_int32),
        ("corner_margin_enbl", C.c_int32),
        ("direct_orien_method", C.c_int32),
        ("multirate_ekf", C.c_int32),
        ("dynamic_meas_delay", C.c_int32),
        ("reserved", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (Makefile in this directory) if the .so is missing or stale."""
    src = os.path.join(_HERE, "ekf_oracle.c")
    hdr = os.path.join(_HERE, "ekf_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in (src, hdr)
    )
    if force or stale:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-s", "-C", _HERE, "-B"], check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int32)
    vp = C.c_void_p
    L.orc_default_params.argtypes = [C.POINTER(OrcParams)]
    L.orc_sizeof_params.restype = C.c_int
    for name in ("orc_quat_exp", "orc_quat_log", "orc_quat_to_rot", "orc_skew"):
        getattr(L, name).argtypes = [dp, dp]
    L.orc_quat_norm.argtypes = [dp]
    L.orc_quat_mul.argtypes = [dp, dp, dp]
    L.orc_create.argtypes = [C.POINTER(OrcParams)]
    L.orc_create.restype = vp
    L.orc_destroy.argtypes = [vp]
    L.orc_set_params.argtypes = [vp, C.POINTER(OrcParams)]
    L.orc_initialize_params.argtypes = [vp]
    L.orc_initialize_state.argtypes = [vp, C.c_int]
    L.orc_set_imu.argtypes = [vp, dp, dp]
    L.orc_set_tag.argtypes = [vp, dp, dp, C.c_double]
    L.orc_filter_update.argtypes = [vp, C.c_double]
    L.orc_get_state.argtypes = [vp, dp]
    L.orc_get_num_states.argtypes = [vp]
    L.orc_get_num_states.restype = C.c_int
    L.orc_get_cov.argtypes = [vp, dp]
    L.orc_set_state.argtypes = [vp, dp, dp]
    L.orc_get_aux.argtypes = [vp, dp]
    L.orc_get_flags.argtypes = [vp, ip]
    L.orc_prediction_step.argtypes = [vp, dp, dp, dp, dp, dp, dp]
    L.orc_correction_step.argtypes = [vp, dp, dp, dp, dp, dp, dp]
    L.orc_batch_create.argtypes = [C.POINTER(OrcParams), C.c_int64]
    L.orc_batch_create.restype = vp
    L.orc_batch_destroy.argtypes = [vp]
    L.orc_batch_set_filter_params.argtypes = [vp, C.c_int, dp]
    L.orc_batch_set_filter_params.restype = C.c_int
    L.orc_batch_run.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, dp, C.c_int64, ip, dp, dp,
                                C.POINTER(C.c_uint8), C.c_double, C.c_int]
    L.orc_batch_get_state.argtypes = [vp, dp]
    L.orc_batch_get_cov.argtypes = [vp, dp]
    L.orc_batch_get_aux.argtypes = [vp, dp]
    L.orc_batch_get_flags.argtypes = [vp, ip]
    L.orc_batch_filter.argtypes = [vp, C.c_int64]
    L.orc_batch_filter.restype = vp
    L.orc_batch_get_counts.argtypes = [vp, C.POINTER(C.c_int64)]
    _lib = L
    return L


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def default_params() -> OrcParams:
    p = OrcParams()
    lib().orc_default_params(C.byref(p))
    return p


def params_from(other) -> OrcParams:
    """Byte-copy a same-layout ctypes params struct (e.g. the product's QekfParams)."""
    assert C.sizeof(other) == C.sizeof(OrcParams)
    return OrcParams.from_buffer_copy(bytes(other))


# ---- helper wrappers ---------------------------------------------------------------------------

def quat_exp(v):
    v = _f64(v); q = np.zeros(4)
    lib().orc_quat_exp(_dp(v), _dp(q))
    return q


def quat_log(q):
    q = _f64(q); v = np.zeros(3)
    lib().orc_quat_log(_dp(q), _dp(v))
    return v


def quat_norm(q):
    q = _f64(q).copy()
    lib().orc_quat_norm(_dp(q))
    return q


def quat_mul(a, b):
    a = _f64(a); b = _f64(b); o = np.zeros(4)
    lib().orc_quat_mul(_dp(a), _dp(b), _dp(o))
    return o


def quat_to_rot(q):
    q = _f64(q); R = np.zeros(9)
    lib().orc_quat_to_rot(_dp(q), _dp(R))
    return R.reshape(3, 3)


def skew(v):
    v = _f64(v); S = np.zeros(9)
    lib().orc_skew(_dp(v), _dp(S))
    return S.reshape(3, 3)


class Filter:
    """One reference-semantics filter (the estimator interface of relative_pose_EKF.hpp:20-144)."""

    def __init__(self, params: OrcParams | None = None, _handle=None, _owner=True):
        self._L = lib()
        self._owner = _owner
        if _handle is not None:
            self._h = _handle
        else:
            p = params if params is not None else default_params()
            self._h = self._L.orc_create(C.byref(p))

    def __del__(self):
        if getattr(self, "_owner", False) and getattr(self, "_h", None):
            self._L.orc_destroy(self._h)
            self._h = None

    @property
    def n(self) -> int:
        return self._L.orc_get_num_states(self._h)

    def set_params(self, p: OrcParams):
        self._L.orc_set_params(self._h, C.byref(p))

    def initialize_state(self, reinit_bias=False):
        self._L.orc_initialize_state(self._h, int(reinit_bias))

    def set_imu(self, accel, gyro):
        a = _f64(accel); w = _f64(gyro)
        self._L.orc_set_imu(self._h, _dp(a), _dp(w))

    def set_tag(self, pos, quat_xyzw, stamp=0.0):
        p = _f64(pos); q = _f64(quat_xyzw)
        self._L.orc_set_tag(self._h, _dp(p), _dp(q), float(stamp))

    def filter_update(self, t_curr=0.0):
        self._L.orc_filter_update(self._h, float(t_curr))

    def state(self):
        x = np.zeros(16)
        self._L.orc_get_state(self._h, _dp(x))
        return x

    def cov(self):
        n = self.n
        P = np.zeros(n * n)
        self._L.orc_get_cov(self._h, _dp(P))
        return P.reshape(n, n)

    def set_state(self, x16, P):
        x = _f64(x16); Pm = _f64(P)
        self._L.orc_set_state(self._h, _dp(x), _dp(Pm))

    def aux(self):
        a = np.zeros(11)
        self._L.orc_get_aux(self._h, _dp(a))
        return {"accel_rel": a[0:3], "r_t_vt_obs": a[3:6], "q_tv_obs": a[6:10],
                "measurement_delay_curr": a[10]}

    def flags(self):
        f = np.zeros(6, dtype=np.int32)
        self._L.orc_get_flags(self._h, _ip(f))
        return {"state_initialized": int(f[0]), "measurement_ready": int(f[1]),
                "performed_correction": int(f[2]), "filter_active": int(f[3]),
                "upds_since_correction": int(f[4]), "hist_len": int(f[5])}

    def prediction_step(self, x, P, u):
        n = self.n
        x = _f64(x); Pm = _f64(P); u = _f64(u)
        xo = np.zeros(16); Po = np.zeros(n * n); acc = np.zeros(3)
        self._L.orc_prediction_step(self._h, _dp(x), _dp(Pm), _dp(u), _dp(xo), _dp(Po), _dp(acc))
        return xo, Po.reshape(n, n), acc

    def correction_step(self, x, P, r_c_tc, q_ct_xyzw):
        n = self.n
        x = _f64(x); Pm = _f64(P); r = _f64(r_c_tc); q = _f64(q_ct_xyzw)
        xo = np.zeros(16); Po = np.zeros(n * n)
        self._L.orc_correction_step(self._h, _dp(x), _dp(Pm), _dp(r), _dp(q), _dp(xo), _dp(Po))
        return xo, Po.reshape(n, n)


class Batch:
    """N independent reference-semantics filters replaying explicit streams (OpenMP over filters)."""

    def __init__(self, params: OrcParams, n_filters: int):
        self._L = lib()
        self.N = int(n_filters)
        self.params = params
        self.n = 15 if params.est_bias else 9
        self._h = self._L.orc_batch_create(C.byref(params), self.N)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_batch_destroy(self._h)
            self._h = None

    def set_filter_params(self, field: int, values):
        v = _f64(values)
        assert v.shape[-1] == self.N
        rc = self._L.orc_batch_set_filter_params(self._h, int(field), _dp(v))
        assert rc == 0

    def run(self, k0, n_steps, imu, tag_step, tag_pose, tag_stamp, tag_valid=None, t_start=0.0,
            n_threads=0):
        imu = _f64(imu)
        T = imu.shape[0]
        assert imu.shape == (T, 6, self.N)
        tag_step = np.ascontiguousarray(tag_step, dtype=np.int32)
        M = tag_step.shape[0]
        tag_pose = _f64(tag_pose)
        assert tag_pose.shape == (M, 7, self.N)
        tag_stamp = _f64(tag_stamp)
        assert tag_stamp.shape == (M,)
        assert k0 + n_steps <= T
        vptr = None
        if tag_valid is not None:
            tag_valid = np.ascontiguousarray(tag_valid, dtype=np.uint8)
            assert tag_valid.shape == (M, self.N)
            vptr = tag_valid.ctypes.data_as(C.POINTER(C.c_uint8))
        self._L.orc_batch_run(self._h, int(k0), int(n_steps), int(T), _dp(imu), int(M), _ip(tag_step),
                              _dp(tag_pose), _dp(tag_stamp), vptr, float(t_start), int(n_threads))

    def state(self):
        x = np.zeros((16, self.N))
        self._L.orc_batch_get_state(self._h, _dp(x))
        return x

    def cov(self):
        P = np.zeros((self.n * self.n, self.N))
        self._L.orc_batch_get_cov(self._h, _dp(P))
        return P.reshape(self.n, self.n, self.N)

    def aux(self):
        a = np.zeros((11, self.N))
        self._L.orc_batch_get_aux(self._h, _dp(a))
        return a

    def flags(self):
        f = np.zeros((6, self.N), dtype=np.int32)
        self._L.orc_batch_get_flags(self._h, _ip(f))
        return f

    def filter(self, i) -> Filter:
        return Filter(_handle=self._L.orc_batch_filter(self._h, int(i)), _owner=False)

    def counts(self):
        c = (C.c_int64 * 2)()
        self._L.orc_batch_get_counts(self._h, c)
        return int(c[0]), int(c[1])
