"""ctypes loader for the host-instantiated product core (tests/host_core/host_core.cu).  TEST ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {64: os.path.join(_HERE, "libhost_core64.so"), 32: os.path.join(_HERE, "libhost_core32.so")}
_SRC = [os.path.join(_HERE, "host_core.cu")] + [
    os.path.join(_HERE, "..", "..", "quadrotor_landing_b200", "csrc", f)
    for f in ("ekf_core.cuh", "ekf_synth.cuh", "ekf_kernels.cuh", "ekf_params.hpp")
]


def _build_one(prec):
    subprocess.run(["nvcc", "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-DHC_ONLY_PREC=%d" % prec,
                    "-Xcompiler", "-fPIC", "-shared", "-o", _LIBS[prec], _SRC[0]], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)


def build(force=False):
    """One library per precision (the instantiation sets are disjoint), compiled in parallel."""
    todo = [pr for pr, lib_ in _LIBS.items()
            if force or not os.path.exists(lib_) or any(os.path.getmtime(s) > os.path.getmtime(lib_) for s in _SRC)]
    if todo:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=2) as ex:
            list(ex.map(_build_one, todo))
    return _LIBS[64]


_libs = {}


def lib(prec=64):
    if prec not in _libs:
        build()
        _libs[prec] = C.CDLL(_LIBS[prec])
    return _libs[prec]


def _dp(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def prediction_step(params, x, P, u, prec=64):
    n = 15 if params.est_bias else 9
    x = _f64(x); P = _f64(P); u = _f64(u)
    xo = np.zeros(16); Po = np.zeros((n, n)); acc = np.zeros(3)
    lib(prec).hc_prediction_step(C.byref(params), int(prec), _dp(x), _dp(P), _dp(u), _dp(xo), _dp(Po), _dp(acc))
    return xo, Po, acc


def correction_step(params, x, P, tag, prec=64):
    n = 15 if params.est_bias else 9
    x = _f64(x); P = _f64(P); tag = _f64(tag)
    xo = np.zeros(16); Po = np.zeros((n, n)); obs = np.zeros(7)
    lib(prec).hc_correction_step(C.byref(params), int(prec), _dp(x), _dp(P), _dp(tag), _dp(xo), _dp(Po), _dp(obs))
    return xo, Po, obs


class HostBatch:
    """N filters advanced by the product's run_filter() on the host (state kept in numpy arrays)."""

    def __init__(self, params, n_filters, prec=64):
        self.p = params
        self.N = int(n_filters)
        self.prec = prec
        self.n = 15 if params.est_bias else 9
        self.np_ = self.n * (self.n + 1) // 2
        self.x = np.zeros((16, self.N)); self.x[9] = 1.0
        self.Ppk = np.zeros((self.np_, self.N))
        ci = [params.r_cov_init, params.v_cov_init, params.ang_cov_init, params.ab_cov_init, params.wb_cov_init]
        e = 0
        for a in range(self.n):          # cov_pert = cov_init (cpp:114), as the product's reset kernel does
            for b in range(a, self.n):
                if a == b:
                    self.Ppk[e] = ci[a // 3]
                e += 1
        self.aux = np.zeros((11, self.N)); self.aux[9] = 1.0
        self.pend = np.zeros((8, self.N))
        self.flags = np.zeros(self.N, dtype=np.int32)
        self.upds = np.zeros(self.N, dtype=np.int32)
        self.pf = None          # per-filter overrides: list of 5 arrays [dim][N] (QEKF_PF_* order)
        self._hist = None

    PF_DIMS = (12, 6, 3, 4, 2)

    def set_filter_params(self, field, values):
        p = self.p
        if self.pf is None:
            N = self.N
            one = lambda v: np.repeat(np.asarray(v, dtype=np.float64)[:, None], N, axis=1)
            self.pf = [one(list(p.Q_a) + list(p.Q_w) + list(p.Q_ab) + list(p.Q_wb)), one(list(p.R_r) + list(p.R_ang)),
                       one(list(p.r_v_cv)), one(list(p.q_vc)), one([p.measurement_delay, p.dyn_measurement_delay_offset])]
        v = _f64(values)
        assert v.shape == (self.PF_DIMS[field], self.N)
        self.pf[field] = v.copy()
        self._hist = None

    def _history(self):
        if self._hist is None:
            geo = np.zeros(2, dtype=np.int32)
            d = _dp(np.ascontiguousarray(self.pf[4][0])) if self.pf is not None else None
            L = lib()
            L.hc_ring_geometry.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_int32)]
            L.hc_ring_geometry(C.byref(self.p), d, self.N, geo.ctypes.data_as(C.POINTER(C.c_int32)))
            dmax, ring_len = int(geo[0]), int(geo[1])
            self._hist = dict(xc=np.zeros((16, self.N)), Pc=np.zeros((self.np_, self.N)),
                              ring=np.zeros((ring_len, 6, self.N)), nh=np.zeros(self.N, dtype=np.int32),
                              hpos=np.zeros(self.N, dtype=np.int32), hlen=np.zeros(self.N, dtype=np.int32),
                              ring_len=ring_len, dmax=dmax)
        return self._hist

    def run(self, k0, n_steps, imu, tag_step, tag_pose, tag_stamp, tag_valid=None, t_start=0.0):
        if self.p.multirate_ekf or self.pf is not None:
            return self._run_ext(k0, n_steps, imu, tag_step, tag_pose, tag_stamp, tag_valid, t_start)
        imu = _f64(imu); tag_pose = _f64(tag_pose); tag_stamp = _f64(tag_stamp)
        tag_step = np.ascontiguousarray(tag_step, dtype=np.int32)
        M = tag_step.shape[0]
        vptr = None
        if tag_valid is not None:
            tag_valid = np.ascontiguousarray(tag_valid, dtype=np.uint8)
            vptr = tag_valid.ctypes.data_as(C.POINTER(C.c_uint8))
        L = lib(self.prec)
        L.hc_run.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_double), C.c_int64,
                             C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                             C.POINTER(C.c_uint8), C.c_double] + [C.POINTER(C.c_double)] * 4 + [C.POINTER(C.c_int32)] * 2
        L.hc_run(C.byref(self.p), int(self.prec), self.N, int(k0), int(n_steps), _dp(imu), M,
                 tag_step.ctypes.data_as(C.POINTER(C.c_int32)), _dp(tag_pose), _dp(tag_stamp), vptr, float(t_start),
                 _dp(self.x), _dp(self.Ppk), _dp(self.aux), _dp(self.pend),
                 self.flags.ctypes.data_as(C.POINTER(C.c_int32)), self.upds.ctypes.data_as(C.POINTER(C.c_int32)))

    def _run_ext(self, k0, n_steps, imu, tag_step, tag_pose, tag_stamp, tag_valid, t_start):
        imu = _f64(imu); tag_pose = _f64(tag_pose); tag_stamp = _f64(tag_stamp)
        tag_step = np.ascontiguousarray(tag_step, dtype=np.int32)
        M = tag_step.shape[0]
        vptr = None
        if tag_valid is not None:
            tag_valid = np.ascontiguousarray(tag_valid, dtype=np.uint8)
            vptr = tag_valid.ctypes.data_as(C.POINTER(C.c_uint8))
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        i32 = lambda a: a.ctypes.data_as(ip)
        L = lib(self.prec)
        L.hc_run_ext.argtypes = ([C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, dp, C.c_int64, ip, dp, dp,
                                  C.POINTER(C.c_uint8), C.c_double] + [dp] * 4 + [ip] * 2 + [dp] * 3 + [ip] * 3 +
                                 [C.c_int32, C.c_int32] + [dp] * 5)
        if self.p.multirate_ekf:
            h = self._history()
            hist = [_dp(h["xc"]), _dp(h["Pc"]), _dp(h["ring"]), i32(h["nh"]), i32(h["hpos"]), i32(h["hlen"]),
                    h["ring_len"], h["dmax"]]
        else:
            hist = [None, None, None, None, None, None, 0, 1]
        pf = [_dp(a) for a in self.pf] if self.pf is not None else [None] * 5
        L.hc_run_ext(C.byref(self.p), int(self.prec), self.N, int(k0), int(n_steps), _dp(imu), M, i32(tag_step),
                     _dp(tag_pose), _dp(tag_stamp), vptr, float(t_start), _dp(self.x), _dp(self.Ppk), _dp(self.aux),
                     _dp(self.pend), i32(self.flags), i32(self.upds), *hist, *pf)

    def history_length(self):
        return self._history()["hlen"].copy() if self.p.multirate_ekf else (self.flags & 1)

    def run_mc(self, scn, noise, k0=0, n_steps=None, stats=None, stride=0, lazy=False):
        """Monte-Carlo replay (shared clean scenario + per-filter noise).  stats: array [32][n_bins][20] or None."""
        n_steps = scn.T - k0 if n_steps is None else n_steps
        imu = _f64(scn.imu_clean); pose = _f64(scn.tag_pose_clean); stamp = _f64(scn.tag_stamp); truth = _f64(scn.truth)
        step = np.ascontiguousarray(scn.tag_step, dtype=np.int32)
        if self.p.multirate_ekf or self.pf is not None:
            return self._run_mc_ext(scn, noise, k0, n_steps, stats, stride, lazy)
        L = lib(self.prec)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.hc_run_mc.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, dp, C.c_int64, ip, dp, dp, dp,
                                C.c_void_p, dp, C.c_int32, C.c_int32, C.c_double, dp, dp, dp, dp, ip, ip]
        sp = _dp(stats) if stats is not None else None
        nb = stats.shape[1] if stats is not None else 0
        L.hc_run_mc(C.byref(self.p), int(self.prec), self.N, int(k0), int(n_steps), _dp(imu), step.shape[0],
                    step.ctypes.data_as(ip), _dp(pose), _dp(stamp), _dp(truth), C.byref(noise), sp, nb, int(stride),
                    float(scn.spec.t_start), _dp(self.x), _dp(self.Ppk), _dp(self.aux), _dp(self.pend),
                    self.flags.ctypes.data_as(ip), self.upds.ctypes.data_as(ip))

    def _run_mc_ext(self, scn, noise, k0, n_steps, stats, stride, lazy):
        """Monte-Carlo replay with delayed fusion / per-filter overrides; lazy: re-synthesise the history inputs
        (run_filter_mrs) instead of reading the ring."""
        imu = _f64(scn.imu_clean); pose = _f64(scn.tag_pose_clean); stamp = _f64(scn.tag_stamp); truth = _f64(scn.truth)
        step = np.ascontiguousarray(scn.tag_step, dtype=np.int32)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        i32 = lambda a: a.ctypes.data_as(ip)
        L = lib(self.prec)
        L.hc_run_mc_ext.argtypes = ([C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, dp, C.c_int64, ip, dp, dp, dp,
                                     C.c_void_p, dp, C.c_int32, C.c_int32, C.c_double] + [dp] * 4 + [ip] * 2 + [dp] * 3 +
                                    [ip] * 3 + [C.c_int32, C.c_int32] + [dp] * 5 + [C.c_int])
        if self.p.multirate_ekf:
            h = self._history()
            hist = [_dp(h["xc"]), _dp(h["Pc"]), _dp(h["ring"]), i32(h["nh"]), i32(h["hpos"]), i32(h["hlen"]),
                    h["ring_len"], h["dmax"]]
        else:
            hist = [None, None, None, None, None, None, 0, 1]
        pf = [_dp(a) for a in self.pf] if self.pf is not None else [None] * 5
        sp = _dp(stats) if stats is not None else None
        nb = stats.shape[1] if stats is not None else 0
        L.hc_run_mc_ext(C.byref(self.p), int(self.prec), self.N, int(k0), int(n_steps), _dp(imu), step.shape[0], i32(step),
                        _dp(pose), _dp(stamp), _dp(truth), C.byref(noise), sp, nb, int(stride), float(scn.spec.t_start),
                        _dp(self.x), _dp(self.Ppk), _dp(self.aux), _dp(self.pend), i32(self.flags), i32(self.upds),
                        *hist, *pf, int(bool(lazy)))

    def state(self):
        return self.x.copy()

    def cov(self):
        n = self.n
        P = np.zeros((n, n, self.N))
        e = 0
        for a in range(n):
            for b in range(a, n):
                P[a, b] = self.Ppk[e]; P[b, a] = self.Ppk[e]; e += 1
        return P


def synthesize(scn, noise, first, count, params=None):
    """Explicit streams of filters [first, first+count) from the product's generator (host-instantiated)."""
    imu = _f64(scn.imu_clean); pose = _f64(scn.tag_pose_clean)
    step = np.ascontiguousarray(scn.tag_step, dtype=np.int32)
    T, M = imu.shape[0], step.shape[0]
    o_imu = np.zeros((T, 6, count)); o_tag = np.zeros((M, 7, count))
    o_val = np.zeros((M, count), dtype=np.uint8); o_bias = np.zeros((6, count))
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L = lib()
    L.hc_synthesize.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, dp, C.c_int64, ip, dp, C.c_int64, C.c_int64, dp, dp,
                                C.POINTER(C.c_uint8), dp]
    L.hc_synthesize(C.byref(params) if params is not None else None, C.byref(noise), T, _dp(imu), M, step.ctypes.data_as(ip), _dp(pose), int(first), int(count),
                    _dp(o_imu), _dp(o_tag), o_val.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(o_bias))
    return dict(imu=o_imu, tag_step=step.copy(), tag_pose=o_tag, tag_stamp=_f64(scn.tag_stamp).copy(), tag_valid=o_val,
                bias=o_bias)


def count_flops(params):
    """(predict, correct) floating-point operations executed by the product's structured code (FMA = 2)."""
    out = np.zeros(2)
    lib().hc_count_flops(C.byref(params), _dp(out))
    return int(out[0]), int(out[1])


# ---- cooperative (three lanes per filter) code, host-instantiated with race tracing (host_coop.cu) ----
_COOP_LIB = os.path.join(_HERE, "libhost_coop.so")
_COOP_SRC = [os.path.join(_HERE, "host_coop.cu")] + [
    os.path.join(_HERE, "..", "..", "quadrotor_landing_b200", "csrc", f)
    for f in ("ekf_core.cuh", "ekf_synth.cuh", "ekf_kernels.cuh", "ekf_coop.cuh", "ekf_duo.cuh", "ekf_params.hpp")
]
_coop = None


def build_coop(force=False):
    if force or not os.path.exists(_COOP_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_COOP_LIB) for s in _COOP_SRC):
        res = subprocess.run(["nvcc", "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler",
                              "-fPIC,-pthread", "-shared", "-o", _COOP_LIB, _COOP_SRC[0]],
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise RuntimeError("host_coop build failed:\n" + res.stdout)
    return _COOP_LIB


def coop_lib():
    global _coop
    if _coop is None:
        build_coop()
        _coop = C.CDLL(_COOP_LIB)
    return _coop


def coop_prediction_step(params, x, P, u, order=0):
    """prediction_step by the three-lane code; lanes run phase by phase in the given order (0..5).
    Returns x, P, accel and (races seen by the tracer, spread between the lanes' replicas of q/ab/wb)."""
    n = 15 if params.est_bias else 9
    x = _f64(x); P = _f64(P); u = _f64(u)
    xo = np.zeros(16); Po = np.zeros((n, n)); acc = np.zeros(3); diag = np.zeros(2)
    coop_lib().hcoop_prediction_step(C.byref(params), int(order), _dp(x), _dp(P), _dp(u), _dp(xo), _dp(Po), _dp(acc), _dp(diag))
    return xo, Po, acc, diag


def coop_correction_step(params, x, P, tag, order=0):
    n = 15 if params.est_bias else 9
    x = _f64(x); P = _f64(P); tag = _f64(tag)
    xo = np.zeros(16); Po = np.zeros((n, n)); obs = np.zeros(7); diag = np.zeros(2)
    coop_lib().hcoop_correction_step(C.byref(params), int(order), _dp(x), _dp(P), _dp(tag), _dp(xo), _dp(Po), _dp(obs), _dp(diag))
    return xo, Po, obs, diag


class CoopHostBatch(HostBatch):
    """N filters advanced by the product's run_filter_coop() on the host: three threads per filter meeting at a barrier,
    every shared word traced.  FP64, single-rate.  self.races = races seen by the tracer in the last run."""

    ENTRY = "hcoop_run"

    def __init__(self, params, n_filters):
        assert not params.multirate_ekf
        super().__init__(params, n_filters, 64)
        self.races = 0

    def _call(self, k0, n_steps, imu, step, pose, stamp, valid, t_start, noise, truth, stats, stride):
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L = coop_lib()
        fn = getattr(L, self.ENTRY)
        fn.argtypes = ([C.c_void_p, C.c_int64, C.c_int64, C.c_int64, dp, C.c_int64, ip, dp, dp, C.POINTER(C.c_uint8),
                                 C.c_double] + [dp] * 4 + [ip] * 2 + [C.c_void_p, dp, dp, C.c_int32, C.c_int32] + [dp] * 6)
        vptr = valid.ctypes.data_as(C.POINTER(C.c_uint8)) if valid is not None else None
        pf = [_dp(a) for a in self.pf] if self.pf is not None else [None] * 5
        diag = np.zeros(2)
        fn(C.byref(self.p), self.N, int(k0), int(n_steps), _dp(imu), step.shape[0], step.ctypes.data_as(ip),
                    _dp(pose), _dp(stamp), vptr, float(t_start), _dp(self.x), _dp(self.Ppk), _dp(self.aux), _dp(self.pend),
                    self.flags.ctypes.data_as(ip), self.upds.ctypes.data_as(ip),
                    C.byref(noise) if noise is not None else None, _dp(truth) if truth is not None else None,
                    _dp(stats) if stats is not None else None, stats.shape[1] if stats is not None else 0, int(stride),
                    *pf, _dp(diag))
        self.races = int(diag[0])

    def run(self, k0, n_steps, imu, tag_step, tag_pose, tag_stamp, tag_valid=None, t_start=0.0):
        imu = _f64(imu); tag_pose = _f64(tag_pose); tag_stamp = _f64(tag_stamp)
        tag_step = np.ascontiguousarray(tag_step, dtype=np.int32)
        if tag_valid is not None:
            tag_valid = np.ascontiguousarray(tag_valid, dtype=np.uint8)
        self._call(k0, n_steps, imu, tag_step, tag_pose, tag_stamp, tag_valid, t_start, None, None, None, 0)

    def run_mc(self, scn, noise, k0=0, n_steps=None, stats=None, stride=0):
        n_steps = scn.T - k0 if n_steps is None else n_steps
        imu = _f64(scn.imu_clean); pose = _f64(scn.tag_pose_clean); stamp = _f64(scn.tag_stamp); truth = _f64(scn.truth)
        step = np.ascontiguousarray(scn.tag_step, dtype=np.int32)
        self._call(k0, n_steps, imu, step, pose, stamp, None, scn.spec.t_start, noise, truth, stats, stride)


class DuoHostBatch(CoopHostBatch):
    """N filters advanced by the product's run_filter_duo() on the host: two threads per filter (role A: covariance core
    and corrections, role B: nominal state and bias columns) meeting at a barrier.  FP64, single-rate."""
    ENTRY = "hduo_run"
