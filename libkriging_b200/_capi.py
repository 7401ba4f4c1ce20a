"""ctypes binding of include/lkgpu.h (liblkgpu.so).  This is the stub a maintainer of the
reference's Python binding would add; no torch types cross this boundary.

There is no CPU fallback: loading fails loudly when the CUDA extension has not been built,
and every call fails with LkgpuError when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# LKGPU_LIB selects another build of the same library (experiment variants, tools/validate_relax.sh)
LIB_PATH = os.environ.get("LKGPU_LIB") or os.path.join(HERE, "liblkgpu.so")

KERNELS = {"gauss": 0, "exp": 1, "matern3_2": 2, "matern5_2": 3}
NOISE = {"none": 0, "nugget": 1, "hetero": 2}
OBJECTIVES = {"LL": 0, "LOO": 1, "LMP": 2}
EXPORTS = {"L": 0, "R": 1, "Rinv": 2, "Fstar": 3, "Rstar": 4, "ystar": 5, "Estar": 6, "Linv": 7, "x": 8,
           "loo_err": 9, "loo_s2": 10, "z": 11}
N_STAGES = 12
STAGE_NAMES = ["cov", "chol", "rcond", "solves", "trtri", "lauum", "grad", "extra", "total"]
COUNTER_NAMES = {"reject_info": 9, "reject_rcond": 10, "rungs_skipped": 11}  # spare stage_ms slots used as counters (include/lkgpu.h)

_dp = C.POINTER(C.c_double)


class LkgpuOut(C.Structure):
    _fields_ = [
        ("sum_log_diagL", C.c_double), ("SSEstar", C.c_double), ("rcond", C.c_double),
        ("n_jitter", C.c_int), ("info", C.c_int),
        ("sum_offdiag_xRx", C.c_double), ("sum_offdiag_RinvR", C.c_double), ("sum_x2", C.c_double),
        ("trace_Rinv", C.c_double), ("sum_noise_Rinv", C.c_double), ("sum_noise_x2", C.c_double),
        ("sum_log_diagLX", C.c_double), ("S2", C.c_double), ("loo", C.c_double),
        ("stage_ms", C.c_double * N_STAGES),
        ("betahat", _dp), ("t1", _dp), ("t2", _dp), ("obj_grad", _dp),
    ]


class LkgpuError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise LkgpuError(f"{LIB_PATH} is missing: build it with `python -m libkriging_b200.build` "
                         "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.lkgpu_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_int]
    L.lkgpu_set_numerics.argtypes = [vp, C.c_double, C.c_int, C.c_double, C.c_int]
    L.lkgpu_set_params.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double]
    L.lkgpu_theta_bounds.argtypes = [vp, C.c_double, C.c_double, C.c_int, _dp, _dp]
    L.lkgpu_sigma2_variogram.argtypes = [vp, _dp]
    L.lkgpu_eval.argtypes = [vp, C.c_int, _dp, C.c_double, C.c_int, C.POINTER(LkgpuOut)]
    L.lkgpu_objective_fun.argtypes = [vp, C.c_int, _dp, C.c_int, C.c_int, _dp, _dp, C.POINTER(LkgpuOut)]
    L.lkgpu_export.argtypes = [vp, C.c_int, _dp]
    L.lkgpu_predict.argtypes = [vp, C.c_int, _dp, _dp, _dp, C.c_double, _dp, _dp]
    L.lkgpu_set_data.argtypes = [vp, _dp, _dp, _dp, _dp]
    L.lkgpu_append_data.argtypes = [vp, C.c_int, _dp, _dp, _dp, _dp]
    L.lkgpu_set_concurrent.argtypes = [vp, C.c_int]
    L.lkgpu_set_ladder_shortcut.argtypes = [vp, C.c_int]
    L.lkgpu_set_fixed_beta.argtypes = [vp, _dp]
    L.lkgpu_commit_model.argtypes = [vp]
    L.lkgpu_restore_model.argtypes = [vp]
    L.lkgpu_last_eval_was_update.argtypes = [vp]
    L.lkgpu_last_eval_was_update.restype = C.c_int
    L.lkgpu_destroy.argtypes = [vp]
    L.lkgpu_destroy.restype = None
    L.lkgpu_last_error.restype = C.c_char_p
    L.lkgpu_abi_version.restype = C.c_int
    L.lkgpu_launch_count.argtypes = [vp]
    L.lkgpu_launch_count.restype = C.c_longlong
    L.lkgpu_get_stream.argtypes = [vp]
    L.lkgpu_get_stream.restype = C.c_void_p
    L.lkgpu_probe_fp64_peak.argtypes = [C.c_int, C.c_int, _dp]
    L.lkgpu_mem_info.argtypes = [C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(_dp) if a is not None and a.size > 0 else None


def _as_trend(F, n):
    """n x p trend matrix, column-major; p = 0 (regmodel 'none') is an n x 0 array."""
    F = np.asarray(F, dtype=np.float64)
    if F.size == 0:
        return np.zeros((n, 0), order="F")
    return np.asfortranarray(F.reshape(n, -1))


def _check(rc):
    if rc != 0:
        raise LkgpuError(lib().lkgpu_last_error().decode("utf-8", "replace"))


def probe_fp64_peak(device=0, mode=0) -> float:
    """TFLOP/s of dependent-free FP64 DMMA (mode 0), DFMA (mode 1) or a mix (mode 2) chains on every SM."""
    v = C.c_double(0.0)
    _check(lib().lkgpu_probe_fp64_peak(device, mode, C.byref(v)))
    return v.value


def mem_info(device=0):
    """(free, total) bytes of device memory."""
    f, t = C.c_ulonglong(0), C.c_ulonglong(0)
    _check(lib().lkgpu_mem_info(device, C.byref(f), C.byref(t)))
    return f.value, t.value


class Engine:
    """One lkgpu handle = the device-resident KModel workspace of one (process, start)."""

    def __init__(self, X, y, F, *, kernel="gauss", noise_model="none", noise=None, device=0):
        X = np.asfortranarray(X, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64).ravel()
        F = _as_trend(F, X.shape[0])
        self.n, self.d = X.shape
        self.p = F.shape[1]
        self.kernel, self.noise_model, self.device = kernel, noise_model, device
        nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64).ravel()
        self._h = C.c_void_p()
        _check(lib().lkgpu_create(C.byref(self._h), device, self.n, self.d, self.p, _ptr(X), _ptr(y), _ptr(F), _ptr(nz),
                                  KERNELS[kernel], NOISE[noise_model]))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().lkgpu_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_numerics(self, num_nugget=1e-10, max_inc_choldiag=10, min_rcond=1e-18, chol_rcond_check=True):
        _check(lib().lkgpu_set_numerics(self._h, num_nugget, max_inc_choldiag, min_rcond, int(chol_rcond_check)))

    def set_params(self, est_sigma2=True, sigma2=1.0, est_nugget=True, nugget=0.0, alpha=1.0):
        _check(lib().lkgpu_set_params(self._h, int(est_sigma2), float(sigma2), int(est_nugget), float(nugget),
                                      float(alpha)))

    def set_data(self, X, y, F, noise=None):
        X = np.asfortranarray(X, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64).ravel()
        F = _as_trend(F, X.shape[0])
        nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64).ravel()
        _check(lib().lkgpu_set_data(self._h, _ptr(X), _ptr(y), _ptr(F), _ptr(nz)))

    def append_data(self, X_u, y_u, F_u, noise_u=None):
        """lkgpu_append_data: extend the data set by n_u observations; the factor of the last evaluation is kept
        for a block extension at the same theta (Kriging::update)."""
        X_u = np.asfortranarray(np.asarray(X_u, dtype=np.float64).reshape(-1, self.d))
        n_u = X_u.shape[0]
        y_u = np.ascontiguousarray(y_u, dtype=np.float64).ravel()
        F_u = _as_trend(F_u, n_u)
        if y_u.size != n_u or F_u.shape[1] != self.p:
            raise LkgpuError("append_data: y_u / F_u do not match X_u")
        nz = None if noise_u is None else np.ascontiguousarray(noise_u, dtype=np.float64).ravel()
        _check(lib().lkgpu_append_data(self._h, n_u, _ptr(X_u), _ptr(y_u), _ptr(F_u), _ptr(nz)))
        self.n += n_u

    def set_concurrent(self, flag=True):
        """This handle is one of several evaluating at the same time on its device (lkgpu_set_concurrent)."""
        _check(lib().lkgpu_set_concurrent(self._h, int(bool(flag))))

    def set_ladder_shortcut(self, flag=True):
        """lkgpu_set_ladder_shortcut: enter safe_chol_lower's ladder one rung below the previously accepted one."""
        _check(lib().lkgpu_set_ladder_shortcut(self._h, int(bool(flag))))

    def set_fixed_beta(self, beta=None):
        """lkgpu_set_fixed_beta: trend coefficients fixed by the caller (None: estimated)."""
        b = None if beta is None else np.ascontiguousarray(beta, dtype=np.float64).ravel()
        if b is not None and b.size != self.p:
            raise LkgpuError("set_fixed_beta: beta must have p entries")
        _check(lib().lkgpu_set_fixed_beta(self._h, _ptr(b)))

    def commit_model(self):
        """Snapshot the model of the last evaluation as the committed model (m_T, m_M, m_z, ... of the reference)."""
        _check(lib().lkgpu_commit_model(self._h))

    def restore_model(self):
        """Make the committed model the live one again (device-to-device copy; no-op if nothing ran since)."""
        _check(lib().lkgpu_restore_model(self._h))

    @property
    def last_eval_was_update(self) -> bool:
        return bool(lib().lkgpu_last_eval_was_update(self._h))

    def theta_bounds(self, lower_factor=0.02, upper_factor=10.0, heuristic=True):
        lo = np.empty(self.d)
        up = np.empty(self.d)
        _check(lib().lkgpu_theta_bounds(self._h, lower_factor, upper_factor, int(heuristic), _ptr(lo), _ptr(up)))
        return lo, up

    def sigma2_variogram(self) -> float:
        v = C.c_double(0.0)
        _check(lib().lkgpu_sigma2_variogram(self._h, C.byref(v)))
        return v.value

    def eval_raw(self, objective, theta, extra=1.0, want_grad=True):
        """lkgpu_eval: the raw device reductions (dict)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        out = LkgpuOut()
        beta = np.zeros(self.p)
        t1 = np.zeros(self.d)
        t2 = np.zeros(self.d)
        og = np.zeros(self.d)
        out.betahat, out.t1, out.t2, out.obj_grad = _ptr(beta), _ptr(t1), _ptr(t2), _ptr(og)
        _check(lib().lkgpu_eval(self._h, OBJECTIVES[objective], _ptr(theta), float(extra), int(want_grad), C.byref(out)))
        res = {f: getattr(out, f) for f, _ in LkgpuOut._fields_ if f not in ("stage_ms", "betahat", "t1", "t2", "obj_grad")}
        res.update(betahat=beta, t1=t1, t2=t2, obj_grad=og,
                   stage_ms={k: out.stage_ms[i] for i, k in enumerate(STAGE_NAMES)})
        return res

    def objective(self, objective, gamma, want_grad=True, with_info=False):
        """The reference's logLikelihood / leaveOneOut / logMargPost value (+ gradient) at gamma."""
        gamma = np.ascontiguousarray(gamma, dtype=np.float64).ravel()
        val = C.c_double(0.0)
        grad = np.zeros(gamma.size) if want_grad else None
        out = LkgpuOut()
        _check(lib().lkgpu_objective_fun(self._h, OBJECTIVES[objective], _ptr(gamma), gamma.size, int(want_grad),
                                         C.byref(val), _ptr(grad), C.byref(out)))
        if with_info:
            info = dict(n_jitter=out.n_jitter, rcond=out.rcond, SSEstar=out.SSEstar, sum_log_diagL=out.sum_log_diagL,
                        S2=out.S2, stage_ms={k: out.stage_ms[i] for i, k in enumerate(STAGE_NAMES)},
                        **{k: int(out.stage_ms[i]) for k, i in COUNTER_NAMES.items()})
            return val.value, grad, info
        return val.value, grad

    def export(self, which):
        n, p = self.n, self.p
        shape = {"L": (n, n), "R": (n, n), "Rinv": (n, n), "Linv": (n, n), "Fstar": (n, p), "Rstar": (p, p),
                 "ystar": (n,), "Estar": (n,), "x": (n,), "loo_err": (n,), "loo_s2": (n,), "z": (n,)}[which]
        buf = np.empty(shape, dtype=np.float64, order="F")
        _check(lib().lkgpu_export(self._h, EXPORTS[which], _ptr(buf)))
        return buf

    def predict(self, Xn, Fn, beta, r_on_factor=1.0, want_var=True):
        Xn = np.asfortranarray(Xn, dtype=np.float64)
        m = Xn.shape[0]
        Fn = _as_trend(Fn, m)
        beta = np.ascontiguousarray(beta, dtype=np.float64).ravel()
        mean = np.empty(m)
        var = np.empty(m) if want_var else None
        _check(lib().lkgpu_predict(self._h, m, _ptr(Xn), _ptr(Fn), _ptr(beta), float(r_on_factor), _ptr(mean), _ptr(var)))
        return mean, var

    @property
    def stream_ptr(self) -> int:
        """cudaStream_t of the handle's launch stream (for CUDA-event timing by the caller)."""
        return int(lib().lkgpu_get_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(lib().lkgpu_launch_count(self._h))
