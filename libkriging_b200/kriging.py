"""Host-side mirror of the reference's `Kriging` class for the fit / objective / predict path.

The reference's host (src/lib/Kriging.cpp) keeps normalisation, the trend basis, bounds, random
starts, the log-theta reparametrisation, the L-BFGS-B loop with restarts, the argmin over starts and
the sigma2 commit on the CPU; every objective evaluation (`fit_ofn`, Kriging.cpp:1980-1987) goes to
the device through the C ABI of include/lkgpu.h.  Same method names, argument meaning and error
behaviour as the reference's Python binding (bindings/Python/.../Kriging_binding.hpp), so the parity
tests read like the reference's own.

L-BFGS-B: the reference drives Lbfgsb.3.0 `setulb` through lbfgsb_cpp (reverse communication,
dependencies/lbfgsb_cpp/include/lbfgsb_cpp/lbfgsb.hpp:199-283).  That dependency is not vendored here;
the same reverse-communication loop is run on SciPy's `setulb` (the same L-BFGS-B 3.0 algorithm),
with the reference's settings m = 10, max_iter = 20, pgtol = 1e-3, factr = 1e10 (LOO: / n^2).
"""
from __future__ import annotations

import math

import numpy as np

from . import optim as _optim
from .trend import regression_model_matrix

NOISE_MODELS = ("none", "nugget", "hetero")
_NOISE_ALIASES = {"none": "none", "nugget": "nugget", "heterogeneous": "hetero", "hetero": "hetero", "noise": "hetero"}


# ----------------------------------------------------------------------------------------------
# objective backend: the device engine (the only one the product constructs)
# ----------------------------------------------------------------------------------------------
class GpuBackend:
    """Thin adaptor over _capi.Engine (liblkgpu.so).  No CPU fallback."""

    def __init__(self, X, y, F, kernel, noise_model, noise, device=0):
        from . import _capi
        self._capi = _capi
        self.engine = _capi.Engine(X, y, F, kernel=kernel, noise_model=noise_model, noise=noise, device=device)
        self.info = {}
        self._scalars_key = None
        self._scalars = None
        # running totals over this handle's objective calls (bench.py reports them for the fit)
        self.stats = dict(evals=0, jitter_rungs=0, device_ms=0.0, reject_info=0, reject_rcond=0, rungs_skipped=0,
                          chol_ms=0.0, rcond_ms=0.0)

    def set_params(self, est_sigma2, sigma2, est_nugget, nugget, alpha):
        self.engine.set_params(est_sigma2, sigma2, est_nugget, nugget, alpha)

    def theta_bounds(self, lower_factor, upper_factor, heuristic):
        return self.engine.theta_bounds(lower_factor, upper_factor, heuristic)

    def sigma2_variogram(self):
        return self.engine.sigma2_variogram()

    def objective(self, name, gamma, want_grad):
        val, grad, info = self.engine.objective(name, gamma, want_grad, with_info=True)
        self.info = info
        self._scalars_key = None  # the device now holds the model at gamma
        st = self.stats
        st["evals"] += 1
        st["jitter_rungs"] += int(info["n_jitter"])
        st["device_ms"] += float(info["stage_ms"]["total"])
        st["reject_info"] += int(info.get("reject_info", 0))
        st["reject_rcond"] += int(info.get("reject_rcond", 0))
        st["rungs_skipped"] += int(info.get("rungs_skipped", 0))
        st["chol_ms"] += float(info["stage_ms"]["chol"])
        st["rcond_ms"] += float(info["stage_ms"]["rcond"])
        return val, grad

    def model_scalars(self, theta, extra):
        """SSEstar and betahat of the model at (theta, extra): one value-only evaluation."""
        key = (tuple(np.asarray(theta, dtype=float).tolist()), float(extra))
        if key != self._scalars_key:
            r = self.engine.eval_raw("LL", theta, extra=extra, want_grad=False)
            self._scalars_key, self._scalars = key, (r["SSEstar"], r["betahat"])
        return self._scalars

    def export(self, which):
        return self.engine.export(which)

    def append_data(self, X_u, y_u, F_u, noise_u=None):
        """Kriging::update, data side: the device keeps the committed factor for a block extension at the same
        theta (lkgpu_append_data)."""
        self.engine.append_data(X_u, y_u, F_u, noise_u)
        self._scalars_key = None

    def set_concurrent(self, flag=True):
        self.engine.set_concurrent(flag)

    def set_fixed_beta(self, beta):
        self.engine.set_fixed_beta(beta)

    def set_ladder_shortcut(self, flag):
        self.engine.set_ladder_shortcut(flag)

    def commit(self):
        """The model of the last evaluation becomes the committed one (Kriging.cpp:2156-2173)."""
        self.engine.commit_model()

    def restore(self):
        """The committed model becomes the live one again (objective calls at other points replaced it)."""
        self.engine.restore_model()
        self._scalars_key = None

    @property
    def used_block_update(self):
        return self.engine.last_eval_was_update

    def max_handles(self):
        """How many handles of this size fit in 70 % of the device memory that is free now (+ this one)."""
        free, _ = self._capi.mem_info(self.engine.device)
        N = (self.engine.n + 127) // 128 * 128
        return 1 + int(0.7 * free // (3 * 8 * N * N + (64 << 20)))

    def predict(self, Xn, Fn, beta, r_on_factor):
        return self.engine.predict(Xn, Fn, beta, r_on_factor)

    def close(self):
        self.engine.close()


def _default_backend_factory(X, y, F, kernel, noise_model, noise, device):
    return GpuBackend(X, y, F, kernel, noise_model, noise, device)


# ----------------------------------------------------------------------------------------------
# L-BFGS-B reverse-communication loop (lbfgsb.hpp:199-283 restated on scipy's setulb)
# ----------------------------------------------------------------------------------------------
class LbfgsbResult:
    __slots__ = ("f_opt", "num_iters", "num_fun", "task", "x")


def lbfgsb_minimize(func, x0, lb, ub, *, m=10, max_iter=20, max_fun=15000, factr=1e10, pgtol=1e-3, maxls=20):
    """Minimise func(x) -> (f, g) within [lb, ub] (nbd = 2 everywhere).  x0 is not modified; returns the
    final x.  Termination mirrors lbfgsb_cpp: stop at NEW_X once isave[29] >= max_iter."""
    import scipy
    from scipy.optimize import _lbfgsb
    from scipy.optimize._lbfgsb_py import status_messages

    # private SciPy API: the C translation of L-BFGS-B 3.0 with integer task codes and the ln_task argument came with
    # SciPy 1.15 (this image has it); older releases take a different setulb signature -- fail loudly, not strangely
    if tuple(int(p) for p in scipy.__version__.split(".")[:2]) < (1, 15):
        raise RuntimeError(f"libkriging_b200 needs SciPy >= 1.15 for its L-BFGS-B loop (found {scipy.__version__}); "
                           "the C++ host (libkriging_b200/host) has no such dependency")
    try:
        from scipy.optimize._lbfgsb_py import HAS_ILP64
    except ImportError:  # pragma: no cover
        HAS_ILP64 = False
    int_dtype = np.int64 if HAS_ILP64 else np.int32
    n = len(x0)
    x = np.clip(np.array(x0, dtype=np.float64), lb, ub)
    low = np.array(lb, dtype=np.float64)
    up = np.array(ub, dtype=np.float64)
    nbd = np.full(n, 2, dtype=int_dtype)
    f = np.array(0.0, dtype=np.float64)
    g = np.zeros(n, dtype=np.float64)
    wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
    iwa = np.zeros(3 * n, dtype=int_dtype)
    task = np.zeros(2, dtype=int_dtype)
    ln_task = np.zeros(2, dtype=int_dtype)
    lsave = np.zeros(4, dtype=int_dtype)
    isave = np.zeros(44, dtype=int_dtype)
    dsave = np.zeros(29, dtype=np.float64)
    n_iter = 0
    n_fun = 0
    while True:
        _lbfgsb.setulb(m, x, low, up, nbd, f, g, factr, pgtol, wa, iwa, task, lsave, isave, dsave, maxls, ln_task)
        if task[0] == 3:  # FG
            fv, gv = func(x.copy())
            n_fun += 1
            f = np.array(fv, dtype=np.float64)
            g = np.array(gv, dtype=np.float64)
        elif task[0] == 1:  # NEW_X
            n_iter += 1
            if n_iter >= max_iter:
                task[0], task[1] = 5, 504
            elif n_fun >= max_fun:
                task[0], task[1] = 5, 502
        else:
            break
    r = LbfgsbResult()
    r.f_opt = float(f)
    r.num_iters = n_iter
    r.num_fun = n_fun
    r.task = "ABNORMAL_TERMINATION_IN_LNSRCH" if task[0] == 8 else status_messages.get(int(task[0]), "?")
    r.x = x.copy()
    return r


# ----------------------------------------------------------------------------------------------
# the reference's Kriging class, fit / objective / predict surface
# ----------------------------------------------------------------------------------------------
class Kriging:
    """Kriging(kernel, noise_model="none").  NuggetKriging == noise_model="nugget",
    NoiseKriging == noise_model="hetero" (reference Kriging.hpp:45-49)."""

    def __init__(self, kernel: str, noise_model: str = "none", *, device: int | None = None, backend_factory=None,
                 concurrent_starts: int | None = None, concurrent_handle: bool = False,
                 ladder_shortcut: bool | None = None):
        if kernel not in ("gauss", "exp", "matern3_2", "matern5_2"):
            raise ValueError(f"Unsupported covariance kernel: {kernel}")
        nm = _NOISE_ALIASES.get(noise_model.lower())
        if nm is None:
            raise ValueError(f"Unsupported noise model: {noise_model}")
        self.m_kernel = kernel
        self.m_noise_model = nm
        self._device = device
        self._backend_factory = backend_factory or _default_backend_factory
        self._backend = None
        self._concurrent_starts = concurrent_starts
        # True: this model is one of several being fitted at the same time on the device (nested.fit_submodels)
        self._concurrent_handle = bool(concurrent_handle)
        # None: the engine's default (on; LKGPU_FULL_LADDER=1 turns it off).  False: safe_chol_lower's plain ladder on
        # every evaluation (lkgpu_set_ladder_shortcut).
        self._ladder_shortcut = ladder_shortcut
        self.config = _optim.OptimConfig.from_env()
        self.m_is_empty = True
        self.fit_log = {}

    # ---- accessors (reference Kriging.hpp:200-274) ----
    def kernel(self): return self.m_kernel
    def theta(self): return self.m_theta.copy()
    def sigma2(self): return self.m_sigma2    # raw members, like the reference's accessors (KrigingImpl.hpp:55-58)
    def nugget(self): return self.m_nugget
    def beta(self): return self.m_beta.copy()
    def is_fitted(self): return not self.m_is_empty
    def T(self): self._need_model(); return self._backend.export("L")
    def M(self): self._need_model(); return self._backend.export("Fstar")
    def z(self): self._need_model(); return self._backend.export("z")
    def circ(self): self._need_model(); return self._backend.export("Rstar")

    def close(self):
        if self._backend is not None:
            self._backend.close()
            self._backend = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- fit ----
    def fit(self, y, X, regmodel="constant", normalize=False, optim="BFGS", objective="LL", parameters=None,
            noise=None, comm=None):
        """Kriging::fit (reference src/lib/Kriging.cpp:1591-2215).  `comm` (optional) is a
        libkriging_b200.parallel.MultistartComm: starts are sharded over its ranks."""
        parameters = dict(parameters or {})
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        y = np.asarray(y, dtype=np.float64).ravel()
        n, d = X.shape
        if y.size != n:
            raise ValueError(f"Dimension of new data should be the same:\n X: ({n}x{d}), y: ({y.size})")
        if self.m_noise_model == "hetero":
            if noise is None:
                raise RuntimeError("fit(y, noise, X, ...) requires a noise vector for NoiseModel::Heterogeneous")
            noise = np.asarray(noise, dtype=np.float64).ravel()
            if noise.size != n:
                raise RuntimeError("noise vector must have the same length as y")
        elif noise is not None:
            raise RuntimeError("fit(y, noise, X, ...) requires NoiseModel::Heterogeneous")
        if objective not in ("LL", "LOO", "LMP"):
            raise ValueError("Unsupported fit objective: " + objective + " (supported here: LL, LOO, LMP)")
        if objective == "LOO" and self.m_noise_model != "none":
            raise ValueError("LOO objective not supported for Nugget/Heterogeneous noise modes")
        if objective == "LMP" and self.m_noise_model == "hetero":
            raise ValueError("LMP objective not supported for Heterogeneous noise mode")
        cfg = self.config
        self.m_objective, self.m_optim, self.m_regmodel = objective, optim, regmodel

        # ---- fit_setup_impl (KrigingImpl.cpp:764-841) ----
        self.m_normalize = bool(normalize)
        if normalize:
            self.m_centerX, self.m_scaleX = X.min(axis=0), X.max(axis=0) - X.min(axis=0)
            self.m_centerY, self.m_scaleY = float(y.min()), float(y.max() - y.min())
        else:
            self.m_centerX, self.m_scaleX = np.zeros(d), np.ones(d)
            self.m_centerY, self.m_scaleY = 0.0, 1.0
        self.m_X = np.asfortranarray((X - self.m_centerX) / self.m_scaleX)
        self.m_y = (y - self.m_centerY) / self.m_scaleY
        self.m_noise = noise  # stored raw even when normalize=True (quirk (iii))
        self.m_F = regression_model_matrix(regmodel, self.m_X)
        p = self.m_F.shape[1]
        if p == 0 and objective != "LL":
            raise NotImplementedError("regmodel='none' (no trend column) is supported for objective='LL' only")
        is_beta_estim = parameters.get("is_beta_estim", True)
        beta = parameters.get("beta")
        self.m_est_beta = True
        self.m_beta = np.zeros(0)
        if (not is_beta_estim) and beta is not None and np.size(beta) > 0:
            self.m_est_beta = False
            self.m_beta = np.asarray(beta, dtype=np.float64).ravel() / (self.m_scaleY if normalize else 1.0)
        theta0 = None
        if parameters.get("theta") is not None:
            theta0 = np.atleast_2d(np.asarray(parameters["theta"], dtype=np.float64))
            if theta0.shape[1] != d and theta0.shape[0] == d:
                theta0 = theta0.T
            if normalize:
                theta0 = theta0 / self.m_scaleX
            if theta0.shape[1] != d:
                raise RuntimeError(f"Dimension of theta should be nx{d} instead of {theta0.shape[0]}x{theta0.shape[1]}")
        scaleY = self.m_scaleY

        self.close()
        dev = self._device if self._device is not None else (comm.device if comm is not None else 0)
        be = self._backend = self._backend_factory(self.m_X, self.m_y, self.m_F, self.m_kernel, self.m_noise_model,
                                                   self.m_noise, dev)
        if self._concurrent_handle and hasattr(be, "set_concurrent"):
            be.set_concurrent(True)
        if self._ladder_shortcut is not None and hasattr(be, "set_ladder_shortcut"):
            be.set_ladder_shortcut(self._ladder_shortcut)
        self.m_sigma2, self.m_nugget, self.m_alpha = 1.0, 0.0, 1.0
        self.m_is_empty = True
        sigma2_p = parameters.get("sigma2")
        nugget_p = parameters.get("nugget")
        is_sigma2_estim = parameters.get("is_sigma2_estim", True)
        is_nugget_estim = parameters.get("is_nugget_estim", True)
        nm = self.m_noise_model

        if optim == "none":
            if theta0 is None:
                raise RuntimeError(f"Theta should be given (1x{d}) matrix, when optim=none")
            self.m_theta = theta0[0].copy()
            self.m_est_theta = False
            sigma2 = -1.0
            self.m_est_sigma2 = bool(is_sigma2_estim)
            if sigma2_p is not None:
                sigma2 = float(sigma2_p) / (scaleY * scaleY if normalize else 1.0)
            else:
                self.m_est_sigma2 = True
            nugget_param = 0.0
            extra = 1.0
            self.m_est_nugget = True
            if nm == "nugget":
                self.m_est_nugget = bool(is_nugget_estim)
                if nugget_p is not None:
                    nugget_param = float(nugget_p) / (scaleY * scaleY if normalize else 1.0)
                if sigma2 > 0 and (sigma2 + nugget_param) > 0:
                    self.m_alpha = sigma2 / (sigma2 + nugget_param)
                else:
                    self.m_alpha = 1.0 - _optim.NUGGET_ALPHA_LOWER
                extra = self.m_alpha
            elif nm == "hetero":
                extra = sigma2 if sigma2 > 0 else self.m_sigma2
            SSE, betahat = be.model_scalars(self.m_theta, extra)
            be.commit()
            self._commit_extra = extra
            self.m_is_empty = False
            if self.m_est_beta:
                self.m_beta = betahat
            if nm == "nugget":
                if self.m_est_sigma2:
                    tv = SSE / n
                    self.m_sigma2 = self.m_alpha * tv
                    self.m_nugget = (1.0 - self.m_alpha) * tv if self.m_est_nugget else nugget_param
                else:
                    self.m_sigma2 = sigma2
                    self.m_nugget = 0.0 if self.m_est_nugget else nugget_param
            elif self.m_est_sigma2:
                self.m_sigma2 = SSE / n
            else:
                self.m_sigma2 = sigma2
            self._push_params()
            return self

        if not optim.startswith("BFGS"):
            raise RuntimeError("Unsupported optim: " + optim + " (supported are: none, BFGS[#])")

        # ---- bounds, starts (Kriging.cpp:1703-1829) ----
        theta_lower, theta_upper = be.theta_bounds(cfg.theta_lower_factor, cfg.theta_upper_factor,
                                                   cfg.variogram_bounds_heuristic)
        rng = _optim.ReferenceRandom(123)
        _, multistart = _optim.parse_method(optim, "BFGS")
        theta0_rand = theta_lower[None, :] + rng.randu_mat(multistart, d) * (theta_upper - theta_lower)[None, :]
        if theta0 is not None:
            multistart = max(multistart, theta0.shape[0])
            theta0 = np.vstack([theta0, theta0_rand])[:multistart]
        else:
            theta0 = theta0_rand
        extra0 = None
        extra_lo, extra_up = 0.0, 1.0
        if nm == "nugget":
            extra_lo, extra_up = _optim.NUGGET_ALPHA_LOWER, 1.0
            if sigma2_p is not None and nugget_p is not None:
                s, nu = float(sigma2_p), float(nugget_p)
                extra0 = np.array([s / (s + nu) if (s > 0 and (s + nu) > 0) else extra_lo + (extra_up - extra_lo) * 0.5])
            else:
                extra0 = extra_lo + (extra_up - extra_lo) * (1.0 - rng.randu_vec(theta0.shape[0]) ** 3.0)
        elif nm == "hetero":
            s2v = self._sigma2_variogram()
            extra_lo = 0.1 * (s2v - float(np.max(noise)))
            extra_up = 10.0 * (s2v - float(np.min(noise)))
            if sigma2_p is not None:
                extra0 = np.array([float(sigma2_p)]) / (scaleY if normalize else 1.0)
            else:
                extra0 = extra_lo + (extra_up - extra_lo) * rng.randu_vec(theta0.shape[0])
        gd = d + (0 if nm == "none" else 1)
        rp = _optim.Reparam(nm, d, cfg.reparametrize)
        lo_full = np.append(theta_lower, extra_lo) if gd > d else theta_lower.copy()
        up_full = np.append(theta_upper, extra_up) if gd > d else theta_upper.copy()
        with np.errstate(divide="ignore", invalid="ignore"):
            gamma_lower, gamma_upper = rp.to(lo_full), rp.to(up_full)

        # ---- estimation flags (Kriging.cpp:1831-1849) ----
        self.m_est_sigma2 = bool(is_sigma2_estim)
        if (not self.m_est_sigma2) and sigma2_p is not None:
            self.m_sigma2 = float(sigma2_p) / (scaleY * scaleY if normalize else 1.0)
        else:
            self.m_est_sigma2 = True
        self.m_est_nugget = True
        if nm == "nugget":
            self.m_est_nugget = bool(is_nugget_estim)
            if (not self.m_est_nugget) and nugget_p is not None:
                self.m_nugget = float(nugget_p) / (scaleY * scaleY if normalize else 1.0)
            else:
                self.m_est_nugget = True
        self._push_params()

        sign = 1.0 if objective == "LOO" else -1.0

        def fit_ofn(gamma, want_grad, be=be):
            v = rp.frm(gamma)
            val, grad = be.objective(objective, v, want_grad)
            if want_grad:
                return sign * val, sign * rp.deriv(v, grad)
            return sign * val, None

        if objective == "LOO":
            pgtol = cfg.gradient_tolerance / (n * n)
            factr = cfg.objective_rel_tolerance / 1e-13 / (n * n)
        else:
            pgtol = cfg.gradient_tolerance
            factr = cfg.objective_rel_tolerance / 1e-13

        # ---- one L-BFGS-B run per start (optimize_worker, Kriging.cpp:1904-2084) ----
        def optimize_worker(start_idx, be=be):
            res = dict(start_index=start_idx, objective_value=math.inf, success=False, n_eval=0)
            try:
                theta_start = theta0[start_idx % multistart].copy()
                full = np.append(theta_start, extra0[start_idx % extra0.size]) if gd > d else theta_start
                gamma_tmp = rp.to(full)
                lo_loc = np.minimum(gamma_tmp, gamma_lower)
                up_loc = np.maximum(gamma_tmp, gamma_upper)
                # (the reference's warm-up populate_Model at theta_start, Kriging.cpp:1943-1948, is skipped: its
                #  result is unconditionally overwritten by the first fit_ofn call at the same point)
                retry = 0
                best_f, best_gamma = math.inf, gamma_tmp.copy()
                counter = [0]

                def fg(x):
                    counter[0] += 1
                    return fit_ofn(x, True, be)

                while retry <= cfg.max_restart:
                    r = lbfgsb_minimize(fg, gamma_tmp, lo_loc, up_loc, max_iter=cfg.max_iteration, pgtol=pgtol,
                                        factr=factr)
                    gamma_tmp = r.x
                    if r.f_opt < best_f:
                        best_f, best_gamma = r.f_opt, gamma_tmp.copy()
                    theta_part = rp.frm(gamma_tmp)[:d]
                    sol_to_lb = float(np.min(np.abs(theta_part - theta_lower)))
                    if (retry < cfg.max_restart) and (r.task.startswith("ABNORMAL_TERMINATION_IN_LNSRCH")
                                                      or r.num_iters <= 2 or sol_to_lb < np.finfo(float).eps
                                                      or r.f_opt > best_f):
                        restart_theta = (theta_start + theta_lower) / (2.0 ** (retry + 1))
                        full = np.append(restart_theta, extra0[start_idx % extra0.size]) if gd > d else restart_theta
                        gamma_tmp = rp.to(full)
                        lo_loc = np.minimum(gamma_tmp, lo_loc)
                        up_loc = np.maximum(gamma_tmp, up_loc)
                        retry += 1
                    else:
                        break
                val, _ = fit_ofn(best_gamma, False, be)  # final evaluation (Kriging.cpp:2044)
                counter[0] += 1
                res.update(objective_value=val, gamma=best_gamma, success=True, n_eval=counter[0], retries=retry)
            except Exception as e:  # one failed start must not kill the fit (Kriging.cpp:2075-2081)
                res.update(success=False, error_message=str(e))
            return res

        # Starts are drawn from a queue: all of them in order without a communicator; with one, the static share
        # s mod G == r when there is at most one start per rank, else tickets from the group's store (parallel.py).
        if comm is None:
            from .parallel import StaticQueue
            queue_ = StaticQueue(range(multistart))
            n_mine = multistart
        else:
            queue_ = comm.start_queue(multistart)
            n_mine = -(-multistart // comm.world)
        ncon = self._concurrency(n_mine, n)
        results = {}

        def drain(b):
            while True:
                s = queue_.next()
                if s is None:
                    return
                results[s] = optimize_worker(s, b)

        if ncon <= 1:
            drain(be)
        else:
            # Batched-occupancy path (BASELINE cfg 5, SURVEY.md §8b "several handles per device on separate
            # streams"): a mid-size factorisation cannot fill 148 SMs (its panel chain is latency-bound), so this
            # rank's starts run concurrently, one engine handle (own workspaces, own streams) and one host thread
            # each.  Every start's trajectory is bit for bit the sequential loop's; the argmin below is still taken in
            # start order.
            from concurrent.futures import ThreadPoolExecutor
            extra_be = []
            try:
                for _ in range(ncon - 1):
                    b = self._backend_factory(self.m_X, self.m_y, self.m_F, self.m_kernel, self.m_noise_model,
                                              self.m_noise, dev)
                    b.set_params(self.m_est_sigma2, self.m_sigma2, self.m_est_nugget, self.m_nugget, self.m_alpha)
                    if self._ladder_shortcut is not None and hasattr(b, "set_ladder_shortcut"):
                        b.set_ladder_shortcut(self._ladder_shortcut)
                    extra_be.append(b)
                with ThreadPoolExecutor(max_workers=ncon) as ex:
                    for f in [ex.submit(drain, b) for b in [be] + extra_be]:
                        f.result()
            finally:
                # statistics of the extra handles are folded into the main one (bench.py reads them there)
                st = getattr(be, "stats", None)
                for b in extra_be:
                    if st is not None and getattr(b, "stats", None):
                        for k_, v_ in b.stats.items():
                            st[k_] += v_
                    b.close()
        my_starts = sorted(results)

        # ---- argmin over successful starts, strict '<' in start order (Kriging.cpp:2097-2114) ----
        if comm is None:
            best_idx, min_ofn = -1, math.inf
            for s in range(multistart):
                r = results[s]
                if r["success"] and r["objective_value"] < min_ofn:
                    min_ofn, best_idx = r["objective_value"], s
            best_gamma = results[best_idx]["gamma"] if best_idx >= 0 else None
            n_eval_total = sum(r["n_eval"] for r in results.values())
        else:
            best_idx, min_ofn, best_gamma, n_eval_total = comm.argmin_exchange(results, multistart, gd)
        if best_idx < 0:
            raise RuntimeError(f"All {multistart} optimization attempts failed")
        self.fit_log = dict(multistart=multistart, best_start=best_idx, objective=min_ofn, n_eval=n_eval_total,
                            theta_lower=theta_lower, theta_upper=theta_upper, theta0=theta0, extra0=extra0,
                            local_starts=my_starts, concurrent_starts=ncon,
                            n_eval_local=sum(r["n_eval"] for r in results.values()))

        # ---- commit (Kriging.cpp:2156-2202).  The model of the best start is rebuilt on this rank's device by
        #      one value-only evaluation at gamma* (bit-reproducible; no n x n matrix crosses NVLink). ----
        v = rp.frm(best_gamma)
        self.m_theta = v[:d].copy()
        self.m_est_theta = True
        extra_param = float(v[d]) if gd > d else 0.0
        commit_extra = extra_param if gd > d else 1.0
        # the model the reference commits is the one its last fit_ofn(best_gamma) built, and _logLikelihood overrides
        # the optimiser's extra parameter there when it is fixed (Kriging.cpp:221-234): m_sigma2 for Heterogeneous with
        # sigma2 fixed, sigma2 / (sigma2 + nugget) for Nugget with both fixed
        if objective == "LL":
            if nm == "hetero" and not self.m_est_sigma2:
                commit_extra = self.m_sigma2
            elif nm == "nugget" and not self.m_est_sigma2 and not self.m_est_nugget:
                commit_extra = self.m_sigma2 / (self.m_sigma2 + self.m_nugget)
        SSE, betahat = be.model_scalars(self.m_theta, commit_extra)
        be.commit()
        self._commit_extra = commit_extra
        self.m_is_empty = False
        if self.m_est_beta:
            self.m_beta = betahat
        if nm == "nugget":
            self.m_alpha = extra_param
            if self.m_est_sigma2:
                if self.m_est_nugget:
                    tv = SSE / n
                    self.m_sigma2 = self.m_alpha * tv
                    if objective == "LMP":
                        self.m_sigma2 = self.m_sigma2 * n / (n - p - 2)
                    self.m_nugget = self.m_sigma2 / self.m_alpha - self.m_sigma2
                else:
                    self.m_sigma2 = self.m_nugget * self.m_alpha / (1.0 - self.m_alpha)
            elif self.m_est_nugget:
                self.m_nugget = self.m_sigma2 * (1.0 - self.m_alpha) / self.m_alpha
        elif nm == "hetero":
            if self.m_est_sigma2:
                self.m_sigma2 = extra_param
        elif self.m_est_sigma2:
            self.m_sigma2 = SSE / n
            if objective == "LMP":
                self.m_sigma2 = SSE / (n - p)
        self._push_params()
        return self

    # ---- update (reference src/lib/Kriging.cpp:2425-2660) ----
    def update(self, y_u, X_u, refit=True, noise_u=None):
        """Kriging::update(y_u, X_u, refit) / update(y_u, noise_u, X_u, refit): add observations to a fitted model.

        * Heterogeneous (noise_u given) and Nugget with refit: a new fit() on the joined, de-normalised data, started
          from the current parameters (Kriging.cpp:2443-2468, 2635-2658).
        * refit (NoiseModel::None): warm restart -- one L-BFGS-B run from the current theta on the extended data
          (:2470-2623).
        * no refit: the model is extended at the current theta (KrigingImpl::update_no_refit_impl,
          KrigingImpl.cpp:576-625).
        In the last two cases the committed factor stays on the device and the first evaluation at the current theta
        is the block extension of LinearAlgebra::update_cholCov / chol_block (lkgpu_append_data)."""
        if self.m_is_empty or self._backend is None:
            raise RuntimeError("Kriging model is not fitted")
        y_u = np.asarray(y_u, dtype=np.float64).ravel()
        X_u = np.asarray(X_u, dtype=np.float64)
        d = self.m_X.shape[1]
        if X_u.ndim == 1:
            X_u = X_u.reshape(-1, d)
        if y_u.size != X_u.shape[0]:
            raise RuntimeError(f"Dimension of new data should be the same:\n X: ({X_u.shape[0]}x{X_u.shape[1]}), "
                               f"y: ({y_u.size})")
        if X_u.shape[1] != d:
            raise RuntimeError(f"Dimension of new data should be the same:\n X: (...x{d}), new X: (...x{X_u.shape[1]})")
        nm = self.m_noise_model
        sY = self.m_scaleY
        y_all = lambda: np.concatenate([self.m_y * sY + self.m_centerY, y_u])          # noqa: E731
        X_all = lambda: np.vstack([self.m_X * self.m_scaleX + self.m_centerX, X_u])    # noqa: E731
        if nm == "hetero":
            if noise_u is None:
                raise RuntimeError("update(y, noise, X) requires NoiseModel::Heterogeneous")
            noise_u = np.asarray(noise_u, dtype=np.float64).ravel()
            if noise_u.size != y_u.size:
                raise RuntimeError("noise_u must have the same length as y_u")
            params = dict(sigma2=self.m_sigma2 * sY * sY, is_sigma2_estim=self.m_est_sigma2,
                          theta=(self.m_theta * self.m_scaleX)[None, :], is_theta_estim=self.m_est_theta,
                          beta=None if self.m_est_beta else self.m_beta * sY, is_beta_estim=self.m_est_beta)
            noise_all = np.concatenate([self.m_noise * sY * sY, noise_u])
            return self.fit(y_all(), X_all(), self.m_regmodel, self.m_normalize, self.m_optim if refit else "none",
                            self.m_objective, params, noise=noise_all)
        if noise_u is not None:
            raise RuntimeError("update(y, noise, X) requires NoiseModel::Heterogeneous")
        if refit and self.m_optim != "none" and nm == "nugget":
            params = {}
            if not (self.m_est_beta and self.m_est_nugget and self.m_est_sigma2 and self.m_est_theta):
                params = dict(sigma2=self.m_sigma2 * sY * sY, is_sigma2_estim=self.m_est_sigma2,
                              theta=(self.m_theta * self.m_scaleX)[None, :], is_theta_estim=self.m_est_theta,
                              nugget=self.m_nugget * sY * sY, is_nugget_estim=self.m_est_nugget)
                if not self.m_est_beta:
                    params.update(beta=self.m_beta * sY, is_beta_estim=False)
            return self.fit(y_all(), X_all(), self.m_regmodel, self.m_normalize, self.m_optim, self.m_objective, params)

        # ---- extend the data with the model's own normalisation; the device keeps the committed factor ----
        be = self._backend
        self._need_model()
        Xn_u = (X_u - self.m_centerX) / self.m_scaleX
        yn_u = (y_u - self.m_centerY) / sY
        F_u = regression_model_matrix(self.m_regmodel, Xn_u)
        be.append_data(Xn_u, yn_u, F_u)
        self.m_X = np.asfortranarray(np.vstack([self.m_X, Xn_u]))
        self.m_y = np.concatenate([self.m_y, yn_u])
        self.m_F = np.vstack([self.m_F, F_u])
        n, p = self.m_F.shape
        extra = self.m_alpha if nm == "nugget" else 1.0

        if not (refit and self.m_optim != "none"):
            # update_no_refit_impl: make_Model(m_theta) -- update_eligible -> block extension
            SSE, betahat = be.model_scalars(self.m_theta, extra)
            be.commit()
            self._commit_extra = extra
            if self.m_est_beta:
                self.m_beta = betahat
            if self.m_est_sigma2:
                self.m_sigma2 = SSE / n
            self._push_params()
            return self

        # ---- warm restart (Kriging.cpp:2470-2623): a single L-BFGS-B run from the current theta ----
        cfg = self.config
        objective = self.m_objective
        theta_lower, theta_upper = be.theta_bounds(cfg.theta_lower_factor, cfg.theta_upper_factor,
                                                   cfg.variogram_bounds_heuristic)
        rp = _optim.Reparam("none", d, cfg.reparametrize)
        with np.errstate(divide="ignore", invalid="ignore"):
            gamma_start = rp.to(self.m_theta.copy())
            gamma_lower = np.minimum(gamma_start, rp.to(theta_lower))
            gamma_upper = np.maximum(gamma_start, rp.to(theta_upper))
        be.model_scalars(self.m_theta, extra)  # the warm-up populate_Model at m_theta (:2544-2547): block extension
        sign = 1.0 if objective == "LOO" else -1.0
        counter = [0]

        def fg(gamma):
            counter[0] += 1
            v = rp.frm(gamma)
            val, grad = be.objective(objective, v, True)
            return sign * val, sign * rp.deriv(v, grad)

        if objective == "LOO":
            pgtol, factr = cfg.gradient_tolerance / (n * n), cfg.objective_rel_tolerance / 1e-13 / (n * n)
        else:
            pgtol, factr = cfg.gradient_tolerance, cfg.objective_rel_tolerance / 1e-13
        r = lbfgsb_minimize(fg, gamma_start, gamma_lower, gamma_upper, max_iter=cfg.max_iteration, pgtol=pgtol,
                            factr=factr)
        self.m_theta = rp.frm(r.x)[:d].copy()
        self.m_est_theta = True
        # (the reference commits whatever model the optimiser's last evaluation left in km; here the model at the
        #  returned point is rebuilt by one value-only evaluation -- the same point unless the line search failed)
        SSE, betahat = be.model_scalars(self.m_theta, extra)
        be.commit()
        self._commit_extra = extra
        if self.m_est_beta:
            self.m_beta = betahat
        if self.m_est_sigma2:
            self.m_sigma2 = SSE / (n - p) if objective == "LMP" else SSE / n
        self.fit_log = dict(multistart=1, best_start=0, objective=r.f_opt, n_eval=counter[0] + 2,
                            theta_lower=theta_lower, theta_upper=theta_upper, warm_restart=True)
        self._push_params()
        return self

    # ---- helpers ----
    def _concurrency(self, n_starts, n):
        """Number of engine handles this process runs with OVERLAPPING evaluations for its multistart rows (the
        batched-occupancy path of BASELINE cfg 5).  A factorisation of n <= 8192 cannot fill 148 SMs -- its panel
        chain is latency-bound -- so by default such fits keep several starts in flight, one handle and one host
        thread each (measured on one B200, n = 5000: 10.2 ms per evaluation alone, 5.1 ms with 4 in flight;
        n = 2500: 4.2 -> 1.2 ms with 8).  Kriging(..., concurrent_starts=K) or LKGPU_CONCURRENT_STARTS=K override;
        the count is limited by the starts this rank owns and by free device memory.  Results do not depend on it:
        overlapping evaluations return the bits of a lone handle, and the argmin is taken in start order."""
        import os
        want = self._concurrent_starts
        if want is None and os.environ.get("LKGPU_CONCURRENT_STARTS"):
            want = int(os.environ["LKGPU_CONCURRENT_STARTS"])
        if want is None:
            want = 8 if n <= 3072 else (4 if n <= 8192 else 1)
        want = max(1, min(int(want), n_starts))
        if want > 1 and hasattr(self._backend, "max_handles"):
            want = max(1, min(want, self._backend.max_handles()))
        return want

    def _push_params(self):
        self._backend.set_params(self.m_est_sigma2, self.m_sigma2, getattr(self, "m_est_nugget", True), self.m_nugget,
                                 self.m_alpha)
        # fixed trend coefficients: the committed z is ystar - M beta (Kriging.cpp:2168-2172)
        if hasattr(self._backend, "set_fixed_beta"):
            self._backend.set_fixed_beta(None if self.m_est_beta else self.m_beta)

    def _need_model(self):
        if self.m_is_empty or self._backend is None:
            raise RuntimeError("Kriging model is not fitted")
        # make sure the live model on the device is the committed one (objective calls may have replaced it)
        self._backend.restore()

    def _sigma2_variogram(self):
        """Heterogeneous sigma2 bounds (Kriging.cpp:1784-1797): half the mean squared increment over the pairs
        whose squared distance is at least the median (all n^2 ordered pairs, diagonal included) -- computed on the
        device by lkgpu_sigma2_variogram (radix select over regenerated pair distances)."""
        return float(self._backend.sigma2_variogram())

    def _gamma_full(self, theta):
        theta = np.asarray(theta, dtype=np.float64).ravel()
        d = self.m_X.shape[1]
        if theta.size == d and self.m_noise_model == "nugget":
            return np.append(theta, self.m_alpha)
        if theta.size == d and self.m_noise_model == "hetero":
            return np.append(theta, self.m_sigma2)
        return theta

    # ---- objective accessors (Kriging.cpp:343-349, 470-476, 650-660) ----
    def logLikelihoodFun(self, theta, return_grad=True):
        val, grad = self._backend.objective("LL", self._gamma_full(theta), bool(return_grad))
        return (val, grad) if return_grad else (val, None)

    def leaveOneOutFun(self, theta, return_grad=True):
        if self.m_noise_model != "none":
            raise ValueError("LOO objective not supported for Nugget/Heterogeneous noise modes")
        val, grad = self._backend.objective("LOO", np.asarray(theta, dtype=np.float64), bool(return_grad))
        return (val, grad) if return_grad else (val, None)

    def leaveOneOutVec(self, theta):
        """(yhat_loo, stdev_loo) at theta: Kriging::leaveOneOutVec (Kriging.cpp:478-484) -- m_y - errorsLOO and
        sqrt(sigma2LOO) * sqrt(m_sigma2), in the model's (normalised) output space like the reference."""
        theta = np.asarray(theta, dtype=np.float64)
        self._backend.objective("LOO", theta, False)
        err = self._backend.export("loo_err")
        s2 = self._backend.export("loo_s2")
        return self.m_y - err, np.sqrt(s2) * math.sqrt(self.m_sigma2)

    def logMargPostFun(self, theta, return_grad=True):
        if self.m_noise_model == "hetero":
            raise ValueError("LMP objective not supported for Heterogeneous noise mode")
        val, grad = self._backend.objective("LMP", self._gamma_full(theta), bool(return_grad))
        return (val, grad) if return_grad else (val, None)

    def logLikelihood(self):
        return self.logLikelihoodFun(self.m_theta, False)[0]

    def leaveOneOut(self):
        return self.leaveOneOutFun(self.m_theta, False)[0]

    def logMargPost(self):
        return self.logMargPostFun(self.m_theta, False)[0]

    # ---- predict mean / stdev (Kriging.cpp:2240-2285 -> KrigingImpl.cpp:145-243) ----
    def predict(self, X_n, return_stdev=True):
        self._need_model()
        X_n = np.asarray(X_n, dtype=np.float64)
        if X_n.ndim == 1:
            X_n = X_n.reshape(-1, self.m_X.shape[1])
        d = self.m_X.shape[1]
        if X_n.shape[1] != d:
            raise RuntimeError(f"Predict locations have wrong dimension: {X_n.shape[1]} instead of {d}")
        n_o, p = self.m_F.shape
        Xn = (X_n - self.m_centerX) / self.m_scaleX
        Fn = regression_model_matrix(self.m_regmodel, Xn)
        lmp_scale = (n_o - p) / (n_o - p - 2.0) if self.m_objective == "LMP" else 1.0
        if self.m_noise_model == "nugget":
            factor, var_scale = self.m_alpha, self.m_sigma2 * lmp_scale / self.m_alpha
        else:
            factor, var_scale = 1.0, self.m_sigma2 * lmp_scale
        mean, var = self._backend.predict(Xn, Fn, self.m_beta, factor)
        mean = self.m_centerY + self.m_scaleY * mean
        if not return_stdev:
            return mean, None
        var = np.where(np.isnan(var) | (var < 0), 0.0, var) * var_scale * self.m_scaleY * self.m_scaleY
        return mean, np.sqrt(var)
