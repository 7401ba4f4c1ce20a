"""oracle/ref.py -- TEST INFRASTRUCTURE.  Python wrapper around
oracle/_ref/ref_driver (the unmodified reference compiled by
oracle/build_ref.sh).  Used by tests/golden/make_golden.py (fixture
generation, build container only) and by bench.py's reference arm /
cpu_baseline (the prebuilt binary travels to the GPU box)."""
from __future__ import annotations

import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "_ref", "ref_driver")


def available() -> bool:
    return os.path.isfile(DRIVER) and os.access(DRIVER, os.X_OK)


def run(X, y, *, kernel="gauss", noise_model="none", noise=None, objective="LL", regmodel="constant",
        normalize=False, mode="eval", optim="none", theta=None, gamma=None, grad=True, reps=1,
        sigma2=None, est_sigma2=None, nugget=None, est_nugget=None, Xn=None, threads=None,
        loovec=False, dump=False, extra_cfg=None, timeout=None, update=None, beta=None):
    """Run the reference on (X, y).  theta: (nt, d) start / evaluation point(s);
    gamma: evaluation point incl. the extra parameter (alpha | sigma2)."""
    if not available():
        raise RuntimeError("oracle/_ref/ref_driver not built (run oracle/build_ref.sh in the build container)")
    X = np.asfortranarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64).ravel()
    n, d = X.shape
    with tempfile.TemporaryDirectory() as wd:
        X.T.ravel().tofile(os.path.join(wd, "X.bin"))  # column-major
        y.tofile(os.path.join(wd, "y.bin"))
        cfg = dict(n=n, d=d, mode=mode, kernel=kernel, noise_model=noise_model, objective=objective,
                   regmodel=regmodel, normalize=int(normalize), optim=optim, grad=int(grad), reps=reps,
                   loovec=int(loovec), dump=int(dump))
        if noise is not None:
            np.ascontiguousarray(noise, dtype=np.float64).tofile(os.path.join(wd, "noise.bin"))
        if theta is not None:
            th = np.atleast_2d(np.asarray(theta, dtype=np.float64))
            cfg["ntheta"] = th.shape[0]
            np.asfortranarray(th).T.ravel().tofile(os.path.join(wd, "theta.bin"))
        if gamma is not None:
            np.ascontiguousarray(gamma, dtype=np.float64).tofile(os.path.join(wd, "gamma.bin"))
        if sigma2 is not None:
            cfg["sigma2"] = repr(float(sigma2)); cfg["est_sigma2"] = int(bool(est_sigma2))
        if nugget is not None:
            cfg["nugget"] = repr(float(nugget)); cfg["est_nugget"] = int(bool(est_nugget))
        if beta is not None:  # fixed trend coefficients (is_beta_estim = false)
            b = np.ascontiguousarray(beta, dtype=np.float64).ravel()
            cfg["beta_n"] = b.size
            b.tofile(os.path.join(wd, "beta.bin"))
        if Xn is not None:
            Xn = np.asfortranarray(Xn, dtype=np.float64)
            cfg["m"] = Xn.shape[0]
            Xn.T.ravel().tofile(os.path.join(wd, "Xn.bin"))
        if update is not None:
            # update = dict(X=..., y=..., refit=bool, noise=...): Kriging::update after the fit
            Xu = np.asfortranarray(update["X"], dtype=np.float64)
            cfg["update_n"] = Xu.shape[0]
            cfg["update_refit"] = int(bool(update.get("refit", False)))
            Xu.T.ravel().tofile(os.path.join(wd, "Xu.bin"))
            np.ascontiguousarray(update["y"], dtype=np.float64).tofile(os.path.join(wd, "yu.bin"))
            if update.get("noise") is not None:
                np.ascontiguousarray(update["noise"], dtype=np.float64).tofile(os.path.join(wd, "noiseu.bin"))
        if extra_cfg:
            cfg.update(extra_cfg)
        with open(os.path.join(wd, "cfg.txt"), "w") as f:
            for k, v in cfg.items():
                f.write(f"{k}={v}\n")
        env = dict(os.environ)
        lp = os.path.join(HERE, "_ref", "ld_library_path.txt")
        if os.path.isfile(lp):
            env["LD_LIBRARY_PATH"] = open(lp).read().strip() + ":" + env.get("LD_LIBRARY_PATH", "")
        if threads is not None:
            env["OPENBLAS_NUM_THREADS"] = str(threads)
            env["OMP_NUM_THREADS"] = str(threads)
        out = subprocess.run([DRIVER, wd], env=env, capture_output=True, text=True, timeout=timeout)
        if out.returncode != 0:
            raise RuntimeError(f"ref_driver failed ({out.returncode}): {out.stderr[-2000:]}")
        res = json.loads(out.stdout.strip().splitlines()[-1])
        if dump:
            n = n + (cfg.get("update_n") or 0)
            res["T"] = np.fromfile(os.path.join(wd, "out_T.bin")).reshape(n, n, order="F")
            res["z"] = np.fromfile(os.path.join(wd, "out_z.bin"))
            M = np.fromfile(os.path.join(wd, "out_M.bin"))
            res["M"] = M.reshape(n, -1, order="F")
        return res
