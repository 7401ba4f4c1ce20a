"""Launcher for libkriging_b200/host/_build/lkgpu_host_driver (the C++ host: Armadillo API + lbfgsb_cpp loop on the
CPU, every objective evaluation on the GPU through liblkgpu.so)."""
from __future__ import annotations

import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(HERE, "_build", "lkgpu_host_driver")


def available() -> bool:
    return os.path.isfile(DRIVER) and os.access(DRIVER, os.X_OK)


def build() -> bool:
    r = subprocess.run(["bash", os.path.join(HERE, "build_host.sh")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build_host.sh failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return available()


def run(X, y, *, kernel="gauss", noise_model="none", noise=None, objective="LL", regmodel="constant", normalize=False,
        mode="eval", optim="none", theta=None, gamma=None, grad=True, sigma2=None, est_sigma2=None, nugget=None,
        est_nugget=None, Xn=None, device=0, timeout=None, update=None, concurrent_starts=None, beta=None,
        world=1, devices=None, env=None):
    """Run lkgpu::Kriging on (X, y): mode='fit' (optim=BFGS[#]) or 'eval' (objective value / gradient at theta or
    gamma).  Returns the driver's JSON (theta, beta, sigma2, nugget, objective_at_fit, pred_mean, pred_sd, ...).
    world > 1: a sharded fit -- `world` driver processes (one per entry of `devices`, default device r for rank r; the
    same device may be named twice) share the multistart rows over lkgpu::ShardComm; returns the list of their JSONs
    in rank order."""
    if not available():
        raise RuntimeError(f"{DRIVER} is missing: run libkriging_b200/host/build_host.sh in the build container")
    X = np.asfortranarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64).ravel()
    n, d = X.shape
    with tempfile.TemporaryDirectory() as wd:
        X.T.ravel().tofile(os.path.join(wd, "X.bin"))  # column-major
        y.tofile(os.path.join(wd, "y.bin"))
        cfg = dict(n=n, d=d, mode=mode, kernel=kernel, noise_model=noise_model, objective=objective, regmodel=regmodel,
                   normalize=int(normalize), optim=optim, grad=int(grad), device=device)
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
        if concurrent_starts is not None:
            cfg["concurrent_starts"] = int(concurrent_starts)
        if beta is not None:
            b = np.ascontiguousarray(beta, dtype=np.float64).ravel()
            cfg["beta_n"] = b.size
            b.tofile(os.path.join(wd, "beta.bin"))
        if Xn is not None:
            Xn = np.asfortranarray(Xn, dtype=np.float64)
            cfg["m"] = Xn.shape[0]
            Xn.T.ravel().tofile(os.path.join(wd, "Xn.bin"))
        if update is not None:
            # update = dict(X=..., y=..., refit=bool, noise=...): lkgpu::Kriging::update after the fit
            Xu = np.asfortranarray(update["X"], dtype=np.float64)
            cfg["update_n"] = Xu.shape[0]
            cfg["update_refit"] = int(bool(update.get("refit", False)))
            Xu.T.ravel().tofile(os.path.join(wd, "Xu.bin"))
            np.ascontiguousarray(update["y"], dtype=np.float64).tofile(os.path.join(wd, "yu.bin"))
            if update.get("noise") is not None:
                np.ascontiguousarray(update["noise"], dtype=np.float64).tofile(os.path.join(wd, "noiseu.bin"))
        with open(os.path.join(wd, "cfg.txt"), "w") as f:
            for k, v in cfg.items():
                f.write(f"{k}={v}\n")
        if world <= 1:
            penv = dict(os.environ, **(env or {}))
            for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):  # a single process, even under torchrun
                penv.pop(k, None)
            out = subprocess.run([DRIVER, wd], capture_output=True, text=True, timeout=timeout, env=penv)
            return _parse(out.stdout, out.stderr, out.returncode)
        devices = list(devices) if devices is not None else list(range(world))
        port = _free_port()
        procs = []
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        visible = [v for v in visible.split(",") if v] if visible else None
        for r in range(world):
            penv = dict(os.environ, **(env or {}))
            # each process sees its own GPU only (the driver initialises one device instead of all of them)
            penv.update(RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                        MASTER_PORT=str(port), LKGPU_COMM_PORT_OFFSET="0", LKGPU_HOST_DEVICE="0",
                        CUDA_VISIBLE_DEVICES=visible[devices[r]] if visible else str(devices[r]))
            procs.append(subprocess.Popen([DRIVER, wd], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=penv))
        outs = []
        try:
            for pr in procs:
                so, se = pr.communicate(timeout=timeout)
                outs.append((so, se, pr.returncode))
        finally:
            for pr in procs:
                if pr.poll() is None:
                    pr.kill()
        return [_parse(*o) for o in outs]


def _free_port() -> int:
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _parse(stdout, stderr, rc):
    lines = stdout.strip().splitlines()
    if not lines:
        raise RuntimeError(f"lkgpu_host_driver produced no output (rc={rc}): {stderr[-1000:]}")
    res = json.loads(lines[-1])
    if "error" in res:
        raise RuntimeError(res["error"])
    return res
