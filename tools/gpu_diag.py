"""Stage-by-stage GPU diagnostic (run under gpurun): compares every exported member of the device model with
numpy/LAPACK for the DMMA/TMA path and for the plain CUDA-core GEMM (LKGPU_DEBUG_SIMPLE_GEMM=1)."""
import json
import os
import sys
import time

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libkriging_b200 import _capi  # noqa: E402
from oracle import kriging_oracle as ko  # noqa: E402
from tests.util import synth  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def check(n, d, kernel, simple, theta=None, full=True):
    os.environ["LKGPU_DEBUG_SIMPLE_GEMM"] = "1" if simple else "0"
    X, y, noise = synth(n, d, 7)
    F = np.ones((n, 1))
    theta = np.full(d, 0.5) if theta is None else theta
    res = dict(n=n, d=d, kernel=kernel, simple=simple)
    with _capi.Engine(X, y, F, kernel=kernel) as e:
        t0 = time.time()
        v, g, info = e.objective("LL", theta, True, with_info=True)
        res["wall_s"] = time.time() - t0
        res["ll"] = v
        res["grad"] = g.tolist()
        res["info"] = info
        if full:
            R = ko.build_R(X, theta, kernel)
            L = sla.cholesky(R, lower=True)
            Linv = sla.solve_triangular(L, np.eye(n), lower=True)
            Rinv = Linv.T @ Linv
            res["err_R"] = rel(e.export("R"), R)
            res["err_L"] = rel(e.export("L"), L)
            res["err_Linv"] = rel(e.export("Linv"), Linv)
            res["err_Rinv"] = rel(e.export("Rinv"), Rinv)
            pb = ko.Problem(X=X, y=y, F=F, kernel=kernel)
            ll, gr = ko.log_likelihood(pb, theta)
            res["ll_ref"] = ll
            res["err_ll"] = abs(v - ll) / abs(ll)
            res["err_grad"] = rel(g, gr)
    print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    out = []
    print("peaks TF/s: dmma", _capi.probe_fp64_peak(0, 0), "dfma", _capi.probe_fp64_peak(0, 1), "mixed",
          _capi.probe_fp64_peak(0, 2), flush=True)
    for simple in (True, False):
        for n in (100, 128, 300, 1000):
            out.append(check(n, 3, "matern5_2", simple))
    for kernel in ("gauss", "exp", "matern3_2"):
        out.append(check(500, 4, kernel, False, theta=np.full(4, 0.4 if kernel != "gauss" else 0.2)))
    for n in (4096, 8192, 20000):
        for rep in range(2):
            out.append(check(n, 10, "matern5_2", False, full=False))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/diag.json", "w"))
