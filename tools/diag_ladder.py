"""Diagnostics of the ladder shortcut (GPU):
 (1) per-evaluation trace of the bench fit (n = 20000): rungs, attempts, chol ms per factorisation;
 (2) the C++ host's BFGS6 fit on the (900, 4, 71) case with 1 and 4 starts in flight, shortcut on / off;
 (3) plain vs shortcut ladder along that fit's trajectory: every evaluation of the sequential fit re-evaluated on a
     fresh-hint handle with the plain ladder (n_jitter must agree)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libkriging_b200 import _capi, kriging  # noqa: E402
from tests.util import synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"

if which in ("all", "cpp"):
    from libkriging_b200 import host
    X, y, _ = synth(900, 4, 71, "smooth")
    for env in ({}, {"LKGPU_FULL_LADDER": "1"}):
        os.environ.pop("LKGPU_FULL_LADDER", None)
        os.environ.update(env)
        for k in (1, 4, 4):
            r = host.run(X, y, kernel="matern5_2", mode="fit", optim="BFGS6", concurrent_starts=k)
            print("cpp", env, "concurrent", k, "n_eval", r["n_eval"], "theta0", r["theta"][0], "obj", r["objective_at_fit"], flush=True)
    os.environ.pop("LKGPU_FULL_LADDER", None)

    # Python host, same data: trace of (gamma, n_jitter) per evaluation, then the plain ladder at the same points
    trace = []

    class Tracing(kriging.GpuBackend):
        def objective(self, name, gamma, want_grad):
            v, g = super().objective(name, gamma, want_grad)
            trace.append((np.array(gamma, float).copy(), self.info["n_jitter"], v, bool(want_grad)))
            return v, g

    kk = kriging.Kriging("matern5_2", backend_factory=Tracing, concurrent_starts=1)
    kk.fit(y, X, optim="BFGS6")
    print("python sequential: evals", len(trace), "n_jitter histogram", np.bincount([t[1] for t in trace]).tolist(), flush=True)
    bad = 0
    with _capi.Engine(X, y, np.ones((900, 1)), kernel="matern5_2") as e:
        e.set_ladder_shortcut(False)
        for gam, nj, v, wg in trace:
            v0, _, i0 = e.objective("LL", gam, wg, with_info=True)
            if i0["n_jitter"] != nj or v0 != v:
                bad += 1
                if bad <= 10:
                    print("  MISMATCH theta", gam[:2], "shortcut n_jitter", nj, "plain", i0["n_jitter"], "values", v, v0, flush=True)
    print("evaluations whose shortcut ladder differs from the plain ladder:", bad, "of", len(trace), flush=True)
    kk.close()

if which in ("all", "fit"):
    from bench import gp_draw, synth as bsynth
    n, d = 20000, 10
    X, y = bsynth(n, d, 123)
    yf = gp_draw(_capi, X, y, "matern5_2", 0.5, 0)
    rows = []

    class Tracing2(kriging.GpuBackend):
        def objective(self, name, gamma, want_grad):
            t0 = time.perf_counter()
            v, g = super().objective(name, gamma, want_grad)
            i = self.info
            nf = i["reject_rcond"] + i["reject_info"] + 1
            rows.append((i["n_jitter"], i["reject_rcond"], i["rungs_skipped"], i["stage_ms"]["chol"], i["stage_ms"]["cov"],
                         i["stage_ms"]["rcond"], i["stage_ms"]["total"], 1e3 * (time.perf_counter() - t0)))
            return v, g

    kk = kriging.Kriging("matern5_2", backend_factory=Tracing2)
    t0 = time.perf_counter()
    kk.fit(yf, X, optim="BFGS")
    print("fit wall", time.perf_counter() - t0, "evals", len(rows))
    print("n_jit rej skipped chol_ms cov_ms rcond_ms total_ms host_ms")
    for r in rows:
        print("%3d %3d %3d %9.1f %7.1f %7.1f %9.1f %9.1f" % r)
    kk.close()
