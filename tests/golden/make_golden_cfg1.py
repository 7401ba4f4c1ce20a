"""tests/golden/make_golden_cfg1.py -- fixture generator (BUILD container only, CPU, ~10 s).

BASELINE configs[0], the reference's own bench shape (bench/bench-kriging.cpp:79-126): Kriging('gauss'), constant trend,
normalize = false, optim = 'BFGS', objective = 'LL', n = 1000, d = 4, y = sum_k sin(2 pi x_k).  The UNMODIFIED reference
(oracle/_ref/ref_driver) runs the FIT; theta, sigma2, beta, the LL at the fitted model, predictions at 20 points and
the wall time go to tests/golden/refgen_cfg1_fit.json.  X is numpy's PCG64(123) uniform sample (tests and bench regenerate
it; Armadillo's randu stream is not reproduced), the start point is the reference's own random draw (seed 123), which
both hosts of this repo reproduce (ReferenceRandom).  The fit ends on numerically singular matrices (sigma2 ~ 1e7, jitter
ladder active on most evaluations) -- the regime BASELINE.md §2 describes for this configuration.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refgen_cfg1_fit.json")


def synth_cfg1(n=1000, d=4, seed=123):
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.random((n, d))
    y = np.sum(np.sin(2.0 * np.pi * X), axis=1)
    return X, y


def main():
    X, y = synth_cfg1()
    Xn = np.random.Generator(np.random.PCG64(1123)).random((20, X.shape[1]))
    threads = len(os.sched_getaffinity(0))
    t0 = time.time()
    r = ref.run(X, y, kernel="gauss", objective="LL", mode="fit", optim="BFGS", Xn=Xn, threads=threads)
    wall = time.time() - t0
    # the objective and its gradient at the fitted theta, as one more fixed-theta evaluation of the reference
    th = np.asarray(r["theta"])
    e = ref.run(X, y, kernel="gauss", objective="LL", mode="eval", theta=th[None, :], gamma=th, grad=True, threads=threads)
    out = dict(source="oracle/_ref/ref_driver (unmodified libKriging): fit(y, X, 'constant', false, 'BFGS', 'LL'), "
                      "then logLikelihoodFun(theta_fit, grad=true)",
               n=1000, d=4, seed=123, kernel="gauss", y="sum_k sin(2 pi x_k)", threads=threads, fit_wall_s=wall,
               fit_s=r.get("fit_s"), theta=r["theta"], sigma2=r["sigma2"], beta=r["beta"],
               objective_at_fit=r.get("objective_at_fit"), LL_at_model=r.get("LL_at_model"),
               pred_mean=r["pred_mean"], pred_sd=r["pred_sd"], value_at_theta_fit=e["value"], grad_at_theta_fit=e["grad"],
               y_sum=float(np.sum(y)), X_sum=float(np.sum(X)))
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: out[k] for k in ("theta", "sigma2", "objective_at_fit", "fit_wall_s", "value_at_theta_fit")}))


if __name__ == "__main__":
    main()
