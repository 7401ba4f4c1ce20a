"""tests/golden/make_golden_update.py -- TEST INFRASTRUCTURE (build container only; needs oracle/_ref/ref_driver).

Generates tests/golden/refgen_updates.json: the UNMODIFIED reference fitted on the first n0 observations (explicit
theta0, or optim = none at a given theta), then extended by Kriging::update(y_u, X_u, refit) with n_u further ones
(src/lib/Kriging.cpp:2425-2660: block extension of the Cholesky factor for refit = false, warm restart for
refit = true, a new fit for the Nugget / Heterogeneous refits).  Recorded: theta, sigma2, nugget, beta, the
log-likelihood of the updated model and its predictions at 25 points.

    python tests/golden/make_golden_update.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from tests.util import synth_update  # noqa: E402

# n0 / n_u chosen to hit the engine's layouts: o inside a 128-panel (250), on a panel boundary (256), below one
# panel (100), several new panels (n_u = 300), a single new row.
CASES = [
    dict(name="upd-m52-n250+30-d3", n0=250, n_u=30, d=3, seed=61, kernel="matern5_2", noise_model="none", objective="LL",
         optim="BFGS", theta0=0.6),
    dict(name="upd-m52-n256+5-d3", n0=256, n_u=5, d=3, seed=62, kernel="matern5_2", noise_model="none", objective="LL",
         optim="BFGS", theta0=0.6),
    dict(name="upd-m32-n100+40-d2", n0=100, n_u=40, d=2, seed=63, kernel="matern3_2", noise_model="none", objective="LL",
         optim="BFGS", theta0=0.5),
    dict(name="upd-exp-n200+300-d4", n0=200, n_u=300, d=4, seed=64, kernel="exp", noise_model="none", objective="LL",
         optim="BFGS", theta0=1.0),
    dict(name="upd-gauss-none-n300+1-d5", n0=300, n_u=1, d=5, seed=65, kernel="gauss", noise_model="none", objective="LL",
         optim="none", theta0=0.3),
    dict(name="upd-m52-nugget-n250+20-d3", n0=250, n_u=20, d=3, seed=66, kernel="matern5_2", noise_model="nugget",
         objective="LL", optim="BFGS", theta0=0.6, refits=(False,)),
    dict(name="upd-m52-hetero-n150+20-d3", n0=150, n_u=20, d=3, seed=67, kernel="matern5_2", noise_model="hetero",
         objective="LL", optim="none", theta0=0.6, sigma2=0.5, refits=(False,)),
    dict(name="upd-lmp-m52-n200+30-d3", n0=200, n_u=30, d=3, seed=68, kernel="matern5_2", noise_model="none",
         objective="LMP", optim="BFGS", theta0=0.6),
    # the first 6 appended points duplicate kept ones up to 1e-8: the Schur complement is numerically singular, so
    # the ladder of safe_chol_lower runs on IT (chol_block, LinearAlgebra.cpp:286), not on the whole matrix --
    # a from-scratch factorisation of the same data gives sigma2 = 1771.8 instead of 3534.5.
    dict(name="upd-m52-jitter-n150+9-d3", n0=150, n_u=9, d=3, seed=69, kernel="matern5_2", noise_model="none",
         objective="LL", optim="none", theta0=0.6, dup=6, refits=(False,)),
    # normalize = true (the update re-uses the model's own centre / scale, Kriging.cpp:2472-2476), a linear trend,
    # and the two re-fit branches: Nugget (new fit from scratch, :2443-2468) and Heterogeneous with refit (:2645-2658)
    dict(name="upd-m52-norm-n200+30-d3", n0=200, n_u=30, d=3, seed=71, kernel="matern5_2", noise_model="none",
         objective="LL", optim="BFGS", theta0=0.6, normalize=True),
    dict(name="upd-m32-linear-n150+25-d2", n0=150, n_u=25, d=2, seed=72, kernel="matern3_2", noise_model="none",
         objective="LL", optim="BFGS", theta0=0.5, regmodel="linear"),
    dict(name="upd-m52-nugget-n150+20-d3", n0=150, n_u=20, d=3, seed=73, kernel="matern5_2", noise_model="nugget",
         objective="LL", optim="BFGS", theta0=0.6, refits=(True,)),
    dict(name="upd-m52-hetero-n150+20-d3-bfgs", n0=150, n_u=20, d=3, seed=74, kernel="matern5_2", noise_model="hetero",
         objective="LL", optim="BFGS", theta0=0.6, sigma2=0.5, refits=(True,)),
    dict(name="upd-m52-jitter-n300+140-d3", n0=300, n_u=140, d=3, seed=70, kernel="matern5_2", noise_model="none",
         objective="LL", optim="none", theta0=0.6, dup=3, refits=(False,)),
]


def main():
    out = []
    for c in CASES:
        n = c["n0"] + c["n_u"]
        X, y, noise = synth_update(c)
        rng = np.random.Generator(np.random.PCG64(c["seed"] + 1000))
        Xn = rng.random((25, c["d"]))
        th0 = np.full((1, c["d"]), c["theta0"])
        n0 = c["n0"]
        for refit in c.get("refits", (False, True)):
            kw = {}
            if c["noise_model"] == "hetero":
                kw.update(noise=noise[:n0], sigma2=c["sigma2"], est_sigma2=False)
            upd = dict(X=X[n0:], y=y[n0:], refit=refit, noise=noise[n0:] if c["noise_model"] == "hetero" else None)
            r = ref.run(X[:n0], y[:n0], kernel=c["kernel"], noise_model=c["noise_model"], objective=c["objective"],
                        mode="fit", optim=c["optim"], theta=th0, Xn=Xn, threads=1, update=upd,
                        normalize=c.get("normalize", False), regmodel=c.get("regmodel", "constant"), **kw)
            rec = dict(c, refit=refit, yfun="smooth", theta=r["theta"], sigma2=r["sigma2"], nugget=r["nugget"],
                       beta=r["beta"], LL_at_model=r["LL_at_model"], pred_mean=r["pred_mean"], pred_sd=r["pred_sd"])
            rec.pop("refits", None)
            rec["name"] = c["name"] + ("-refit" if refit else "-norefit")
            print(rec["name"], r["theta"], r["sigma2"], r["LL_at_model"])
            out.append(rec)
    with open(os.path.join(HERE, "refgen_updates.json"), "w") as f:
        json.dump(dict(source="oracle/_ref/ref_driver (unmodified libKriging, OpenBLAS 0.3.15, 1 thread): fit on n0 rows, "
                              "then Kriging::update with n_u rows", updates=out), f, indent=1)


if __name__ == "__main__":
    main()
