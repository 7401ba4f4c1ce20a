"""tests/golden/make_golden.py -- fixture generator (run in the BUILD container only).

(1) Converts the reference's own golden CSVs (/root/reference/tests/references/
    data1-scal-*, data2-grad-{1..10}-*; protocol in
    bindings/Python/pylibkriging/tests/binding_consistency_test.py:46-95) into
    tests/golden/reference_vectors.json.
(2) Runs the unmodified reference (oracle/_ref/ref_driver, built by
    oracle/build_ref.sh) on seeded synthetic cases that have no golden file
    (exp / matern kernels, nugget / heterogeneous noise, LOO, LMP, trends,
    jitter retries, fits, predict) and stores inputs' seeds + outputs in
    tests/golden/refgen_vectors.json.

Usage: python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

REFDIR = "/root/reference/tests/references"
OUT = os.path.dirname(os.path.abspath(__file__))


def synth(n, d, seed, yfun="prodsin"):
    """Seeded inputs shared by the generator and the tests (tests/util.py re-implements this)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.random((n, d))
    if yfun == "prodsin":  # reference golden generator: test-binding-consistency.R:86
        y = np.prod(np.sin((X - 0.5) ** 2), axis=1)
    elif yfun == "sumsin":  # bench/bench-kriging.cpp:71-77
        y = np.sum(np.sin(2 * np.pi * X), axis=1)
    elif yfun == "smooth":
        y = np.sin(3.0 * X[:, 0]) + np.sum(X * X, axis=1) + 0.05 * rng.standard_normal(n)
    else:
        raise ValueError(yfun)
    noise = 0.01 + 0.05 * rng.random(n)
    return X, y, noise


def main():
    # ---- (1) reference golden CSVs ----
    sets = {}
    def rd(name):
        return np.genfromtxt(os.path.join(REFDIR, name), delimiter=",")
    X = rd("data1-scal-X.csv").reshape(-1, 1)
    sets["data1-scal"] = dict(
        X=X.tolist(), y=rd("data1-scal-y.csv").ravel().tolist(),
        ll=float(rd("data1-scal-result-logLikelihood.csv")),
        ll_grad=np.atleast_1d(rd("data1-scal-result-logLikelihoodGrad.csv")).tolist(),
        loo=float(rd("data1-scal-result-leaveOneOut.csv")),
        loo_grad=np.atleast_1d(rd("data1-scal-result-leaveOneOutGrad.csv")).tolist())
    for i in range(1, 11):
        p = f"data2-grad-{i}"
        sets[p] = dict(X=rd(p + "-X.csv").tolist(), y=rd(p + "-y.csv").ravel().tolist(),
                       ll=float(rd(p + "-result-logLikelihood.csv")),
                       ll_grad=np.atleast_1d(rd(p + "-result-logLikelihoodGrad.csv")).tolist())
    with open(os.path.join(OUT, "reference_vectors.json"), "w") as f:
        json.dump(dict(source="libKriging tests/references/*.csv", kernel="gauss", theta_value=0.3,
                       regmodel="constant", normalize=False, tolerance=1e-12, sets=sets), f)
    print("reference_vectors.json:", len(sets), "sets")

    # ---- (2) reference-generated vectors ----
    cases = []
    def add(name, **kw):
        cases.append(dict(name=name, **kw))
    kernels = ["gauss", "exp", "matern3_2", "matern5_2"]
    th = {"gauss": 0.35, "exp": 0.8, "matern3_2": 0.6, "matern5_2": 0.5}
    # LL, three noise models, four kernels
    for k in kernels:
        add(f"ll-none-{k}", n=150, d=3, seed=11, kernel=k, noise_model="none", objective="LL", theta=[th[k]] * 3)
        add(f"ll-nugget-{k}", n=120, d=3, seed=12, kernel=k, noise_model="nugget", objective="LL",
            theta=[th[k]] * 3, extra=0.9, yfun="smooth")
        add(f"ll-hetero-{k}", n=120, d=2, seed=13, kernel=k, noise_model="hetero", objective="LL",
            theta=[th[k]] * 2, extra=0.7, yfun="smooth")
        add(f"loo-none-{k}", n=100, d=3, seed=14, kernel=k, noise_model="none", objective="LOO", theta=[th[k]] * 3)
        add(f"lmp-none-{k}", n=100, d=3, seed=15, kernel=k, noise_model="none", objective="LMP", theta=[th[k]] * 3)
    add("lmp-nugget-matern5_2", n=100, d=3, seed=16, kernel="matern5_2", noise_model="nugget", objective="LMP",
        theta=[0.5] * 3, extra=0.85, yfun="smooth")
    # ragged sizes around the 128-tile boundary, anisotropic theta, trends
    for n in (1, 2, 5, 127, 128, 129, 257, 300):
        if n < 3:
            continue
        add(f"ll-size-{n}", n=n, d=4, seed=100 + n, kernel="matern5_2", noise_model="none", objective="LL",
            theta=[0.4, 0.7, 0.55, 0.9])
    add("ll-linear-trend", n=200, d=3, seed=21, kernel="matern3_2", noise_model="none", objective="LL",
        theta=[0.6, 0.5, 0.8], regmodel="linear")
    add("ll-quadratic-trend", n=200, d=2, seed=22, kernel="gauss", noise_model="none", objective="LL",
        theta=[0.05, 0.06], regmodel="quadratic")
    add("loo-linear-trend", n=150, d=2, seed=23, kernel="matern5_2", noise_model="none", objective="LOO",
        theta=[0.5, 0.6], regmodel="linear")
    add("lmp-linear-trend", n=150, d=2, seed=24, kernel="matern5_2", noise_model="none", objective="LMP",
        theta=[0.5, 0.6], regmodel="linear")
    add("ll-d10", n=400, d=10, seed=25, kernel="matern5_2", noise_model="none", objective="LL", theta=[0.5] * 10)
    add("ll-d20-gauss", n=300, d=20, seed=26, kernel="gauss", noise_model="none", objective="LL", theta=[1.2] * 20)
    # ill-conditioned: jitter ladder fires (reported, gated loosely)
    add("ll-jitter-gauss", n=300, d=2, seed=27, kernel="gauss", noise_model="none", objective="LL",
        theta=[1.5, 1.5], ill=True)
    out = []
    for c in cases:
        X, y, noise = synth(c["n"], c["d"], c["seed"], c.get("yfun", "prodsin"))
        gamma = list(c["theta"]) + ([c["extra"]] if c["noise_model"] != "none" else [])
        kw = dict(kernel=c["kernel"], noise_model=c["noise_model"], objective=c["objective"],
                  regmodel=c.get("regmodel", "constant"), theta=np.array(c["theta"]), gamma=np.array(gamma),
                  loovec=(c["objective"] == "LOO"), threads=1)
        if c["noise_model"] == "hetero":
            kw["noise"] = noise
            kw.update(sigma2=c["extra"], est_sigma2=True)  # optim=none needs a positive sigma2 to build diag
        r = ref.run(X, y, **kw)
        c.update(value=r["value"], grad=r["grad"])
        if "loo_mean" in r:
            c.update(loo_mean=r["loo_mean"], loo_sd=r["loo_sd"], sigma2_at_theta=r["sigma2"])
        out.append(c)
        print(c["name"], r["value"], r["grad"][:3])

    # ---- fits + predict ----
    fits = []
    def addfit(name, **kw):
        fits.append(dict(name=name, **kw))
    addfit("fit-ll-m52-n200-d3", n=200, d=3, seed=31, kernel="matern5_2", noise_model="none", objective="LL", optim="BFGS")
    addfit("fit-ll-gauss-n100-d2", n=100, d=2, seed=32, kernel="gauss", noise_model="none", objective="LL", optim="BFGS")
    addfit("fit-ll-exp-n150-d4-ms4", n=150, d=4, seed=33, kernel="exp", noise_model="none", objective="LL", optim="BFGS4")
    addfit("fit-ll-m32-nugget-n150-d3", n=150, d=3, seed=34, kernel="matern3_2", noise_model="nugget", objective="LL",
           optim="BFGS", yfun="smooth")
    addfit("fit-ll-m52-hetero-n120-d2", n=120, d=2, seed=35, kernel="matern5_2", noise_model="hetero", objective="LL",
           optim="BFGS", yfun="smooth")
    addfit("fit-loo-m52-n100-d2", n=100, d=2, seed=36, kernel="matern5_2", noise_model="none", objective="LOO", optim="BFGS")
    addfit("fit-lmp-m52-n100-d2", n=100, d=2, seed=37, kernel="matern5_2", noise_model="none", objective="LMP", optim="BFGS")
    addfit("fit-ll-m52-n300-d5-norm-lin", n=300, d=5, seed=38, kernel="matern5_2", noise_model="none", objective="LL",
           optim="BFGS", normalize=True, regmodel="linear", yfun="smooth")
    addfit("fit-ll-m52-n500-d10-ms8", n=500, d=10, seed=39, kernel="matern5_2", noise_model="none", objective="LL",
           optim="BFGS8", yfun="smooth")
    fout = []
    for c in fits:
        X, y, noise = synth(c["n"], c["d"], c["seed"], c.get("yfun", "prodsin"))
        rng = np.random.Generator(np.random.PCG64(c["seed"] + 1000))
        Xn = rng.random((25, c["d"]))
        kw = dict(kernel=c["kernel"], noise_model=c["noise_model"], objective=c["objective"], mode="fit",
                  optim=c["optim"], regmodel=c.get("regmodel", "constant"), normalize=c.get("normalize", False),
                  Xn=Xn, threads=1)
        if c["noise_model"] == "hetero":
            kw["noise"] = noise
        r = ref.run(X, y, **kw)
        c.update(theta=r["theta"], sigma2=r["sigma2"], nugget=r["nugget"], beta=r["beta"],
                 objective_at_fit=r["objective_at_fit"], pred_mean=r["pred_mean"], pred_sd=r["pred_sd"])
        fout.append(c)
        print(c["name"], r["theta"], r["sigma2"], r["objective_at_fit"])
    with open(os.path.join(OUT, "refgen_vectors.json"), "w") as f:
        json.dump(dict(source="oracle/_ref/ref_driver (unmodified libKriging, OpenBLAS 0.3.15, 1 thread)",
                       evals=out, fits=fout), f)


if __name__ == "__main__":
    main()
