"""tests/golden/make_golden_fullsize.py -- fixture generator (BUILD container only, CPU, tens of minutes).

Runs the UNMODIFIED reference (oracle/_ref/ref_driver) ONCE at the full BASELINE sizes that fit this container's
host memory and stores value + gradient + the measured wall time in tests/golden/refgen_fullsize.json:

  cfg2   Kriging('matern5_2') LL + gradient, n = 20000, d = 10, theta = 0.5   (BASELINE configs[1]; ~50 GB host RAM)
  cfg3   Kriging('exp') LOO + gradient,     n = 10000, d = 6,  theta = 0.8    (BASELINE configs[2])
  cfg5   Kriging('gauss') LL + gradient,    n = 5000,  d = 20, theta = 1.2    (BASELINE configs[4] shape)
  big    cfg-2 shape at n = 18000 (what fits this container; cfg2 itself was generated on the GPU box's host, 16 cores)
  mid    several n = 1500..3000 cases (12..24 panels of 128: look-ahead / outer blocks / TRTRI recursion depth > 1)

Inputs are the seeded generator shared with tests/util.py:synth(n, d, seed, 'smooth') and bench.py:synth (same X,
same y), so that the tests and bench.py regenerate them on the GPU box without reading any file.

cfg 4 (n = 40000) cannot be allocated by the reference (ARMA_32BIT_WORD, 102 GB of dX): see
make_golden_cfg4_oracle.py (numpy oracle, in-place LAPACK).

Usage: python tests/golden/make_golden_fullsize.py [case ...]       (cases: cfg2 cfg3 cfg5 mid; default all)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

OUT = os.environ.get("GOLDEN_OUT") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "refgen_fullsize.json")


def synth(n, d, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.random((n, d))
    y = np.sin(3.0 * X[:, 0]) + np.sum(X * X, axis=1) + 0.05 * rng.standard_normal(n)
    return X, y


CASES = {
    "cfg2": dict(n=20000, d=10, seed=123, kernel="matern5_2", noise_model="none", objective="LL", theta=0.5),
    # the largest cfg-2-shaped case this container's 62 GB hold (the reference peaks at 19.6 x 8 n^2 bytes): 141 panels
    "big-ll-m52-n18000": dict(n=18000, d=10, seed=123, kernel="matern5_2", noise_model="none", objective="LL", theta=0.5),
    "cfg3": dict(n=10000, d=6, seed=123, kernel="exp", noise_model="none", objective="LOO", theta=0.8),
    "cfg5": dict(n=5000, d=20, seed=123, kernel="gauss", noise_model="none", objective="LL", theta=1.2),
    "mid-ll-m52-n3000": dict(n=3000, d=10, seed=124, kernel="matern5_2", noise_model="none", objective="LL", theta=0.5),
    "mid-ll-m32-nugget-n2500": dict(n=2500, d=8, seed=125, kernel="matern3_2", noise_model="nugget", objective="LL",
                                    theta=0.6, extra=0.9),
    "mid-lmp-m52-n2000": dict(n=2000, d=6, seed=126, kernel="matern5_2", noise_model="none", objective="LMP", theta=0.5),
    "mid-loo-exp-n1500": dict(n=1500, d=6, seed=127, kernel="exp", noise_model="none", objective="LOO", theta=0.8),
    "mid-ll-gauss-n2100": dict(n=2100, d=12, seed=128, kernel="gauss", noise_model="none", objective="LL", theta=0.9),
}


def main():
    want = sys.argv[1:] or ["mid", "cfg5", "cfg3", "cfg2"]
    names = []
    for w in want:
        names += [k for k in CASES if k == w or (w == "mid" and k.startswith("mid-"))]
    res = {}
    if os.path.isfile(OUT):
        res = json.load(open(OUT)).get("cases", {})
    threads = len(os.sched_getaffinity(0))
    for name in names:
        c = dict(CASES[name])
        X, y = synth(c["n"], c["d"], c["seed"])
        th = np.full(c["d"], c["theta"])
        gamma = np.concatenate([th, [c["extra"]]]) if c["noise_model"] != "none" else th
        t0 = time.time()
        r = ref.run(X, y, kernel=c["kernel"], noise_model=c["noise_model"], objective=c["objective"], theta=th[None, :],
                    gamma=gamma, grad=True, reps=1, threads=threads, loovec=False)
        c.update(value=r["value"], grad=r["grad"], eval_s=r["eval_s_all"][0], populate_s=r["fit_s"],
                 wall_s=time.time() - t0, threads=threads, y_sum=float(np.sum(y)), X_sum=float(np.sum(X)))
        res[name] = c
        print(name, repr(r["value"]), r["grad"], "eval_s", r["eval_s_all"], flush=True)
        with open(OUT, "w") as f:
            json.dump(dict(source="oracle/_ref/ref_driver (unmodified libKriging, OpenBLAS 0.3.15, %d threads), "
                                  "logLikelihoodFun / leaveOneOutFun / logMargPostFun(theta, grad=true) after "
                                  "fit(optim='none')" % threads,
                           generator="tests/golden/make_golden_fullsize.py", inputs="synth(n, d, seed): PCG64, "
                           "X = U[0,1)^(n x d), y = sin(3 x0) + sum x^2 + 0.05 N(0,1)", cases=res), f, indent=1)


if __name__ == "__main__":
    main()
