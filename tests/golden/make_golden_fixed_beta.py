"""tests/golden/make_golden_fixed_beta.py -- fixture generator (BUILD container only).

Reference runs with FIXED trend coefficients (Parameters::beta, is_beta_estim = false): the committed
z = ystar - M beta (src/lib/Kriging.cpp:1680-1686, 2168-2172) enters predict's mean, the objective keeps its GLS
estimate.  Writes tests/golden/refgen_fixed_beta.json.   Usage: python tests/golden/make_golden_fixed_beta.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from tests.util import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refgen_fixed_beta.json")

CASES = [
    dict(name="fixedbeta-const-none", n=150, d=3, seed=51, kernel="matern5_2", regmodel="constant", normalize=False,
         optim="none", theta=[0.5, 0.6, 0.7], beta=[0.8]),
    dict(name="fixedbeta-linear-norm", n=200, d=2, seed=52, kernel="matern5_2", regmodel="linear", normalize=True,
         optim="none", theta=[0.4, 0.3], beta=[1.2, 0.5, -0.3]),
    dict(name="fixedbeta-const-bfgs", n=120, d=2, seed=53, kernel="matern3_2", regmodel="constant", normalize=False,
         optim="BFGS", theta=[0.5, 0.5], beta=[1.1]),
]


def main():
    out = []
    for c in CASES:
        X, y, _ = synth(c["n"], c["d"], c["seed"], "smooth")
        Xn = np.random.Generator(np.random.PCG64(c["seed"] + 1000)).random((20, c["d"]))
        r = ref.run(X, y, kernel=c["kernel"], regmodel=c["regmodel"], normalize=c["normalize"], mode="fit",
                    optim=c["optim"], theta=np.array(c["theta"])[None, :], beta=c["beta"], Xn=Xn, threads=1, dump=True)
        c = dict(c, theta_fit=r["theta"], sigma2=r["sigma2"], beta_out=r["beta"], pred_mean=r["pred_mean"],
                 pred_sd=r["pred_sd"], z=r["z"].tolist(), objective_at_fit=r["objective_at_fit"])
        out.append(c)
        print(c["name"], r["theta"], r["sigma2"], r["beta"], r["pred_mean"][:2])
    json.dump(dict(source="oracle/_ref/ref_driver (unmodified libKriging), Parameters{beta, is_beta_estim=false}",
                   generator="tests/golden/make_golden_fixed_beta.py", cases=out), open(OUT, "w"))


if __name__ == "__main__":
    main()
