"""tests/golden/make_golden_none_trend.py -- fixture generator (BUILD container only): reference fits with
regmodel = "none" (Trend::RegressionModel::None, src/lib/Trend.cpp:39: F has no column, p = 0).
Writes tests/golden/refgen_none_trend.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from tests.util import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refgen_none_trend.json")
CASES = [
    dict(name="none-trend-m52-bfgs", n=150, d=3, seed=91, kernel="matern5_2", optim="BFGS", theta0=0.5),
    dict(name="none-trend-m32-fixed-theta", n=200, d=2, seed=92, kernel="matern3_2", optim="none", theta0=0.4),
]


def main():
    out = []
    for c in CASES:
        X, y, _ = synth(c["n"], c["d"], c["seed"], "smooth")
        Xn = np.random.Generator(np.random.PCG64(c["seed"] + 1000)).random((20, c["d"]))
        r = ref.run(X, y, kernel=c["kernel"], regmodel="none", mode="fit", optim=c["optim"],
                    theta=np.full((1, c["d"]), c["theta0"]), Xn=Xn, threads=1)
        out.append(dict(c, theta=r["theta"], sigma2=r["sigma2"], objective_at_fit=r["objective_at_fit"],
                        pred_mean=r["pred_mean"], pred_sd=r["pred_sd"]))
        print(c["name"], r["theta"], r["sigma2"], r["objective_at_fit"])
    json.dump(dict(source="oracle/_ref/ref_driver (unmodified libKriging), regmodel='none'",
                   generator="tests/golden/make_golden_none_trend.py", cases=out), open(OUT, "w"))


if __name__ == "__main__":
    main()
