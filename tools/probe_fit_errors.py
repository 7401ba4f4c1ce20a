"""Measured deviations of the device fits from the reference's fits (the fixtures of tests/test_gpu_fit.py)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libkriging_b200.kriging import Kriging  # noqa: E402
from tests.util import GOLDEN, load_refgen, relerr, synth  # noqa: E402

GEN = load_refgen()
WC = json.load(open(os.path.join(GOLDEN, "refgen_fits_wc.json")))["fits"]
print("default random starts (gates: objective 1e-6, theta 5e-3, sigma2 5e-2)")
for c in GEN["fits"]:
    X, y, noise = synth(c["n"], c["d"], c["seed"], c.get("yfun", "prodsin"))
    k = Kriging(c["kernel"], c["noise_model"])
    k.fit(y, X, c.get("regmodel", "constant"), c.get("normalize", False), c["optim"], c["objective"],
          noise=noise if c["noise_model"] == "hetero" else None)
    obj = {"LL": k.logLikelihood, "LOO": k.leaveOneOut, "LMP": k.logMargPost}[c["objective"]]()
    print(f"  {c['name']:34s} objective {relerr(obj, c['objective_at_fit']):.1e}  theta {relerr(k.theta(), c['theta']):.1e}  "
          f"sigma2 {relerr(k.sigma2(), c['sigma2']):.1e}", flush=True)
    k.close()
print("well-conditioned explicit starts (gates: 1e-6)")
for c in WC:
    X, y, _ = synth(c["n"], c["d"], c["seed"], "smooth")
    k = Kriging(c["kernel"], c["noise_model"])
    k.fit(y, X, "constant", False, "BFGS", c["objective"], parameters={"theta": np.full((1, c["d"]), c["theta0"])})
    obj = {"LL": k.logLikelihood, "LMP": k.logMargPost}[c["objective"]]()
    print(f"  {c['name']:34s} objective {relerr(obj, c['objective_at_fit']):.1e}  theta {relerr(k.theta(), c['theta']):.1e}  "
          f"sigma2 {relerr(k.sigma2(), c['sigma2']):.1e}  path_min_rcond2 {c['path_min_rcond2']:.1e}", flush=True)
    k.close()
