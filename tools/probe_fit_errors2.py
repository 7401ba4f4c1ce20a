"""Measured deviations: cfg-1 fit (both hosts) and the C++ host's default-start fits against the reference's fits."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libkriging_b200.host import driver as host  # noqa: E402
from libkriging_b200.kriging import Kriging  # noqa: E402
from tests.golden.make_golden_cfg1 import synth_cfg1  # noqa: E402
from tests.util import GOLDEN, load_refgen, relerr, relerr_vec, synth  # noqa: E402

c = json.load(open(os.path.join(GOLDEN, "refgen_cfg1_fit.json")))
X, y = synth_cfg1()
Xn = np.random.Generator(np.random.PCG64(1123)).random((20, 4))
k = Kriging("gauss")
k.fit(y, X, "constant", False, "BFGS", "LL")
mean, sd = k.predict(Xn, True)
v, g = k.logLikelihoodFun(np.asarray(c["theta"]), True)
print(f"cfg1 python host: LL {relerr(k.logLikelihood(), c['objective_at_fit']):.1e} theta {relerr(k.theta(), c['theta']):.1e} "
      f"sigma2 {relerr(k.sigma2(), c['sigma2']):.1e} pred mean {relerr_vec(mean, c['pred_mean']):.1e} sd {relerr_vec(sd, c['pred_sd']):.1e} "
      f"value at ref theta {relerr(v, c['value_at_theta_fit']):.1e} grad {relerr_vec(g, c['grad_at_theta_fit']):.1e}", flush=True)
k.close()
r = host.run(X, y, kernel="gauss", mode="fit", optim="BFGS", Xn=Xn)
print(f"cfg1 C++ host:    LL {relerr(r['objective_at_fit'], c['objective_at_fit']):.1e} theta {relerr(r['theta'], c['theta']):.1e} "
      f"sigma2 {relerr(r['sigma2'], c['sigma2']):.1e} pred mean {relerr_vec(r['pred_mean'], c['pred_mean']):.1e}", flush=True)
GEN = load_refgen()
print("C++ host, default random starts")
for c in GEN["fits"]:
    if c["noise_model"] == "hetero" and False:
        continue
    X, y, noise = synth(c["n"], c["d"], c["seed"], c.get("yfun", "prodsin"))
    try:
        r = host.run(X, y, kernel=c["kernel"], noise_model=c["noise_model"], noise=noise if c["noise_model"] == "hetero" else None,
                     objective=c["objective"], regmodel=c.get("regmodel", "constant"), normalize=c.get("normalize", False),
                     mode="fit", optim=c["optim"])
        print(f"  {c['name']:34s} objective {relerr(r['objective_at_fit'], c['objective_at_fit']):.1e}  theta {relerr(r['theta'], c['theta']):.1e}  "
              f"sigma2 {relerr(r['sigma2'], c['sigma2']):.1e}", flush=True)
    except Exception as ex:
        print("  ", c["name"], "failed:", str(ex)[:100])
