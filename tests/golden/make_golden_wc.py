"""tests/golden/make_golden_wc.py -- TEST INFRASTRUCTURE (build container only; needs oracle/_ref/ref_driver).

Generates tests/golden/refgen_fits_wc.json: fits of the UNMODIFIED reference started from an explicit,
well-conditioned theta0 (parameters.theta), on data whose optimum is well-conditioned too, so that the whole
L-BFGS-B trajectory stays at rcond_1(L)^2 >= 1e-12.  Only on such paths is north_star's
"fitted theta within 1e-6" a property of the objective implementation: the default random starts of the other
fit fixtures land at numerically singular points (rcond^2 ~ 1e-18) where any two LAPACKs disagree at 1e-5
and the trajectories bifurcate.  The script also records, with the oracle as objective provider, the smallest
rcond seen on the path and the sensitivity of the fitted theta to a 1e-11 relative perturbation of the objective.

    python tests/golden/make_golden_wc.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from tests.util import synth  # noqa: E402

CASES = [
    dict(name="wc-ll-m52-n300-d4", n=300, d=4, seed=51, kernel="matern5_2", noise_model="none", objective="LL", theta0=0.6),
    dict(name="wc-ll-m32-n250-d3", n=250, d=3, seed=52, kernel="matern3_2", noise_model="none", objective="LL", theta0=0.5),
    dict(name="wc-ll-exp-n300-d5", n=300, d=5, seed=53, kernel="exp", noise_model="none", objective="LL", theta0=1.0),
    dict(name="wc-ll-gauss-n200-d6", n=200, d=6, seed=54, kernel="gauss", noise_model="none", objective="LL", theta0=0.25),
    dict(name="wc-ll-m52-nugget-n250-d3", n=250, d=3, seed=55, kernel="matern5_2", noise_model="nugget", objective="LL",
         theta0=0.6),
    dict(name="wc-ll-m52-n1000-d10", n=1000, d=10, seed=56, kernel="matern5_2", noise_model="none", objective="LL",
         theta0=0.8),
    dict(name="wc-lmp-m52-n200-d3", n=200, d=3, seed=57, kernel="matern5_2", noise_model="none", objective="LMP",
         theta0=0.6),
]


def main():
    out = []
    for c in CASES:
        X, y, noise = synth(c["n"], c["d"], c["seed"], "smooth")
        rng = np.random.Generator(np.random.PCG64(c["seed"] + 1000))
        Xn = rng.random((25, c["d"]))
        th0 = np.full((1, c["d"]), c["theta0"])
        r = ref.run(X, y, kernel=c["kernel"], noise_model=c["noise_model"], objective=c["objective"], mode="fit",
                    optim="BFGS", theta=th0, Xn=Xn, threads=1)
        c = dict(c, yfun="smooth", optim="BFGS", theta=r["theta"], sigma2=r["sigma2"], nugget=r["nugget"], beta=r["beta"],
                 objective_at_fit=r["objective_at_fit"], pred_mean=r["pred_mean"], pred_sd=r["pred_sd"])
        # path diagnostics with the oracle as the objective provider
        from libkriging_b200 import kriging
        from tests.oracle_backend import OracleBackend
        rconds = []

        class Diag(OracleBackend):
            eps = 0.0

            def objective(self, name, gamma, want_grad):
                v, g = super().objective(name, gamma, want_grad)
                from oracle import kriging_oracle as ko
                ex = gamma[c["d"]] if c["noise_model"] != "none" else None
                rconds.append(ko.populate_model(self.pb, np.asarray(gamma)[:c["d"]], ex).rcond2)
                if self.eps:
                    v = v * (1.0 + self.eps * np.sin(1e6 * float(np.sum(gamma))))
                return v, g

        thetas = []
        for eps in (0.0, 1e-11):
            Diag.eps = eps
            k = kriging.Kriging(c["kernel"], c["noise_model"], backend_factory=Diag)
            k.fit(y, X, "constant", False, "BFGS", c["objective"], parameters={"theta": th0})
            thetas.append(k.theta())
        c["path_min_rcond2"] = float(min(rconds))
        c["oracle_theta_relerr"] = float(np.max(np.abs(thetas[0] - np.array(c["theta"])) / np.array(c["theta"])))
        c["theta_sensitivity_1e-11"] = float(np.max(np.abs(thetas[1] - thetas[0]) / thetas[0]))
        print(c["name"], r["theta"], "min rcond2 %.2e" % c["path_min_rcond2"], "oracle relerr %.2e" % c["oracle_theta_relerr"],
              "sens %.2e" % c["theta_sensitivity_1e-11"], "n_evals", len(rconds) // 2)
        out.append(c)
    with open(os.path.join(HERE, "refgen_fits_wc.json"), "w") as f:
        json.dump(dict(source="oracle/_ref/ref_driver (unmodified libKriging, OpenBLAS 0.3.15, 1 thread), explicit theta0",
                       fits=out), f)


if __name__ == "__main__":
    main()
