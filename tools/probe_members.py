import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from oracle import kriging_oracle as ko
from tests.util import synth, relerr, relerr_vec
from libkriging_b200 import _capi as capi
for kernel,n,d in [("matern5_2", 333, 4), ("gauss", 200, 2), ("exp", 512, 6), ("matern3_2", 129, 3)]:
    X, y, _ = synth(n, d, 5, "smooth")
    F = ko.regression_matrix("linear", X)
    theta = np.full(d, 0.45 if kernel != "gauss" else 0.08)
    pb = ko.Problem(X=X, y=y, F=F, kernel=kernel)
    m = ko.populate_model(pb, theta)
    c = np.linalg.cond(m.R)
    with capi.Engine(X, y, F, kernel=kernel) as e:
        r = e.eval_raw("LL", theta, want_grad=True)
        out = {k: relerr_vec(e.export(k), getattr(m, k)) for k in ("R","L","Rinv","Fstar","ystar","Estar")}
        out["SSE"] = relerr(r["SSEstar"], m.SSEstar); out["beta"] = relerr_vec(r["betahat"], m.betahat)
        out["|Estar|/|ystar|"] = float(np.linalg.norm(m.Estar)/np.linalg.norm(m.ystar))
    print(kernel, n, "cond(R)=%.3g cond*eps=%.2g" % (c, c*2.2e-16), {k: "%.2g" % v for k,v in out.items()}, flush=True)
