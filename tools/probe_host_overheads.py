"""Host-side overheads around the cfg-5 fit (n = 5000, d = 20): handle creation, C++ host single process and sharded
(two processes on one device)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth  # noqa: E402
from libkriging_b200 import _capi  # noqa: E402
from libkriging_b200.host import driver as cpp  # noqa: E402
from libkriging_b200.kriging import Kriging  # noqa: E402

n, d = 5000, 20
X, y = synth(n, d, 123)
F = np.ones((n, 1))
e0 = _capi.Engine(X, y, F, kernel="gauss")
ts = []
es = []
for i in range(7):
    t0 = time.perf_counter()
    es.append(_capi.Engine(X, y, F, kernel="gauss"))
    ts.append(time.perf_counter() - t0)
print("handle creation s:", [round(t, 4) for t in ts], flush=True)
for e in es:
    e.close()
e0.close()
for opt in ("BFGS8",):
    for rep in range(2):
        k = Kriging("gauss", concurrent_starts=8)
        t0 = time.perf_counter()
        k.fit(y, X, "constant", False, opt, "LL")
        print("python host", opt, "fit s", round(time.perf_counter() - t0, 3), "evals", k.fit_log["n_eval"], "LL", k.logLikelihood(), flush=True)
        k.close()
    r = cpp.run(X, y, kernel="gauss", mode="fit", optim=opt, concurrent_starts=8, timeout=600)
    print("cpp host 1 process", opt, {q: r[q] for q in ("fit_s", "cuda_init_s", "n_eval", "objective_at_fit")}, flush=True)
    rs = cpp.run(X, y, kernel="gauss", mode="fit", optim=opt, concurrent_starts=4, world=2, devices=[0, 0], timeout=600)
    for r in rs:
        print("cpp host 2 processes", opt, {q: r[q] for q in ("rank", "fit_s", "cuda_init_s", "n_eval", "local_n_eval", "local_starts", "objective_at_fit")}, flush=True)
