"""Throughput of T handles evaluating concurrently on one GPU (BASELINE cfg 5 shape by default), bit-for-bit
checked against a lone handle.
    python tools/bench_concurrent.py [n] [d] [kernel] [reps] [T1,T2,...]
Prints one JSON line per T: evaluations/s, ms per evaluation (wall / evaluations), mismatches vs the lone handle."""
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libkriging_b200 import _capi  # noqa: E402
from tests.util import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 20
kernel = sys.argv[3] if len(sys.argv) > 3 else "gauss"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
Ts = [int(t) for t in (sys.argv[5] if len(sys.argv) > 5 else "1,2,4,8").split(",")]
FLAG = os.environ.get("DIAG_FLAG", "1") == "1"
X, y, _ = synth(n, d, 505, "smooth")
F = np.ones((n, 1))
th0 = 1.0 if kernel == "gauss" else 0.5
thetas = [np.full(d, th0) * (1 + 0.1 * k) for k in range(3)]

with _capi.Engine(X, y, F, kernel=kernel) as e:
    ref = [e.objective("LL", th, True) for th in thetas]
    ref = [(v, g.copy()) for v, g in ref]

for T in Ts:
    engines = [_capi.Engine(X, y, F, kernel=kernel) for _ in range(T)]
    for e in engines:
        e.set_concurrent(FLAG and T > 1)
        e.objective("LL", thetas[0], True)  # warm-up
    bad = [0]
    dev_ms = [0.0]
    lock = threading.Lock()

    def worker(t):
        e = engines[t]
        b, ms = 0, 0.0
        for r in range(reps):
            for k, th in enumerate(thetas):
                v, g, info = e.objective("LL", th, True, with_info=True)
                ms += info["stage_ms"]["total"]
                if v != ref[k][0] or not np.array_equal(g, ref[k][1]):
                    b += 1
        with lock:
            bad[0] += b
            dev_ms[0] += ms

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    wall = time.perf_counter() - t0
    nev = T * reps * len(thetas)
    print(json.dumps({"n": n, "d": d, "kernel": kernel, "handles": T, "evals": nev, "wall_s": round(wall, 4),
                      "evals_per_s": round(nev / wall, 2), "ms_per_eval_wall": round(1e3 * wall / nev, 3),
                      "ms_per_eval_device_mean": round(dev_ms[0] / nev, 3), "mismatching": bad[0]}), flush=True)
    for e in engines:
        e.close()
