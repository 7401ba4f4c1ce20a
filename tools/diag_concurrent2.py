"""Diagnostic: are evaluations bitwise reproducible when several handles run concurrently on one GPU?"""
import os
import sys
import threading

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libkriging_b200 import _capi  # noqa: E402
from tests.util import synth  # noqa: E402

n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 5000, 20
nthreads = int(sys.argv[2]) if len(sys.argv) > 2 else 4
FLAG = os.environ.get("DIAG_FLAG", "1") == "1"   # 1: handles flagged by lkgpu_set_concurrent (overlapping evaluations)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
X, y, _ = synth(n, d, 505, "smooth")
F = np.ones((n, 1))
thetas = [np.full(d, 1.0) * (1 + 0.1 * k) for k in range(3)]

with _capi.Engine(X, y, F, kernel="gauss") as e:
    e.set_concurrent(FLAG)  # the reference takes the kernels the concurrent handles take: bit-for-bit comparison
    ref = []
    for th in thetas:
        v, g = e.objective("LL", th, True)
        ref.append((v, g.copy(), e.export("L"), e.export("Linv"), e.export("Rinv"), e.export("x"), e.export("Estar"),
                    e.export("ystar")))
    v2, g2 = e.objective("LL", thetas[0], True)
    print("sequential repeat identical:", v2 == ref[0][0] and np.array_equal(g2, ref[0][1]), flush=True)

engines = [_capi.Engine(X, y, F, kernel="gauss") for _ in range(nthreads)]
for e in engines:
    e.set_concurrent(FLAG)
bad = []
lock = threading.Lock()


def worker(t):
    e = engines[t]
    for r in range(reps):
        for k, th in enumerate(thetas):
            v, g = e.objective("LL", th, True)
            if v != ref[k][0] or not np.array_equal(g, ref[k][1]):
                L, Li, Ri = e.export("L"), e.export("Linv"), e.export("Rinv")
                xx, es, ys = e.export("x"), e.export("Estar"), e.export("ystar")
                print("   vectors: x != %d (first %s)  Estar != %d (first %s)  ystar != %d" % (
                    np.count_nonzero(xx != ref[k][5]), np.flatnonzero(xx != ref[k][5])[:2].tolist(),
                    np.count_nonzero(es != ref[k][6]), np.flatnonzero(es != ref[k][6])[:2].tolist(),
                    np.count_nonzero(ys != ref[k][7])), flush=True)
                with lock:
                    bad.append((t, r, k, v == ref[k][0], float(np.max(np.abs(g - ref[k][1]))),
                                int(np.count_nonzero(L != ref[k][2])), int(np.count_nonzero(Li != ref[k][3])),
                                int(np.count_nonzero(Ri != ref[k][4])),
                                np.argwhere(L != ref[k][2])[:3].tolist(), np.argwhere(Li != ref[k][3])[:3].tolist(),
                                np.argwhere(Ri != ref[k][4])[:3].tolist()))


ths = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
for t in ths:
    t.start()
for t in ths:
    t.join()
print("concurrent evaluations:", nthreads * reps * len(thetas), "mismatching:", len(bad))
for b in bad[:12]:
    print("  thread %d rep %d theta %d value_same=%s max|dg|=%.3e  #L!= %d  #Linv!= %d  #Rinv!= %d  at %s %s %s" % b)
for e in engines:
    e.close()
