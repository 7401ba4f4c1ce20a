"""Diagnostic: origin block of a non-reproducible triangular sweep under concurrent handles."""
import os
import sys
import threading

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libkriging_b200 import _capi  # noqa: E402
from tests.util import synth  # noqa: E402

n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 5000, 20
nthreads = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
X, y, _ = synth(n, d, 505, "smooth")
F = np.ones((n, 1))
th = np.full(d, 1.0)
with _capi.Engine(X, y, F, kernel="gauss") as e:
    e.objective("LL", th, True)
    ref = {k: e.export(k) for k in ("x", "Estar", "ystar", "Fstar")}
engines = [_capi.Engine(X, y, F, kernel="gauss") for _ in range(nthreads)]
lock = threading.Lock()
cnt = [0]


def worker(t):
    e = engines[t]
    for r in range(reps):
        e.objective("LL", th, True)
        for k, fwd in (("ystar", True), ("Fstar", True), ("Estar", True), ("x", False)):
            v = e.export(k).ravel()
            bad = np.flatnonzero(v != ref[k].ravel())
            if bad.size:
                org = bad[0] if fwd else bad[-1]
                blk = org // 128
                inblk = bad[(bad >= blk * 128) & (bad < (blk + 1) * 128)] - blk * 128
                with lock:
                    cnt[0] += 1
                    print("thread %d rep %d %s (%s): %d entries differ; origin block %d; rows in that block: %d (%d..%d) ; |diff| at origin %.3e (value %.3e)"
                          % (t, r, k, "fwd" if fwd else "bwd", bad.size, blk, inblk.size, inblk[0], inblk[-1],
                             abs(v[org] - ref[k].ravel()[org]), ref[k].ravel()[org]), flush=True)
                    print("      in-block pattern (16-row groups):", "".join(
                        "X" if ((inblk >= g) & (inblk < g + 16)).any() else "." for g in range(0, 128, 16)),
                        " rel.diff per 16-row group: " + " ".join("%.0e" % (np.max(np.abs(v[blk * 128 + g:blk * 128 + g + 16] - ref[k].ravel()[blk * 128 + g:blk * 128 + g + 16])) / (np.max(np.abs(ref[k].ravel()[blk * 128:blk * 128 + 128])) + 1e-300)) for g in range(0, min(128, v.size - blk * 128), 16)), flush=True)


ths = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
for t in ths:
    t.start()
for t in ths:
    t.join()
print("evaluations", nthreads * reps, "deviating vectors:", cnt[0])
