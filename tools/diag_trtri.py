"""Diagnostic: locate the origin of a non-reproducible L^-1 under concurrent handles (tile map of the deepest
recursion node of TRTRI that holds a deviating entry)."""
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
nb = (n + 127) // 128

nodes = []
todo = [(0, nb, 0)]
while todo:
    a, b, dep = todo.pop()
    if b - a <= 1:
        continue
    m = a + (b - a + 1) // 2
    nodes.append((a, m, b, dep))
    todo += [(a, m, dep + 1), (m, b, dep + 1)]

with _capi.Engine(X, y, F, kernel="gauss") as e:
    e.objective("LL", th, True)
    ref = e.export("Linv")
engines = [_capi.Engine(X, y, F, kernel="gauss") for _ in range(nthreads)]
lock = threading.Lock()
found = []


def analyse(Li, tag):
    diff = Li != ref
    best = None
    for a, m, b, dep in nodes:
        blk = diff[m * 128:min(b * 128, n), a * 128:m * 128]
        if blk.any() and (best is None or dep > best[3]):
            best = (a, m, b, dep)
    dg = [j for j in range(nb) if diff[j * 128:(j + 1) * 128, j * 128:(j + 1) * 128].any()]
    with lock:
        print(tag, "diag blocks differing:", dg, "deepest node with deviating W21:", best, flush=True)
        if best:
            a, m, b, dep = best
            blk = diff[m * 128:min(b * 128, n), a * 128:m * 128]
            ad = np.abs(Li - ref)[m * 128:min(b * 128, n), a * 128:m * 128]
            rows = np.flatnonzero(blk.any(axis=1))
            cols = np.flatnonzero(blk.any(axis=0))
            print("   node rows %d..%d cols %d..%d ; deviating rows (rel. to node) %d..%d count %d ; cols %d..%d count %d ; max |diff| %.3e (ref max %.3e)"
                  % (m * 128, b * 128, a * 128, m * 128, rows[0], rows[-1], rows.size, cols[0], cols[-1], cols.size,
                     ad.max(), np.abs(ref[m * 128:min(b * 128, n), a * 128:m * 128]).max()))
            # 8 x 8 block map of the first deviating 64 x 128 tile
            r0 = (rows[0] // 64) * 64
            c0 = (cols[0] // 128) * 128
            t = blk[r0:r0 + 64, c0:c0 + 128]
            print("   first deviating tile at node-relative (%d, %d): 8x8-block map (rows = 8-row groups j, cols = 8-col groups):" % (r0, c0))
            for j in range(0, t.shape[0], 8):
                print("     ", "".join("X" if t[j:j + 8, c:c + 8].any() else "." for c in range(0, t.shape[1], 8)))


def worker(t):
    e = engines[t]
    for r in range(reps):
        e.objective("LL", th, True)
        Li = e.export("Linv")
        if not np.array_equal(Li, ref):
            analyse(Li, "thread %d rep %d:" % (t, r))
            found.append(1)


ths = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
for t in ths:
    t.start()
for t in ths:
    t.join()
print("evaluations", nthreads * reps, "with deviating L^-1:", len(found))
