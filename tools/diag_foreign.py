"""Diagnostic: does ONE handle deviate when an unrelated kernel keeps the GPU busy on another stream?
    python tools/diag_foreign.py <mode> [evals]     mode: none | elementwise | dgemm | smem | engine"""
import os
import sys
import threading

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libkriging_b200 import _capi  # noqa: E402
from tests.util import synth  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "none"
evals = int(sys.argv[2]) if len(sys.argv) > 2 else 300
n, d = 5000, 20
X, y, _ = synth(n, d, 505, "smooth")
F = np.ones((n, 1))
thetas = [np.full(d, 1.0) * (1 + 0.1 * k) for k in range(3)]
stop = threading.Event()


def foreign():
    torch.cuda.set_device(0)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        if mode == "elementwise":
            a = torch.ones(1 << 26, dtype=torch.float64, device="cuda")
            while not stop.is_set():
                for _ in range(20):
                    a.mul_(1.0000001)
                st.synchronize()
        elif mode == "dgemm":
            a = torch.randn(2048, 2048, dtype=torch.float64, device="cuda")
            b = torch.randn(2048, 2048, dtype=torch.float64, device="cuda")
            while not stop.is_set():
                for _ in range(10):
                    torch.mm(a, b)
                st.synchronize()
        elif mode == "smem":
            # small convolution-free op with shared memory: a softmax over short rows
            a = torch.randn(1 << 16, 1000, dtype=torch.float32, device="cuda")
            while not stop.is_set():
                for _ in range(10):
                    torch.softmax(a, dim=1)
                st.synchronize()
        elif mode == "engine":
            with _capi.Engine(X, y, F, kernel="gauss") as e2:
                e2.set_concurrent(True)
                k = 0
                while not stop.is_set():
                    e2.objective("LL", thetas[k % 3], True)
                    k += 1


with _capi.Engine(X, y, F, kernel="gauss") as e:
    # DIAG_FLAG=1 (default): flagged handle = launch-chain sweeps, and the gate lets the foreign engine overlap;
    # DIAG_FLAG=0: the default exclusive mode with the wavefront sweep kernel
    e.set_concurrent(os.environ.get("DIAG_FLAG", "1") == "1")
    ref = []
    for th in thetas:
        v, g = e.objective("LL", th, True)
        ref.append((v, g.copy(), e.export("L"), e.export("Linv"), e.export("Rinv"), e.export("x"), e.export("Estar")))
    t = threading.Thread(target=foreign)
    if mode != "none":
        t.start()
    bad = 0
    for i in range(evals):
        k = i % 3
        v, g, inf = e.objective("LL", thetas[k], True, with_info=True)
        if v != ref[k][0] or not np.array_equal(g, ref[k][1]):
            bad += 1
            if bad <= 6:
                print("  info: n_jitter=%d rcond=%.3e reject_info=%d reject_rcond=%d stage_ms=%s" % (
                    inf["n_jitter"], inf["rcond"], inf["reject_info"], inf["reject_rcond"],
                    {k_: round(x, 2) for k_, x in inf["stage_ms"].items()}), flush=True)
                L, Li, Ri, xx, es = e.export("L"), e.export("Linv"), e.export("Rinv"), e.export("x"), e.export("Estar")
                dl = np.argwhere(L != ref[k][2])
                print("  eval %d: value_same=%s max|dg|=%.2e  #L!=%d (first %s, max rel %.1e)  #Linv!=%d  #Rinv!=%d  x!=%d  Estar!=%d" % (
                    i, v == ref[k][0], float(np.max(np.abs(g - ref[k][1]))), len(dl), dl[:2].tolist(),
                    float(np.max(np.abs(L - ref[k][2])) / np.max(np.abs(ref[k][2]))),
                    int(np.count_nonzero(Li != ref[k][3])), int(np.count_nonzero(Ri != ref[k][4])),
                    int(np.count_nonzero(xx != ref[k][5])), int(np.count_nonzero(es != ref[k][6]))), flush=True)
    stop.set()
    if mode != "none":
        t.join()
print(f"foreign={mode}: {evals} evaluations, deviating: {bad}", flush=True)
