#!/bin/bash
# Reproducers of round 1's stale-operand bug, re-run after the TMA producer rewrite (elected lane, 3D boxes):
# one handle next to foreign kernels (was 95 % deviating with the bug), overlapping handles bit for bit (was 0.5-3 %)
O=gpurun_out/r02c18; mkdir -p $O
for mode in elementwise smem dgemm engine; do
  echo "== diag_foreign $mode" | tee -a $O/repro.log
  timeout 300 python tools/diag_foreign.py $mode 150 2>&1 | tail -3 | tee -a $O/repro.log
done
echo "== diag_concurrent2 5000 8 40" | tee -a $O/repro.log
timeout 600 python tools/diag_concurrent2.py 5000 8 40 2>&1 | tail -3 | tee -a $O/repro.log
echo "== diag_concurrent2 2500 16 40" | tee -a $O/repro.log
timeout 600 python tools/diag_concurrent2.py 2500 16 40 2>&1 | tail -3 | tee -a $O/repro.log
echo "== n = 20000 next to a bandwidth-bound foreign kernel: repeatability" | tee -a $O/repro.log
timeout 300 python - <<'PY' 2>&1 | tee -a $O/repro.log
import threading, numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from bench import synth
from libkriging_b200 import _capi
n, d = 20000, 10
X, y = synth(n, d, 123)
stop = threading.Event()
def foreign():
    torch.cuda.set_device(0)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        a = torch.ones(1 << 26, dtype=torch.float64, device="cuda")
        while not stop.is_set():
            for _ in range(20):
                a.mul_(1.0000001)
            st.synchronize()
with _capi.Engine(X, y, np.ones((n, 1)), kernel="matern5_2") as e:
    ref = e.objective("LL", np.full(d, 0.5), True)
    t = threading.Thread(target=foreign); t.start()
    bad = 0
    for i in range(12):
        v, g = e.objective("LL", np.full(d, 0.5), True)
        bad += int(v != ref[0] or not np.array_equal(g, ref[1]))
    stop.set(); t.join()
    print("n = 20000: evaluations next to the foreign kernel: 12, deviating from the lone evaluation:", bad)
PY
