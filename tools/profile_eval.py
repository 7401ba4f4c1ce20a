"""Minimal driver for ncu: N evaluations of the bench workload through the C ABI (no torch, no probes).
    python tools/profile_eval.py [n] [d] [evals] [objective] [kernel]
Prints the number of kernels each evaluation launched (to locate launches in the ncu list)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth  # noqa: E402
from libkriging_b200 import _capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 10
evals = int(sys.argv[3]) if len(sys.argv) > 3 else 2
objective = sys.argv[4] if len(sys.argv) > 4 else "LL"
kernel = sys.argv[5] if len(sys.argv) > 5 else "matern5_2"
X, y = synth(n, d, 123)
with _capi.Engine(X, y, np.ones((n, 1)), kernel=kernel) as e:
    prev = e.launch_count
    for i in range(evals):
        v, g, info = e.objective(objective, np.full(d, 0.5), True, with_info=True)
        print(f"eval {i}: value={v:.12g} launches={e.launch_count - prev} stage_ms={info['stage_ms']}", flush=True)
        prev = e.launch_count
