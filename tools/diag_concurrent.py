"""Diagnostic: is a multistart fit bitwise the same with 1 and 4 handles in flight? (BASELINE cfg 5 shape)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libkriging_b200.kriging import GpuBackend, Kriging  # noqa: E402
from tests.util import synth  # noqa: E402

n, d = 5000, 20
X, y, _ = synth(n, d, 505, "smooth")
log = []


class Tracing(GpuBackend):
    def objective(self, name, gamma, want_grad):
        v, g = super().objective(name, gamma, want_grad)
        log.append((tuple(float(t) for t in gamma), float(v), int(self.info["n_jitter"]), float(self.info["rcond"])))
        return v, g


def run(con):
    log.clear()
    k = Kriging("gauss", concurrent_starts=con, backend_factory=Tracing)
    k.config.max_iteration = 4
    k.fit(y, X, optim="BFGS4", objective="LL")
    out = (k.fit_log["best_start"], k.fit_log["objective"], k.fit_log["n_eval"])
    k.close()
    return out, dict((g, (v, nj, rc)) for g, v, nj, rc in log)


for tag in sys.argv[1:] or ["1", "4", "4", "1"]:
    r, ev = run(int(tag))
    print("con", tag, r, "evals", len(ev), "jitter rungs", sum(v[1] for v in ev.values()), flush=True)
    if "ref" not in globals():
        ref = ev
    else:
        common = [g for g in ev if g in ref]
        bad = [(g, ev[g], ref[g]) for g in common if ev[g] != ref[g]]
        print("   common points", len(common), "differing", len(bad))
        for g, a, b in bad[:5]:
            print("     ", a, b)
