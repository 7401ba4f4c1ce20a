"""Stage timings of one evaluation at each BASELINE.json config shape, and (optionally) a timed trace of the bench
fit.  Diagnostic only -- numbers printed here are not bench values.
    python tools/config_sweep.py [cfg ...]        cfg in {1,2,3,4,5,fit}
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth  # noqa: E402
from libkriging_b200 import _capi, kriging  # noqa: E402

CFG = {
    "1": dict(kernel="gauss", noise_model="none", n=1000, d=4, obj="LL", theta=0.3),
    "2": dict(kernel="matern5_2", noise_model="none", n=20000, d=10, obj="LL", theta=0.5),
    "3": dict(kernel="exp", noise_model="none", n=10000, d=6, obj="LOO", theta=0.5),
    "4": dict(kernel="matern3_2", noise_model="nugget", n=40000, d=8, obj="LL", theta=0.5),
    "5": dict(kernel="gauss", noise_model="none", n=5000, d=20, obj="LL", theta=1.0),
}


def run_cfg(name):
    c = CFG[name]
    X, y = synth(c["n"], c["d"], 123)
    F = np.ones((c["n"], 1))
    t0 = time.perf_counter()
    with _capi.Engine(X, y, F, kernel=c["kernel"], noise_model=c["noise_model"]) as e:
        t_create = time.perf_counter() - t0
        gamma = np.full(c["d"], c["theta"])
        if c["noise_model"] == "nugget":
            gamma = np.append(gamma, 0.9)
        for want_grad in (False, True):
            best = None
            for rep in range(3):
                t0 = time.perf_counter()
                v, g, info = e.objective(c["obj"], gamma, want_grad, with_info=True)
                wall = (time.perf_counter() - t0) * 1e3
                if best is None or wall < best[0]:
                    best = (wall, v, info)
            wall, v, info = best
            st = {k: round(x, 3) for k, x in info["stage_ms"].items() if x > 0}
            print(f"cfg{name} {c['kernel']}/{c['noise_model']}/{c['obj']} n={c['n']} d={c['d']} grad={want_grad}: "
                  f"wall {wall:.2f} ms value={v:.10g} jitter={info['n_jitter']} rcond={info['rcond']:.3g} "
                  f"create={t_create:.2f}s stages={json.dumps(st)}", flush=True)


def run_fit(n=20000, d=10):
    X, y = synth(n, d, 123)
    F = np.ones((n, 1))
    with _capi.Engine(X, y, F, kernel="matern5_2") as e0:
        e0.objective("LL", np.full(d, 0.5), False)
        L = e0.export("L")
    z = np.random.Generator(np.random.PCG64(321)).standard_normal(n)
    y = 1.5 + 2.0 * (L @ z)
    del L
    trace = []

    class Tracing(kriging.GpuBackend):
        def objective(self, name, gamma, want_grad):
            t0 = time.perf_counter()
            v, g = super().objective(name, gamma, want_grad)
            trace.append(dict(wall_ms=(time.perf_counter() - t0) * 1e3, f=float(v), grad=bool(want_grad),
                              theta=[float(np.exp(t)) for t in gamma], n_jitter=self.info["n_jitter"],
                              rcond=self.info["rcond"], stage_ms=self.info["stage_ms"]))
            return v, g

    k = kriging.Kriging("matern5_2", backend_factory=Tracing)
    t0 = time.perf_counter()
    k.fit(y, X, optim="BFGS", objective="LL")
    wall = time.perf_counter() - t0
    print(f"fit n={n} d={d}: wall {wall:.2f} s, {len(trace)} objective calls, sum of call walls "
          f"{sum(t['wall_ms'] for t in trace) / 1e3:.2f} s", flush=True)
    for i, t in enumerate(trace):
        st = t["stage_ms"]
        print(f"  {i:3d} wall {t['wall_ms']:8.1f} ms f={t['f']:.8g} jit={t['n_jitter']} rcond={t['rcond']:.2e} "
              f"theta[0..2]={[round(x, 3) for x in t['theta'][:3]]} cov={st['cov']:.1f} chol={st['chol']:.1f} "
              f"rcond_ms={st['rcond']:.1f} trtri={st['trtri']:.1f} lauum={st['lauum']:.1f} total={st['total']:.1f}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(dict(wall_s=wall, trace=trace), open("gpurun_out/fit_trace_bench.json", "w"))


if __name__ == "__main__":
    for a in (sys.argv[1:] or ["1", "2", "3", "4", "5"]):
        if a == "fit":
            run_fit()
        else:
            run_cfg(a)
