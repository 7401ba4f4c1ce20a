"""Dump the L-BFGS-B trajectory (gamma, f, g per objective call) of one fixture fit, for the device backend
(default) or the oracle backend (--oracle, CPU).  Debug aid for fit parity."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libkriging_b200 import kriging  # noqa: E402
from tests.util import load_refgen, synth  # noqa: E402

name = sys.argv[1]
oracle = "--oracle" in sys.argv
c = [c for c in load_refgen()["fits"] if c["name"] == name][0]
X, y, noise = synth(c["n"], c["d"], c["seed"], c.get("yfun", "prodsin"))
trace = []
if oracle:
    from tests.oracle_backend import OracleBackend as Base
else:
    Base = kriging.GpuBackend


class Tracing(Base):
    def objective(self, name, gamma, want_grad):
        v, g = super().objective(name, gamma, want_grad)
        trace.append(dict(gamma=[float(t) for t in gamma], f=float(v), g=None if g is None else [float(t) for t in g]))
        return v, g


k = kriging.Kriging(c["kernel"], c["noise_model"], backend_factory=Tracing)
k.fit(y, X, c.get("regmodel", "constant"), c.get("normalize", False), c["optim"], c["objective"],
      noise=noise if c["noise_model"] == "hetero" else None)
out = dict(name=name, backend="oracle" if oracle else "gpu", theta=[float(t) for t in k.theta()], ref_theta=c["theta"],
           trace=trace)
os.makedirs("gpurun_out", exist_ok=True)
fn = f"gpurun_out/trace_{name}_{'oracle' if oracle else 'gpu'}.json"
json.dump(out, open(fn, "w"))
print(fn, len(trace), out["theta"], c["theta"])
