"""tests/golden/make_golden_nested.py -- fixture generator (BUILD container only).

Runs the UNMODIFIED reference's NestedKriging (oracle/_ref/ref_nested_driver: src/lib/NestedKriging.cpp compiled where it
lies) with a Random partition and stores the partition it drew, every sub-model's fitted hyper-parameters after the
unification step and the unified (theta, sigma2, beta0) in tests/golden/refgen_nested.json (SURVEY.md §8 row f4:
NestedKriging.cpp:262-331).  Usage: python tests/golden/make_golden_nested.py"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.util import synth  # noqa: E402

DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_nested_driver")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refgen_nested.json")

CASES = [
    dict(name="nested-m52-n600-d3-p4", n=600, d=3, seed=61, kernel="matern5_2", groups=4, theta0=0.6),
    dict(name="nested-m32-n900-d2-p6", n=900, d=2, seed=62, kernel="matern3_2", groups=6, theta0=0.5),
    dict(name="nested-exp-n500-d4-p2-randomstart", n=500, d=4, seed=63, kernel="exp", groups=2),
]


def run(c):
    X, y, _ = synth(c["n"], c["d"], c["seed"], "smooth")
    with tempfile.TemporaryDirectory() as wd:
        np.asfortranarray(X).T.ravel().tofile(os.path.join(wd, "X.bin"))
        y.tofile(os.path.join(wd, "y.bin"))
        with open(os.path.join(wd, "cfg.txt"), "w") as f:
            for k in ("n", "d", "groups", "kernel"):
                f.write(f"{k}={c[k]}\n")
            if "theta0" in c:
                f.write(f"theta0={c['theta0']!r}\n")
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
        lp = os.path.join(ROOT, "oracle", "_ref", "ld_library_path.txt")
        env["LD_LIBRARY_PATH"] = open(lp).read().strip() + ":" + env.get("LD_LIBRARY_PATH", "")
        out = subprocess.run([DRIVER, wd], env=env, capture_output=True, text=True)
        if out.returncode != 0:
            raise RuntimeError(out.stderr[-2000:])
        return json.loads(out.stdout.strip().splitlines()[-1])


def main():
    res = []
    for c in CASES:
        r = run(c)
        res.append(dict(c, **r))
        print(c["name"], r["theta"], r["sigma2"], r["beta0"], [len(g) for g in r["groups"]])
    json.dump(dict(source="oracle/_ref/ref_nested_driver (unmodified libKriging NestedKriging, Partition::Random, PoE)",
                   generator="tests/golden/make_golden_nested.py", cases=res), open(OUT, "w"))


if __name__ == "__main__":
    main()
