"""Phase timing of the Cholesky panel kernel (potf2_inv_kernel) from clock64() stamps.
Needs a profiling build:  LKGPU_BUILD_FLAGS=-DLKGPU_POTF2_PROFILE LKGPU_BUILD_OUT=libkriging_b200/_variants/lib_prof.so \
                          python -m libkriging_b200.build ;  LKGPU_LIB=$PWD/libkriging_b200/_variants/lib_prof.so python tools/potf2_phases.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libkriging_b200 import _capi  # noqa: E402
from tests.util import synth  # noqa: E402

n, d = 128, 3
X, y, _ = synth(n, d, 5, "smooth")
names = ["load", "colsteps0", "rank32_0", "colsteps1", "rank32_1", "colsteps2", "rank32_2", "colsteps3", "(none)",
         "logdet+store L", "offdiag inverse", "store W"]
with _capi.Engine(X, y, np.ones((n, 1)), kernel="matern5_2") as e:
    for rep in range(3):
        e.objective("LL", np.full(d, 0.5), False)
    buf = (C.c_longlong * 32)()
    rc = _capi.lib().lkgpu_debug_potf2_profile(buf)
    t = np.array(buf[:13], dtype=np.int64)
    dt = np.diff(t)
    mhz = 1965.0
    for nm, c in zip(names, dt):
        print(f"{nm:18s} {int(c):8d} cycles  {c / mhz:7.2f} us")
    print(f"{'total':18s} {int(t[12] - t[0]):8d} cycles  {(t[12] - t[0]) / mhz:7.2f} us (at {mhz:.0f} MHz)")
