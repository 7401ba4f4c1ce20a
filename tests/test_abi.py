"""CPU-side checks of the drop-in boundary: liblkgpu.so loads, exports every symbol that
include/lkgpu.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lkgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lkgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from libkriging_b200 import build
    lib = ctypes.CDLL(build.build())
    names = _declared()
    assert len(names) >= 14
    for nme in names:
        assert hasattr(lib, nme), f"{nme} declared in include/lkgpu.h but not exported"
    lib.lkgpu_abi_version.restype = ctypes.c_int
    assert lib.lkgpu_abi_version() == 1


def test_out_struct_layout_matches_header():
    from libkriging_b200 import _capi
    src = open(os.path.join(ROOT, "include", "lkgpu.h")).read()
    body = src[src.index("typedef struct lkgpu_out {"):src.index("} lkgpu_out;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(?:double\*?|int)\s+\*?([A-Za-z0-9_]+)(?:\[[A-Z_]+\])?;", body)
    assert fields == [f for f, _ in _capi.LkgpuOut._fields_]


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from libkriging_b200 import _capi
    X = np.random.rand(10, 2)
    with pytest.raises(_capi.LkgpuError, match="no CUDA device"):
        _capi.Engine(X, X[:, 0], np.ones((10, 1)))
    with pytest.raises(_capi.LkgpuError):
        _capi.probe_fp64_peak()
