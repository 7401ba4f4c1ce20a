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


def test_ring_release_is_scheduled_after_the_stage_math():
    """SASS guard for the TMA ring kernels (gemm_dmma_kernel, trsv_wave_kernel): the release of a ring slot
    (SYNCS.ARRIVE...A1T0 on the empty barrier) must come after the last math instruction that consumes the slot's
    shared-memory reads.  ptxas 12.9 hoisted it to just after the last LDS issue when nothing stopped it, and under
    heavy shared-memory traffic the refill then overtook queued loads ("Concurrent handles"); the kernels
    keep it in place with a never-taken fence.  This test fails if a compiler change undoes that."""
    import shutil
    import subprocess
    from libkriging_b200 import build
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True, check=True).stdout
    seen = 0
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        name = f.split("\n", 1)[0].strip()
        if "gemm_dmma_kernel" not in name and "trsv_wave_kernel" not in name:
            continue
        ins = [m.group(1).strip() for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(.*?);", f)]
        math = "DMMA" if "gemm_dmma" in name else "DFMA"
        rel = [i for i, t in enumerate(ins) if "SYNCS.ARRIVE.TRANS64.A1T0" in t]
        assert len(rel) == 1, (name, len(rel))
        r = rel[0]
        # backwards from the release to the consumer's full-barrier wait: the stage's math sits in between ...
        w = max(i for i in range(r) if "SYNCS.PHASECHK" in ins[i])
        n_math = sum(1 for t in ins[w:r] if re.match(r"(@!?U?P\d+\s+)?" + math, t))
        assert n_math >= (128 if math == "DMMA" else 16), (name, n_math)
        assert any("FENCE" in t for t in ins[r - 8:r]), name
        # ... and none of it after the release, up to the next barrier wait / branch
        tail = []
        for t in ins[r + 1:]:
            if "SYNCS.PHASECHK" in t or re.match(r"(@!?U?P\d+\s+)?BRA", t):
                break
            tail.append(t)
        assert not [t for t in tail if re.match(r"(@!?U?P\d+\s+)?" + math, t)], (name, tail)
        seen += 1
    assert seen == 11  # 3 GEMM layouts + 8 sweep variants
