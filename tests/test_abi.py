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


def test_ring_release_is_ordered_by_a_real_fence():
    """SASS guard for the TMA ring kernels (gemm_dmma_kernel, trsv_wave_kernel): between the last shared-memory read
    of a ring slot and the slot's release (SYNCS.ARRIVE...A1T0 on the empty barrier) every lane executes an
    UNCONDITIONAL memory barrier + async-proxy fence (fence.proxy.async.shared::cta -> MEMBAR.ALL.CTA +
    FENCE.VIEW.ASYNC.S).  The barrier cannot complete before the warp's LDS have, and the fence orders those
    generic-proxy reads before the TMA refill (async proxy) that the release allows: the order holds by construction,
    not by instruction scheduling.  (Round 1 pinned the release behind the stage's DMMAs with a never-taken branch;
    without anything ptxas 12.9 hoisted it to just after the last LDS *issue* and refills overtook queued loads under
    heavy shared-memory traffic, profiles/r01c_ring_release.md.)"""
    import shutil
    import subprocess
    from libkriging_b200 import build
    assert shutil.which("cuobjdump") is not None, "cuobjdump (CUDA toolkit) is required for the SASS guard"
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True, check=True).stdout
    seen = 0
    pred = r"(@!?U?P\d+\s+)?"
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        name = f.split("\n", 1)[0].strip()
        if "gemm_dmma_kernel" not in name and "trsv_wave_kernel" not in name:
            continue
        ins = [m.group(1).strip() for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(.*?);", f)]
        math = "DMMA" if "gemm_dmma" in name else "DFMA"
        rel = [i for i, t in enumerate(ins) if "SYNCS.ARRIVE.TRANS64.A1T0" in t]
        assert len(rel) == 1, (name, len(rel))
        r = rel[0]
        # the consumer's stage: from its full-barrier wait to the release
        w = max(i for i in range(r) if "SYNCS.PHASECHK" in ins[i])
        seg = ins[w:r]
        n_math = sum(1 for t in ins[w:] if re.match(pred + math, t))
        assert n_math >= (128 if math == "DMMA" else 16), (name, n_math)
        # every read of the slot is a shared-memory load (no generic loads on the ring) ...
        assert not [t for t in seg if re.match(pred + r"LD\.E", t) and "STRONG" in t and "0x3" in t], name
        lds = [i for i, t in enumerate(seg) if re.match(pred + "LDS", t)]
        assert lds, name
        after = seg[lds[-1] + 1:]
        # ... and after the last of them, before the release: an unpredicated barrier + proxy fence, not skippable
        mb = [i for i, t in enumerate(after) if t.startswith("MEMBAR.ALL")]
        fv = [i for i, t in enumerate(after) if t.startswith("FENCE.VIEW.ASYNC")]
        assert mb and fv and mb[0] < fv[0], (name, after)
        assert not [t for t in after if re.match(pred + r"(BRA|BRX|JMP)", t)], (name, after)
        seen += 1
    assert seen == 11  # 3 GEMM layouts + 8 sweep variants
