"""Build liblkgpu.so (the C-ABI engine of include/lkgpu.h) in-tree with nvcc for sm_100a.

    python -m libkriging_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblkgpu.so")
SOURCES = ["engine.cu"]
HEADERS = ["common.cuh", "tile_tables.hpp", "cov.cuh", "gemm_dmma.cuh", "potrf_panel.cuh", "trsv.cuh", "trsv_wave.cuh", "variogram.cuh", "lmp_loo.cuh", "objective.inl",
           os.path.join("..", "..", "include", "lkgpu.h")]


def needs_build() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS + [os.path.join("..", "build.py")]:
        fp = os.path.join(CSRC, f)
        if os.path.isfile(fp) and os.path.getmtime(fp) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False, out: str | None = None) -> str:
    """Build the library.  `out` (or LKGPU_BUILD_OUT) names another output file: experiment variants next to the
    product library, selected at run time with LKGPU_LIB (libkriging_b200/_capi.py)."""
    out = out or os.environ.get("LKGPU_BUILD_OUT")
    if out:
        force = True
    elif not force and not needs_build():
        return LIB
    lib_out = os.path.abspath(out) if out else LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
           "-o", lib_out] + [os.path.join(CSRC, s) for s in SOURCES]
    # experiments: LKGPU_BUILD_FLAGS adds flags, LKGPU_BUILD_DROP removes defaults, LKGPU_BUILD_OUT names the output, e.g.
    #   LKGPU_BUILD_FLAGS="-DLKGPU_RING_RELEASE=0" LKGPU_BUILD_OUT=libkriging_b200/_variants/lib_A.so python -m libkriging_b200.build
    for drop in os.environ.get("LKGPU_BUILD_DROP", "").split():
        while drop in cmd:
            i = cmd.index(drop)
            if i > 0 and cmd[i - 1] == "-Xptxas":
                del cmd[i - 1:i + 1]
            else:
                del cmd[i]
    extra = os.environ.get("LKGPU_BUILD_FLAGS", "").split()
    if extra:
        cmd[1:1] = extra
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building liblkgpu.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return lib_out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
