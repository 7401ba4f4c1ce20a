#!/bin/bash
# Round-2 starting point: validate the removal of the defensive measures that predate the ring-release fix
# ("The ring release, and concurrent handles"), one at a time, with the two reproducers:
#   tools/diag_foreign.py elementwise N   one handle next to a bandwidth-bound foreign kernel (was 95 % deviating)
#   tools/diag_concurrent2.py n T R       T overlapping handles, bit-for-bit against a lone reference (was 0.5-3 %)
# Host-side switches (no rebuild): LKGPU_OVERLAP_DEFAULT=1, LKGPU_TRTRI_NOSYNC=1, LKGPU_WAVE_WHEN_SHARED=1.
# Compile-time measures to try afterwards, one rebuild each: fence_writes_for_tma() (common.cuh), -dlcm=cg and
# -D__restrict__= (libkriging_b200/build.py), the device-side fills / copies (engine.cu dev_zero / dev_copy).
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" 2>&1 | grep -E "deviating|mismatching|thread" | head -6; }
run LKGPU_WAVE_WHEN_SHARED=1 python tools/diag_concurrent2.py 5000 8 40
run LKGPU_WAVE_WHEN_SHARED=1 LKGPU_TRTRI_NOSYNC=1 python tools/diag_concurrent2.py 5000 8 40
run LKGPU_WAVE_WHEN_SHARED=1 LKGPU_TRTRI_NOSYNC=1 python tools/diag_foreign.py elementwise 60
run DIAG_FLAG=0 LKGPU_OVERLAP_DEFAULT=1 LKGPU_WAVE_WHEN_SHARED=1 LKGPU_TRTRI_NOSYNC=1 python tools/diag_concurrent2.py 5000 8 40
# Cholesky experiment: persistent trailing update with r SMs left to the panel chain
for r in 0 4 8 12 16; do echo "== LKGPU_PERSISTENT_UPDATE=$r"; LKGPU_PERSISTENT_UPDATE=$r python tools/profile_eval.py 20000 10 3 2>&1 | tail -1 | grep -o "chol.: [0-9.]*"; done
