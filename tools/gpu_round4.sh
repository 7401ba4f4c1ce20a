#!/bin/bash
mkdir -p gpurun_out
echo "== determinism under concurrency (all launch-chain so that the single-handle reference is comparable bit for bit)"
for i in 1 2; do LKGPU_STEP_TRSV=1 python tools/diag_concurrent2.py 5000 4 10 2>&1 | grep -E "mismatching|vectors|thread"; done
LKGPU_STEP_TRSV=1 python tools/diag_concurrent2.py 5000 8 8 2>&1 | grep -E "mismatching|vectors|thread"
echo "== pytest update + nested" ; (time timeout 900 python -m pytest tests/test_gpu_update.py tests/test_nested.py -m gpu -q -s) > gpurun_out/pytest_update.log 2>&1 ; tail -8 gpurun_out/pytest_update.log; grep "sub-model fits\|block extension chol" gpurun_out/pytest_update.log
echo "== pytest -m gpu" ; (time timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_update.py --deselect tests/test_nested.py) > gpurun_out/pytest_gpu.log 2>&1 ; tail -8 gpurun_out/pytest_gpu.log
echo "== bench" ; (time timeout 900 python bench.py) > gpurun_out/bench.log 2>&1 ; tail -4 gpurun_out/bench.log | cut -c1-600
