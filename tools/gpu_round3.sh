#!/bin/bash
mkdir -p gpurun_out
echo "== determinism under concurrency"
for i in 1 2 3 4; do python tools/diag_concurrent2.py 5000 4 10 2>&1 | grep -E "mismatching|vectors|thread"; done
for i in 1 2; do python tools/diag_concurrent2.py 5000 8 8 2>&1 | grep -E "mismatching|vectors|thread"; done
echo "== pytest update" ; (time timeout 900 python -m pytest tests/test_gpu_update.py -x -q -s) > gpurun_out/pytest_update.log 2>&1 ; tail -8 gpurun_out/pytest_update.log
echo "== pytest -m gpu" ; (time timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_update.py) > gpurun_out/pytest_gpu.log 2>&1 ; tail -8 gpurun_out/pytest_gpu.log
echo "== bench" ; (time timeout 900 python bench.py) > gpurun_out/bench.log 2>&1 ; tail -4 gpurun_out/bench.log
