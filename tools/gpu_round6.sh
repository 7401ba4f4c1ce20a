#!/bin/bash
mkdir -p gpurun_out
echo "== exclusive gate (default): 8 threads, unflagged handles vs a lone reference, bit for bit"
for i in 1 2; do DIAG_FLAG=0 python tools/diag_concurrent2.py 5000 8 40 2>&1 | grep -E "mismatching|thread"; done
echo "== pytest -m gpu" ; (time timeout 1800 python -m pytest tests -m gpu -q -s) > gpurun_out/pytest_gpu.log 2>&1 ; grep -E "passed|failed|sub-model fits|block extension" gpurun_out/pytest_gpu.log | tail -5
echo "== bench" ; (time timeout 900 python bench.py) > gpurun_out/bench.log 2>&1 ; tail -4 gpurun_out/bench.log | cut -c1-300
