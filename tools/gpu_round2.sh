#!/bin/bash
# GPU round: update-path tests first (new code), then the whole GPU suite, bench.
mkdir -p gpurun_out
echo "== pytest update" ; (time timeout 900 python -m pytest tests/test_gpu_update.py -x -q -s) > gpurun_out/pytest_update.log 2>&1 ; tail -15 gpurun_out/pytest_update.log
echo "== pytest -m gpu" ; (time timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_update.py) > gpurun_out/pytest_gpu.log 2>&1 ; tail -5 gpurun_out/pytest_gpu.log
echo "== bench" ; (time timeout 900 python bench.py) > gpurun_out/bench.log 2>&1 ; tail -4 gpurun_out/bench.log
