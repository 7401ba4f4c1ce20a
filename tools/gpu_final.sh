#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; (time timeout 1800 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu.log 2>&1 ; tail -4 gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; tail -2 gpurun_out/smoke.log
echo "== bench" ; (time timeout 900 python bench.py) > gpurun_out/bench.log 2>&1 ; tail -4 gpurun_out/bench.log | cut -c1-400
echo "== reference arm" ; (time timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/bench_ref.log 2>&1 ; tail -2 gpurun_out/bench_ref.log | cut -c1-600
bash tools/gpu_profile.sh
