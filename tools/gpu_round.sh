#!/bin/bash
# One gpurun call: parity tests, smoke, bench, ncu launch list + one full capture of the dominant kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; (time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1 ; tail -5 gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; tail -2 gpurun_out/smoke.log
echo "== bench" ; (time timeout 900 python bench.py) > gpurun_out/bench.log 2>&1 ; tail -4 gpurun_out/bench.log
echo "== ncu launches" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv python tools/profile_eval.py 20000 10 2 > gpurun_out/ncu_launches.log 2>&1 ; tail -3 gpurun_out/ncu_launches.log
echo "== ncu full (trailing update, early panels)" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 30 -c 3 -o gpurun_out/prof_gemm -f python tools/profile_eval.py 20000 10 1 > gpurun_out/ncu_full.log 2>&1 ; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
