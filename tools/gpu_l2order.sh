#!/bin/bash
mkdir -p gpurun_out
echo "== old order"; LKGPU_NO_L2_ORDER=1 python tools/profile_eval.py 20000 10 3 2>&1 | tail -2
echo "== L2 order";  python tools/profile_eval.py 20000 10 3 2>&1 | tail -2
echo "== dram traffic (L2 order)"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_dmma --clock-control none -c 367 --csv --log-file gpurun_out/gemm_traffic_l2.csv python tools/profile_eval.py 20000 10 1 > gpurun_out/ncu_traffic_l2.log 2>&1 ; tail -1 gpurun_out/ncu_traffic_l2.log
echo "== parity subset"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -3
