#!/bin/bash
# LAUUM / TRTRI table orders: time (stage events) and DRAM traffic (ncu, 2 metrics) per order
O=gpurun_out/r02c14; mkdir -p $O
for lo in 0 1 2; do
  echo "== LAUUM_ORDER=$lo" | tee -a $O/order.log
  LKGPU_LAUUM_ORDER=$lo timeout 300 python tools/profile_eval.py 20000 10 3 2>&1 | tail -1 | tee -a $O/order.log
  LKGPU_LAUUM_ORDER=$lo timeout 300 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:gemm_dmma -s 353 -c 1 --clock-control none python tools/profile_eval.py 20000 10 1 2>&1 | grep -E "dram__bytes_read|hit_rate|gpu__time" | tee -a $O/order.log
done
for ts in 0 1; do
  echo "== TRTRI_SERP=$ts" | tee -a $O/order.log
  LKGPU_TRTRI_SERP=$ts timeout 300 python tools/profile_eval.py 20000 10 3 2>&1 | tail -1 | tee -a $O/order.log
  LKGPU_TRTRI_SERP=$ts timeout 300 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:gemm_dmma -s 351 -c 2 --clock-control none python tools/profile_eval.py 20000 10 1 2>&1 | grep -E "dram__bytes_read|hit_rate|gpu__time" | tee -a $O/order.log
done
echo "== racecheck after the panel write-back change"; timeout 600 compute-sanitizer --tool racecheck python tools/diag_concurrent2.py 1500 3 2 > $O/racecheck.log 2>&1; tail -3 $O/racecheck.log
echo "== fixed beta tests"; timeout 600 python -m pytest tests/test_gpu_fit.py tests/test_cpp_host.py -m gpu -q -k "fixed_beta" 2>&1 | tail -3
