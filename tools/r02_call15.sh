#!/bin/bash
# LAUUM band orders
O=gpurun_out/r02c15; mkdir -p $O
for band in 4 8 16 32; do
  echo "== LAUUM_ORDER=3 BAND=$band" | tee -a $O/order.log
  LKGPU_LAUUM_ORDER=3 LKGPU_LAUUM_BAND=$band timeout 300 python tools/profile_eval.py 20000 10 3 2>&1 | tail -1 | tee -a $O/order.log
  LKGPU_LAUUM_ORDER=3 LKGPU_LAUUM_BAND=$band timeout 300 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:gemm_dmma -s 353 -c 1 --clock-control none python tools/profile_eval.py 20000 10 1 2>&1 | grep -E "dram__bytes_read|hit_rate|gpu__time" | tee -a $O/order.log
done
echo "== n=5000 / 40000 with band 8 + trtri serp"; 
LKGPU_TRTRI_SERP=1 LKGPU_LAUUM_ORDER=3 timeout 300 python tools/profile_eval.py 5000 20 4 LL gauss 2>&1 | tail -1 | tee -a $O/order.log
LKGPU_TRTRI_SERP=1 LKGPU_LAUUM_ORDER=3 timeout 300 python tools/profile_eval.py 40000 10 2 2>&1 | tail -1 | tee -a $O/order.log
