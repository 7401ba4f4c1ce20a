#!/bin/bash
# TRTRI table orders (LAUUM order 3 / band 16 now the default)
O=gpurun_out/r02c16; mkdir -p $O
for to in 1 2; do
  echo "== TRTRI_ORDER=$to" | tee -a $O/order.log
  LKGPU_TRTRI_ORDER=$to timeout 300 python tools/profile_eval.py 20000 10 3 2>&1 | tail -1 | tee -a $O/order.log
  LKGPU_TRTRI_ORDER=$to timeout 300 python tools/profile_eval.py 5000 20 4 LL gauss 2>&1 | tail -1 | tee -a $O/order.log
  LKGPU_TRTRI_ORDER=$to timeout 300 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:gemm_dmma -s 337 -c 16 --clock-control none python tools/profile_eval.py 20000 10 1 2>&1 | grep -E "dram__bytes_read|gpu__time" | awk '{a[$1]+=$3} END {for (k in a) print k, a[k]}' | tee -a $O/order.log
done
echo "== parity (TRTRI_ORDER=2)"; (time LKGPU_TRTRI_ORDER=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_reference.py -m gpu -q -x) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
