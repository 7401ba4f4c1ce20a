#!/bin/bash
# A/B: rest update after / beside the look-ahead update; outer block size at mid n
O=gpurun_out/r02c13; mkdir -p $O
for ual in 1 0; do
  echo "== UPDATE_AFTER_LOOKAHEAD=$ual" | tee -a $O/ual.log
  LKGPU_UPDATE_AFTER_LOOKAHEAD=$ual timeout 300 python tools/profile_eval.py 20000 10 3 2>&1 | tail -1 | tee -a $O/ual.log
  LKGPU_UPDATE_AFTER_LOOKAHEAD=$ual timeout 300 python tools/profile_eval.py 10000 10 3 2>&1 | tail -1 | tee -a $O/ual.log
  LKGPU_UPDATE_AFTER_LOOKAHEAD=$ual timeout 300 python tools/profile_eval.py 5000 20 4 LL gauss 2>&1 | tail -1 | tee -a $O/ual.log
  LKGPU_UPDATE_AFTER_LOOKAHEAD=$ual timeout 300 python tools/profile_eval.py 2500 10 4 2>&1 | tail -1 | tee -a $O/ual.log
  LKGPU_UPDATE_AFTER_LOOKAHEAD=$ual timeout 300 python tools/profile_eval.py 1000 4 4 LL gauss 2>&1 | tail -1 | tee -a $O/ual.log
done
for ob in 3 6 8; do
  echo "== OUTER_PANELS=$ob (n = 5000, 2500, 10000)" | tee -a $O/ual.log
  LKGPU_OUTER_PANELS=$ob timeout 300 python tools/profile_eval.py 5000 20 4 LL gauss 2>&1 | tail -1 | tee -a $O/ual.log
  LKGPU_OUTER_PANELS=$ob timeout 300 python tools/profile_eval.py 2500 10 4 2>&1 | tail -1 | tee -a $O/ual.log
  LKGPU_OUTER_PANELS=$ob timeout 300 python tools/profile_eval.py 10000 10 3 2>&1 | tail -1 | tee -a $O/ual.log
done
echo "== concurrent"; for ual in 1 0; do LKGPU_UPDATE_AFTER_LOOKAHEAD=$ual timeout 300 python tools/bench_concurrent.py 5000 20 gauss 10 1,8 2>&1 | tail -2 | tee -a $O/ual.log; done
