#!/bin/bash
# Round 2, GPU call 1: A/B of the ring release + removal of the defensive measures, one library variant each
#   A = round-1 release (never-taken fence), all defensive flags       B = real proxy fence (product), all defensive flags
#   C = B without the writer-side fences                                D = C without -dlcm=cg / -D__restrict__=
# plus the persistent-update sweep, concurrent-handle throughput, and ncu full captures of the side kernels.
O=gpurun_out/r02c1; mkdir -p $O
V=libkriging_b200/_variants
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
free -g > $O/host.txt; nproc >> $O/host.txt
libs=("A:$V/lib_A.so" "B:libkriging_b200/liblkgpu.so" "C:$V/lib_C.so" "D:$V/lib_D.so")
for e in "${libs[@]}"; do
  k=${e%%:*}; L=$PWD/${e#*:}
  echo "== variant $k: timing n=20000" | tee -a $O/variants.log
  LKGPU_LIB=$L timeout 300 python tools/profile_eval.py 20000 10 4 2>&1 | tail -2 | tee -a $O/variants.log
  echo "== variant $k: foreign elementwise (wave + nosync)" | tee -a $O/variants.log
  LKGPU_LIB=$L LKGPU_WAVE_WHEN_SHARED=1 LKGPU_TRTRI_NOSYNC=1 timeout 300 python tools/diag_foreign.py elementwise 90 2>&1 | tail -3 | tee -a $O/variants.log
  echo "== variant $k: 8 overlapping handles (wave + nosync), bitwise" | tee -a $O/variants.log
  LKGPU_LIB=$L LKGPU_WAVE_WHEN_SHARED=1 LKGPU_TRTRI_NOSYNC=1 timeout 300 python tools/diag_concurrent2.py 5000 8 20 2>&1 | grep -E "identical|mismatching|thread" | head -5 | tee -a $O/variants.log
done
echo "== D, overlap as default (unflagged handles), 8 handles" | tee -a $O/variants.log
DIAG_FLAG=0 LKGPU_LIB=$PWD/$V/lib_D.so LKGPU_OVERLAP_DEFAULT=1 LKGPU_WAVE_WHEN_SHARED=1 LKGPU_TRTRI_NOSYNC=1 timeout 300 python tools/diag_concurrent2.py 5000 8 20 2>&1 | grep -E "identical|mismatching|thread" | head -5 | tee -a $O/variants.log
echo "== concurrent throughput (D, wave + nosync)" | tee -a $O/variants.log
LKGPU_LIB=$PWD/$V/lib_D.so LKGPU_WAVE_WHEN_SHARED=1 LKGPU_TRTRI_NOSYNC=1 timeout 300 python tools/bench_concurrent.py 5000 20 gauss 8 1,2,4,8 2>&1 | tee -a $O/concurrent.log
LKGPU_LIB=$PWD/$V/lib_D.so LKGPU_WAVE_WHEN_SHARED=1 LKGPU_TRTRI_NOSYNC=1 timeout 300 python tools/bench_concurrent.py 2500 6 matern5_2 8 1,4,8 2>&1 | tee -a $O/concurrent.log
echo "== stage timings small n (B)" | tee -a $O/small.log
for cfg in "5000 20 3 LL gauss" "2500 6 3 LL matern5_2" "1000 4 3 LL gauss" "10000 6 2 LOO exp"; do timeout 300 python tools/profile_eval.py $cfg 2>&1 | tail -1 | tee -a $O/small.log; done
echo "== persistent update sweep (D)" | tee -a $O/persist.log
for r in 0 4 8 16; do echo "LKGPU_PERSISTENT_UPDATE=$r" | tee -a $O/persist.log; LKGPU_LIB=$PWD/$V/lib_D.so LKGPU_PERSISTENT_UPDATE=$r timeout 300 python tools/profile_eval.py 20000 10 3 2>&1 | tail -1 | tee -a $O/persist.log; done
echo "== parity subset on B"
(time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q) > $O/pytest_parity.log 2>&1; tail -3 $O/pytest_parity.log
echo "== ncu full captures of the side kernels (B)"
for k in cov_build grad_reduce trsv_wave potf2_inv; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o $O/prof_$k -f python tools/profile_eval.py 20000 10 1 > $O/ncu_$k.log 2>&1; tail -1 $O/ncu_$k.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trsv_wave -s 2 -c 1 -o $O/prof_trsv_wave_bwd -f python tools/profile_eval.py 20000 10 1 > $O/ncu_trsv_bwd.log 2>&1
ls -la $O
