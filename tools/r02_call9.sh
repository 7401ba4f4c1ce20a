#!/bin/bash
# A/B of the Cholesky trailing-update tile order (LKGPU_CHOL_BAND: 0 = closed-form column order, else rows per band / 64)
O=gpurun_out/r02c9; mkdir -p $O
for band in 0 32 16 64; do
  echo "== band $band"; LKGPU_CHOL_BAND=$band timeout 300 python tools/profile_eval.py 20000 10 3 2>&1 | tail -2 | tee -a $O/band.log
done
echo "== band 32, 8 outer panels"; LKGPU_OUTER_PANELS=8 timeout 300 python tools/profile_eval.py 20000 10 3 2>&1 | tail -1 | tee -a $O/band.log
echo "== band 32, 5 outer panels"; LKGPU_OUTER_PANELS=5 timeout 300 python tools/profile_eval.py 20000 10 3 2>&1 | tail -1 | tee -a $O/band.log
echo "== n=40000"; timeout 300 python tools/profile_eval.py 40000 10 2 2>&1 | tail -1 | tee -a $O/band.log
echo "== n=40000 band 0"; LKGPU_CHOL_BAND=0 timeout 300 python tools/profile_eval.py 40000 10 2 2>&1 | tail -1 | tee -a $O/band.log
echo "== parity"; (time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_reference.py -m gpu -q -x) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
echo "== ncu syrk"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 12 -c 1 -f -o $O/prof_syrk_band python tools/profile_eval.py 20000 10 1 > $O/ncu_syrk.log 2>&1; tail -2 $O/ncu_syrk.log
