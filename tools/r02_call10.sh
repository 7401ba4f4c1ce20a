#!/bin/bash
# A/B of the TMA producer rewrite (elected lane + 3D boxes for M-major operands; LKGPU_NO_MM3=1: 2D boxes)
O=gpurun_out/r02c10; mkdir -p $O
echo "== mm3"; timeout 300 python tools/profile_eval.py 20000 10 3 2>&1 | tail -2 | tee -a $O/mm3.log
echo "== no mm3"; LKGPU_NO_MM3=1 timeout 300 python tools/profile_eval.py 20000 10 3 2>&1 | tail -2 | tee -a $O/mm3.log
echo "== n=5000 d=20 gauss"; timeout 300 python tools/profile_eval.py 5000 20 3 LL gauss 2>&1 | tail -1 | tee -a $O/mm3.log
echo "== n=1000 d=4 gauss"; timeout 300 python tools/profile_eval.py 1000 4 3 LL gauss 2>&1 | tail -1 | tee -a $O/mm3.log
echo "== parity"; (time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_reference.py -m gpu -q -x) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
echo "== cpp sharded"; (time timeout 900 python -m pytest tests/test_cpp_host.py -m gpu -q -x -k "sharded or concurrent") > $O/pytest_cpp.log 2>&1; tail -15 $O/pytest_cpp.log
echo "== ncu syrk"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 12 -c 1 -f -o $O/prof_syrk_mm3 python tools/profile_eval.py 20000 10 1 > $O/ncu_syrk.log 2>&1; tail -2 $O/ncu_syrk.log
