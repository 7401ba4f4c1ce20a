#!/bin/bash
# A/B of programmatic dependent launch levels; cfg-1 fit fixture; C++ sharded fit timing
O=gpurun_out/r02c11; mkdir -p $O
for pdl in 0 1 2 3; do
  echo "== PDL $pdl" | tee -a $O/pdl.log
  LKGPU_PDL=$pdl timeout 300 python tools/profile_eval.py 20000 10 3 2>&1 | tail -1 | tee -a $O/pdl.log
  LKGPU_PDL=$pdl timeout 300 python tools/profile_eval.py 5000 20 4 LL gauss 2>&1 | tail -1 | tee -a $O/pdl.log
  LKGPU_PDL=$pdl timeout 300 python tools/profile_eval.py 1000 4 4 LL gauss 2>&1 | tail -1 | tee -a $O/pdl.log
done
echo "== concurrent throughput PDL 0/1/2"; for pdl in 0 1 2; do LKGPU_PDL=$pdl timeout 300 python tools/bench_concurrent.py 2>&1 | tail -4 | tee -a $O/pdl_concurrent.log; done
echo "== tests"; (time LKGPU_PDL=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_fullsize_reference.py -m gpu -q -x) > $O/pytest.log 2>&1; tail -6 $O/pytest.log
