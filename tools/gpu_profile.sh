#!/bin/bash
# ncu evidence for profiles/: launch list of two evaluations, DRAM traffic of every DMMA GEMM launch of one
# evaluation, and one full capture each of the two largest launches (LAUUM = launch 353 of the 354 DMMA launches; the first k = 768 trailing update = launch 12).
mkdir -p gpurun_out
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python tools/profile_eval.py 20000 10 2 > gpurun_out/ncu_launches.log 2>&1 ; tail -2 gpurun_out/ncu_launches.log
echo "== dram traffic of the gemm launches of one evaluation"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_dmma --clock-control none -c 354 --csv --log-file gpurun_out/gemm_traffic.csv python tools/profile_eval.py 20000 10 1 > gpurun_out/ncu_traffic.log 2>&1 ; tail -1 gpurun_out/ncu_traffic.log
echo "== full capture: LAUUM"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 353 -c 1 -o gpurun_out/prof_lauum -f python tools/profile_eval.py 20000 10 1 > gpurun_out/ncu_full1.log 2>&1 ; tail -1 gpurun_out/ncu_full1.log
echo "== full capture: first trailing update (k = 768)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 12 -c 1 -o gpurun_out/prof_syrk -f python tools/profile_eval.py 20000 10 1 > gpurun_out/ncu_full2.log 2>&1 ; tail -1 gpurun_out/ncu_full2.log
ls -la gpurun_out | tail -12
