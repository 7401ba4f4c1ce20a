#!/bin/bash
# Round-2 evidence run on one B200: whole GPU suite, smoke, the bench line of BASELINE configs 2 (default), 1, 3, 5,
# ncu launch list of two evaluations, DRAM traffic of every DMMA launch of one evaluation, full captures of LAUUM, the
# first k = 768 trailing update and the largest TRTRI launch, compute-sanitizer on overlapping evaluations.
O=gpurun_out/${1:-r02final}; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
echo "== pytest -m gpu"; (time timeout 1500 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
echo "== bench (default)"; (time timeout 1500 python bench.py) > $O/bench.json 2> $O/bench.err; tail -c 500 $O/bench.json; tail -3 $O/bench.err
for c in 5 1 3; do echo "== bench --config $c"; (time timeout 900 python bench.py --config $c --no-cpu) > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; tail -c 300 $O/bench_cfg$c.json; tail -3 $O/bench_cfg$c.err; done
echo "== ncu launch list (2 evaluations)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches.csv python tools/profile_eval.py 20000 10 2 > $O/ncu_launches.log 2>&1; tail -2 $O/ncu_launches.log
echo "== dram traffic of the gemm launches of one evaluation"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_dmma --clock-control none -c 354 --csv --log-file $O/gemm_traffic.csv python tools/profile_eval.py 20000 10 1 > $O/ncu_traffic.log 2>&1; tail -1 $O/ncu_traffic.log
echo "== ncu full captures"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 353 -c 1 -f -o $O/prof_lauum python tools/profile_eval.py 20000 10 1 > $O/ncu_lauum.log 2>&1; tail -1 $O/ncu_lauum.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 12 -c 1 -f -o $O/prof_syrk python tools/profile_eval.py 20000 10 1 > $O/ncu_syrk.log 2>&1; tail -1 $O/ncu_syrk.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 351 -c 2 -f -o $O/prof_trtri python tools/profile_eval.py 20000 10 1 > $O/ncu_trtri.log 2>&1; tail -1 $O/ncu_trtri.log
echo "== two overlapping n = 20000 handles"; timeout 600 python tools/bench_concurrent.py 20000 10 matern5_2 3 1,2 2>&1 | tail -2 | tee $O/concurrent_n20000.log
echo "== 1..16 overlapping n = 5000 handles"; timeout 600 python tools/bench_concurrent.py 5000 20 gauss 10 1,2,4,8,16 2>&1 | tail -5 | tee $O/concurrent_n5000.log
echo "== compute-sanitizer (overlapping evaluations, n = 1500)"
timeout 600 compute-sanitizer --tool racecheck python tools/diag_concurrent2.py 1500 3 2 > $O/racecheck.log 2>&1; tail -3 $O/racecheck.log
timeout 600 compute-sanitizer --tool memcheck python tools/diag_concurrent2.py 1500 3 2 > $O/memcheck.log 2>&1; tail -3 $O/memcheck.log
ls -la $O
