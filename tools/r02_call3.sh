#!/bin/bash
# Round 2, GPU call 3: whole GPU suite (no -x), panel-kernel phase timing, the bench line of every BASELINE config,
# ncu captures of the rewritten side kernels and of the trailing update.
O=gpurun_out/r02c3; mkdir -p $O
echo "== pytest -m gpu"; (time timeout 1500 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; tail -8 $O/pytest_gpu.log
echo "== potf2 phases"; LKGPU_LIB=$PWD/libkriging_b200/_variants/lib_prof.so timeout 120 python tools/potf2_phases.py 2>&1 | tee $O/potf2_phases.log
echo "== bench (default)"; (time timeout 1200 python bench.py) > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.json; tail -3 $O/bench.err
for c in 5 1 3 4; do echo "== bench --config $c"; (time timeout 1200 python bench.py --config $c --no-cpu) > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; tail -c 400 $O/bench_cfg$c.json; tail -3 $O/bench_cfg$c.err; done
echo "== ncu launch list (2 evaluations)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches.csv python tools/profile_eval.py 20000 10 2 > $O/ncu_launches.log 2>&1; tail -2 $O/ncu_launches.log
echo "== ncu full captures"
for k in cov_build grad_reduce trsv_wave; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o $O/prof_$k -f python tools/profile_eval.py 20000 10 1 > $O/ncu_$k.log 2>&1; tail -1 $O/ncu_$k.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 12 -c 1 -o $O/prof_syrk -f python tools/profile_eval.py 20000 10 1 > $O/ncu_syrk.log 2>&1; tail -1 $O/ncu_syrk.log
ls -la $O
