#!/bin/bash
O=gpurun_out/r02c5; mkdir -p $O
echo "== pytest -m gpu"; (time timeout 1500 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; tail -8 $O/pytest_gpu.log
echo "== potf2 phases"; LKGPU_LIB=$PWD/libkriging_b200/_variants/lib_prof.so timeout 120 python tools/potf2_phases.py 2>&1 | tee $O/potf2_phases.log
echo "== small n stages"; for cfg in "5000 20 3 LL gauss" "2500 6 3 LL matern5_2" "1000 4 3 LL gauss"; do timeout 300 python tools/profile_eval.py $cfg 2>&1 | tail -1 | tee -a $O/small.log; done
echo "== ladder trace"; timeout 900 python tools/diag_ladder.py fit > $O/diag_fit.log 2>&1; head -3 $O/diag_fit.log; awk 'NR>2{n++; c+=$4} END{print "evals",n,"chol_ms",c}' $O/diag_fit.log
echo "== bench (default)"; (time timeout 1200 python bench.py) > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.json; tail -3 $O/bench.err
echo "== concurrent"; timeout 300 python tools/bench_concurrent.py 5000 20 gauss 8 1,4,8,16 2>&1 | tee $O/concurrent.log
ls $O
