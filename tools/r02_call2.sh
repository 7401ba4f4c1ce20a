#!/bin/bash
# Round 2, GPU call 2: full GPU suite on the relaxed default build (+ ladder shortcut, overlap by default, fixed beta),
# the reference arm MEASURED at n = 20000 (CPU, in parallel with the test-suite; writes the cfg2 golden), then the
# bench line and the concurrency experiments.
O=gpurun_out/r02c2; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
( time timeout 1700 python bench.py --impl reference --steps 20 --warmup 5 --golden-out $O/cfg2_golden.json > $O/bench_ref.json 2> $O/bench_ref.err ) 2> $O/bench_ref.time &
REFPID=$!
echo "== pytest -m gpu"; (time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
wait $REFPID; echo "== reference arm"; tail -c 1500 $O/bench_ref.json; cat $O/bench_ref.time
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
echo "== bench (default)"; (time timeout 1200 python bench.py) > $O/bench.json 2> $O/bench.err; tail -c 3000 $O/bench.json; tail -3 $O/bench.err
echo "== fit with the plain ladder (LKGPU_FULL_LADDER=1)"; LKGPU_FULL_LADDER=1 timeout 900 python bench.py --steps 3 --no-cpu --no-update --no-batched > $O/bench_full_ladder.json 2>$O/bench_full_ladder.err; python - <<'PY'
import json
for f in ("gpurun_out/r02c2/bench.json","gpurun_out/r02c2/bench_full_ladder.json"):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); ft=j["fit"]
        print(f, "wall", ft["wall_s"], "evals", ft["n_eval_all_ranks"], "LL", ft["LL_at_fit"], "theta0", ft["theta"][:3], ft["ladder"])
    except Exception as e: print(f, "failed", e)
PY
echo "== concurrency experiments (n=5000 d=20 gauss)"
for env in "X=1" "LKGPU_NO_PERSISTENT=1" "CUDA_DEVICE_MAX_CONNECTIONS=32" "CUDA_DEVICE_MAX_CONNECTIONS=32 LKGPU_NO_PERSISTENT=1"; do
  echo "-- $env" | tee -a $O/concurrent.log
  env $env timeout 300 python tools/bench_concurrent.py 5000 20 gauss 8 1,2,4,8,16 2>&1 | tee -a $O/concurrent.log
done
echo "-- n=2500 d=6 m52, n=1000 d=4 gauss" | tee -a $O/concurrent.log
timeout 300 python tools/bench_concurrent.py 2500 6 matern5_2 8 1,8,16 2>&1 | tee -a $O/concurrent.log
timeout 300 python tools/bench_concurrent.py 1000 4 gauss 16 1,8,16 2>&1 | tee -a $O/concurrent.log
ls -la $O
