#!/bin/bash
# 2 GPUs: the default bench line at N = 2 with the C++ host's sharded fit (the path the round-end scaling run takes)
O=gpurun_out/r02c17; mkdir -p $O
echo "== bench N=2 (default)"; (time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --no-cpu --no-plain-ladder --no-update --steps 5) > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 300 $O/bench_n2.json; tail -3 $O/bench_n2.err
echo "== multirank tests"; (time timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q) > $O/pytest_multirank.log 2>&1; tail -3 $O/pytest_multirank.log
