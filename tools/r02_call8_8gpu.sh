#!/bin/bash
# 8 GPUs: BASELINE cfg 5 (gauss n = 5000, BFGS64: 8 starts in flight per GPU, dynamic queue; Python host and the C++
# host's sharded fit) and cfg 4 (Nugget matern3_2 n = 40000, BFGS8: one start per GPU) as fits, with per-rank balance.
O=gpurun_out/r02c8; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
export NCCL_DEBUG=WARN
echo "== bench N=8 cfg 5"; (time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --config 5 --no-cpu --no-update) > $O/bench_cfg5_n8.json 2> $O/bench_cfg5_n8.err; tail -c 600 $O/bench_cfg5_n8.json; tail -3 $O/bench_cfg5_n8.err
echo "== bench N=8 cfg 4"; (time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config 4 --no-cpu --no-cpp-host --no-update --no-batched --steps 3) > $O/bench_cfg4_n8.json 2> $O/bench_cfg4_n8.err; tail -c 600 $O/bench_cfg4_n8.json; tail -3 $O/bench_cfg4_n8.err
ls $O
