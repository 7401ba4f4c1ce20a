#!/bin/bash
# 2 GPUs: on-device NCCL test of the sharded fit, the default bench line at N = 2, cfg 5 at N = 2
O=gpurun_out/r02c7; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
echo "== potf2 phases"; LKGPU_LIB=$PWD/libkriging_b200/_variants/lib_prof.so timeout 120 python tools/potf2_phases.py 2>&1 | tee $O/potf2_phases.log
echo "== multirank test"; (time timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_nested.py -m gpu -q) > $O/pytest_multirank.log 2>&1; tail -5 $O/pytest_multirank.log
echo "== bench N=2 (default)"; (time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu) > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 400 $O/bench_n2.json; tail -3 $O/bench_n2.err
echo "== bench N=2 cfg 5"; (time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 5 --no-cpu --no-cpp-host) > $O/bench_cfg5_n2.json 2> $O/bench_cfg5_n2.err; tail -c 400 $O/bench_cfg5_n2.json; tail -3 $O/bench_cfg5_n2.err
ls $O
