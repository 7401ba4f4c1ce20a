#!/bin/bash
O=gpurun_out/r02c6; mkdir -p $O
echo "== pytest -m gpu"; (time timeout 1500 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; tail -8 $O/pytest_gpu.log
echo "== bench (default)"; (time timeout 1500 python bench.py) > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.json; tail -3 $O/bench.err
echo "== cpp diag"; timeout 600 python tools/diag_ladder.py cpp 2>&1 | tee $O/diag_cpp.log | tail -12
ls $O
