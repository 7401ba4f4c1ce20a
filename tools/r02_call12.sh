#!/bin/bash
O=gpurun_out/r02c12; mkdir -p $O
echo "== chol trace n=5000"; LKGPU_TRACE_CHOL=1 timeout 300 python tools/profile_eval.py 5000 20 2 LL gauss > $O/trace5000.out 2> $O/trace5000.err; tail -1 $O/trace5000.out; grep -c trace $O/trace5000.err
echo "== chol trace n=20000"; LKGPU_TRACE_CHOL=1 timeout 300 python tools/profile_eval.py 20000 10 2 > $O/trace20000.out 2> $O/trace20000.err; tail -1 $O/trace20000.out
echo "== host overheads"; timeout 900 python tools/probe_host_overheads.py 2>&1 | tee $O/host_overheads.log
