#!/bin/bash
O=gpurun_out/r02c4; mkdir -p $O
timeout 900 python tools/diag_ladder.py cpp 2>&1 | tee $O/diag_cpp.log
timeout 900 python tools/diag_ladder.py fit 2>&1 | tee $O/diag_fit.log
