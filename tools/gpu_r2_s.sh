#!/bin/bash
# round 2, GPU call S: the complete BASELINE configs[3] fit on one B200 (r = 64 x m = 4096 x 32 starts to convergence)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1100 python tools/full_fit.py 64 4096 32 > $O/r2s_full_fit.json 2> $O/r2s_full_fit.err; echo "rc=$?"; cat $O/r2s_full_fit.json; tail -3 $O/r2s_full_fit.err
