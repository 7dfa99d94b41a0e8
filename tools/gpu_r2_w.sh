#!/bin/bash
# round 2, GPU call W: where does the side-stream overlap stop paying?  GPBO_OVERLAP_MAX = 48 (default) vs 160
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
export TAIL_BATCHES=40,56,64,80,96,128,148
for rep in 1 2; do
  GPBO_OVERLAP_MAX=0 timeout 200 python tools/tail_bench.py 4096 2>/dev/null
  GPBO_OVERLAP_MAX=160 timeout 200 python tools/tail_bench.py 4096 2>/dev/null
done | tee $O/r2w_tail.log
