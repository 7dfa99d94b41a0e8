#!/bin/bash
# round 2, GPU call U: reference experiment configurations, this build (padding pivots skipped) vs the round-start build
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for rep in 1 2; do
  timeout 300 python tools/refcfg_bench.py 2>/dev/null
  GPBO_LIB=$PWD/tools/ab/libgpbo_head.so timeout 300 python tools/refcfg_bench.py 2>/dev/null
done | tee $O/r2u_refcfg.log
