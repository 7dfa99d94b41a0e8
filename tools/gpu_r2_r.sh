#!/bin/bash
# round 2, GPU call R: side-stream overlap with stream priorities; side split-K using all / half / a quarter of the SMs
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for rep in 1 2; do
  GPBO_OVERLAP_MAX=0 timeout 300 python tools/tail_bench.py 4096 2>/dev/null
  for d in 1 2 4; do GPBO_SIDE_SPLIT_DIV=$d timeout 300 python tools/tail_bench.py 4096 2>/dev/null; done
done 2>&1 | tee $O/r2r_tail.log
GPBO_OVERLAP_MAX=0 timeout 300 python tools/tail_bench.py 4096 fit 2>/dev/null | tee -a $O/r2r_tail.log
GPBO_SIDE_SPLIT_DIV=2 timeout 300 python tools/tail_bench.py 4096 fit 2>/dev/null | tee -a $O/r2r_tail.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q > $O/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2r_pytest.log
