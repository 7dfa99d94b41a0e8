#!/bin/bash
# round 2, GPU call B: full GPU tests + phase timing of the small-matrix path
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2b_pytest.log
GPBO_SMALL_DBG=1 timeout 300 python tools/small_dbg.py > gpurun_out/r2b_small_dbg.log 2>&1; echo "small dbg rc=$?"
tail -30 gpurun_out/r2b_small_dbg.log
