#!/bin/bash
# round 2, GPU call G: tests after epilogue prefill / flat assembly / warp chol32 inverse; per-class timing at m=4096; assembly
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2g_pytest.log
for B in 1 8 1036; do timeout 300 python tools/quick_bench.py 4096 $B skip 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('B',d['B'],'eval_s',round(d['eval_seconds'],5),'evals/s',round(d['evals_per_s'],2),{k:round(v[0],2) for k,v in d['profile_ms'].items() if v[1]})"; done | tee gpurun_out/r2g_quick.log
timeout 600 python tools/asm_bench.py > gpurun_out/r2g_asm.log 2>&1; echo "asm rc=$?"; grep -c . gpurun_out/r2g_asm.log
GPBO_SMALL_DBG=1 timeout 300 python tools/small_dbg.py 2>&1 | grep -v "lml_grad m=" | awk '!seen[$0]++' | tail -12 | tee gpurun_out/r2g_small.log
timeout 300 python tools/pred_bench.py 8 4096 > gpurun_out/r2g_pred.json 2> gpurun_out/r2g_pred.err; echo "pred rc=$?"; cut -c1-1500 gpurun_out/r2g_pred.json
GPBO_NO_SWEEP=1 timeout 300 python tools/pred_bench.py 8 4096 > gpurun_out/r2g_pred_nosweep.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2g_pred_nosweep.json')); print('no sweep trsm', d['trsm']['frac'], d['device_ms']['cross_panel'])"
timeout 200 tools/mma_tma > gpurun_out/r2g_mma_tma.txt 2>&1; echo "mma_tma rc=$?"; cat gpurun_out/r2g_mma_tma.txt
timeout 200 tools/mma_sweep 2>&1 | grep -E "peak|bk16_s4_mbar" > gpurun_out/r2g_mma_sweep.txt; cat gpurun_out/r2g_mma_sweep.txt
