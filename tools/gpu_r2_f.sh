#!/bin/bash
# round 2, GPU call F: tests after the warp-level 32x32 Cholesky, small-path phase cycles, single-pair latency
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2f_pytest.log
GPBO_SMALL_DBG=1 timeout 300 python tools/small_dbg.py 2>&1 | grep -v "^\[gpbo small lml_grad m=2[02][04]\]" | awk '!seen[$0]++' > gpurun_out/r2f_small_dbg.log; echo "small dbg rc=$?"
grep -v "lml_grad m=" gpurun_out/r2f_small_dbg.log | tail -12; grep "lml_grad m=224\]" gpurun_out/r2f_small_dbg.log | head -1
for B in 1 8 1036; do timeout 300 python tools/quick_bench.py 4096 $B skip 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('B',d['B'],'eval_s',round(d['eval_seconds'],5),'evals/s',round(d['evals_per_s'],2),{k:round(v[0],2) for k,v in d['profile_ms'].items() if v[1]})"; done
