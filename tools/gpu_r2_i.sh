#!/bin/bash
# round 2, GPU call I: A/B of the in-shared potf2 variants (tools/ab/*.so built with -DGPBO_CHOL32_* / -DGPBO_DIAGINV_*)
cd "$(dirname "$0")/.."
O=gpurun_out
for V in default v0 v2 v3 v4 tma tma_off; do
  unset GPBO_NO_TMA
  if [ $V = default ]; then unset GPBO_LIB; elif [ $V = tma_off ]; then export GPBO_LIB=$PWD/tools/ab/libgpbo_tma.so GPBO_NO_TMA=1; else export GPBO_LIB=$PWD/tools/ab/libgpbo_$V.so; fi
  echo "== variant $V"
  for B in 1 1036; do timeout 300 python tools/quick_bench.py 4096 $B skip 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('B',d['B'],'eval_s',round(d['eval_seconds'],5),'evals/s',round(d['evals_per_s'],2),{k:round(v[0],2) for k,v in d['profile_ms'].items() if v[1]})"; done
  timeout 300 python tools/small_dbg.py 2>&1 | grep -E "single pair|592 pairs|fit [0-9.]+ s" | grep -E "m=(20|90|200)|fit"
done 2>&1 | tee $O/r2i_ab.log
unset GPBO_NO_TMA; export GPBO_LIB=$PWD/tools/ab/libgpbo_tma.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q > $O/r2i_pytest_tma.log 2>&1; echo "pytest tma rc=$?"; tail -3 $O/r2i_pytest_tma.log
timeout 300 python tools/quick_bench.py 8192 148 skip 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('tma m=8192 B',d['B'],'eval_s',round(d['eval_seconds'],5),'evals/s',round(d['evals_per_s'],2),{k:round(v[0],2) for k,v in d['profile_ms'].items() if v[1]})" | tee -a $O/r2i_ab.log
GPBO_NO_TMA=1 timeout 300 python tools/quick_bench.py 8192 148 skip 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('no-tma m=8192 B',d['B'],'eval_s',round(d['eval_seconds'],5),'evals/s',round(d['evals_per_s'],2),{k:round(v[0],2) for k,v in d['profile_ms'].items() if v[1]})" | tee -a $O/r2i_ab.log
