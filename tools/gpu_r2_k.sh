#!/bin/bash
# round 2, GPU call K: tests; symmetric-assembly regression hunt (this build vs the round-2a build, interleaved); narrow
# assembly kernel on/off; trtri tile order; prediction roofline with the single-site sweep kernel
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2k_pytest.log
asm() {  # $1 label
  timeout 300 python tools/asm_bench.py 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
r={}
for a in d['assemble']:
    r.setdefault((a['n1'],a['n2'],a['sym']),[]).append(round(a['frac'],3))
for k,v in r.items(): print('$1',k,v)"
}
for rep in 1 2; do
  unset GPBO_LIB GPBO_NO_NARROW; asm new
  GPBO_LIB=$PWD/tools/ab/libgpbo_r2d.so asm r2d
done 2>&1 | tee $O/r2k_asm_ab.log
GPBO_NO_NARROW=1 asm no_narrow 2>&1 | grep "200" | tee -a $O/r2k_asm_ab.log
run() {  # $1 label, $2 m, $3 B
  timeout 300 python tools/quick_bench.py $2 $3 skip 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1 m',d['m'],'B',d['B'],'evals/s',round(d['evals_per_s'],2),{k:round(v[0],2) for k,v in d['profile_ms'].items() if v[1]})"
}
for rep in 1 2; do
  GPBO_TRTRI_LPT=0 run pairmajor 8192 148; GPBO_TRTRI_LPT=1 run lpt 8192 148
  GPBO_TRTRI_LPT=0 run pairmajor 8192 256; GPBO_TRTRI_LPT=1 run lpt 8192 256
done 2>&1 | tee $O/r2k_lpt.log
timeout 300 python tools/pred_bench.py 8 4096 > $O/r2k_pred.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2k_pred.json')); print('trsm', d['trsm']['frac'], d['device_ms'], 'schur', d['schur']['frac'])"
timeout 300 python tools/pred_bench.py 6 200 > $O/r2k_pred_small.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2k_pred_small.json')); print('m=200', d['device_ms'])"
