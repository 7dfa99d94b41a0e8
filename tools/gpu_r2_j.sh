#!/bin/bash
# round 2, GPU call J: TMA staging on/off, interleaved and repeated (the boxes drift by 2-3 % as they warm up), tests with
# TMA on, prediction sweep scheduling modes
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2j_pytest.log
run() {  # $1 label, $2 m, $3 B
  timeout 300 python tools/quick_bench.py $2 $3 skip 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1 m',d['m'],'B',d['B'],'evals/s',round(d['evals_per_s'],2),{k:round(v[0],2) for k,v in d['profile_ms'].items() if v[1]})"
}
for rep in 1 2 3; do
  unset GPBO_NO_TMA; run tma 4096 1036; GPBO_NO_TMA=1 run ldgsts 4096 1036
  unset GPBO_NO_TMA; run tma 8192 148; GPBO_NO_TMA=1 run ldgsts 8192 148
done 2>&1 | tee $O/r2j_tma_ab.log
unset GPBO_NO_TMA
for MODE in 1 0 1 0; do
  GPBO_SWEEP_MODE=$MODE timeout 300 python tools/pred_bench.py 8 4096 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('sweep mode $MODE trsm', round(d['trsm']['frac'],4), d['device_ms']['cross_panel'], 'schur', round(d['schur']['frac'],4), 'chol', d['device_ms']['chol_diag'], d['device_ms']['chol_panel'])"
done 2>&1 | tee $O/r2j_sweep.log
