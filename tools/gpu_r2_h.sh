#!/bin/bash
# round 2, GPU call H: the driver's N=1 command, its ncu launch list (short form of the same command), ncu --set full
# captures of the dominant kernels at the headline size (m = 8192) and of the new prediction / assembly kernels
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2h_bench_n1.json 2> $O/r2h_bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2h_bench_n1.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','steps','fits_per_s','gpu_launches')}, d['e2e']['value'], d['clocks'], d.get('seconds_total'))
r=d['roofline']; print(r['kernel'], r['frac'], r['whole_eval_frac'], r['class_frac_algorithmic'])
print('configs3', d.get('evals_configs3'))
p=d.get('roofline_prediction'); print('pred', p if isinstance(p,str) else {k:p[k] for k in ('trsm','schur','mean','std','sqrtw_ms','sqrtw_iterations')})
a=d.get('roofline_assembly'); print('asm', a if isinstance(a,str) else (a['min_frac'], {k:v for k,v in a['frac'].items() if v<0.8}))
print('cpu', d.get('cpu_baseline')); print('fit_sample', d.get('fit_sample'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/r2h_launches.csv \
    python bench.py --gpus 1 --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/r2h_ncu_list.log 2>&1; echo "ncu list rc=$?"
for K in trtri_row_kernel lauum_grad_kernel chol_panel_kernel chol_diag_kernel; do
  S=40; [ $K = lauum_grad_kernel ] && S=0
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o $O/prof_$K \
      python tools/quick_bench.py 8192 148 skip > $O/r2h_ncu_$K.log 2>&1; echo "ncu $K rc=$?"
  python tools/ncu_summary.py $O/prof_$K.ncu-rep > $O/r2h_ncu_summary_$K.txt 2>&1; rm -f $O/prof_$K.ncu-rep
done
ncu --set full --clock-control none --import-source on -k regex:cross_sweep_kernel -s 1 -c 1 -f -o $O/prof_cross_sweep \
    python tools/pred_bench.py 8 4096 > $O/r2h_ncu_cross_sweep.log 2>&1; echo "ncu sweep rc=$?"
python tools/ncu_summary.py $O/prof_cross_sweep.ncu-rep > $O/r2h_ncu_summary_cross_sweep_kernel.txt 2>&1; rm -f $O/prof_cross_sweep.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:assemble_flat_kernel -c 1 -f -o $O/prof_flat \
    python tools/asm_bench.py > $O/r2h_ncu_flat.log 2>&1; echo "ncu flat rc=$?"
python tools/ncu_summary.py $O/prof_flat.ncu-rep > $O/r2h_ncu_summary_assemble_flat_kernel.txt 2>&1; rm -f $O/prof_flat.ncu-rep
