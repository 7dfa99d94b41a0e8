#!/bin/bash
# round 2, GPU call D: the driver's N=1 command, the reference arm (short), launch list of one round under ncu
mkdir -p gpurun_out
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','steps','fits_per_s','gpu_launches')}, d['e2e']['value'], d['clocks'])
print(d['roofline']['whole_eval_frac'], d['roofline']['kernel'], d['roofline']['frac'], d['fit']['seconds_rounds'], d['fit']['seconds_moments'])
print({k: (v if isinstance(v,(str,float,int)) else '...') for k,v in d.items() if k in ('dgemm_cublas_tflops','fit_sample','cpu_baseline')})
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r2d_bench_ref.json | cut -c1-400
