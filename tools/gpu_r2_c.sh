#!/bin/bash
# round 2, GPU call C: full GPU tests, prediction / sqrtW / reference-config numbers after the fixes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --points 2048 --modes 16 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2c_bench_small.json 2> gpurun_out/r2c_bench_small.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench_small.json'))
print(json.dumps(d.get('roofline_prediction'))[:1500])
print(json.dumps(d.get('fit_reference_configs'))[:2500])
PY
