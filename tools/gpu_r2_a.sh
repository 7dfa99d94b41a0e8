#!/bin/bash
# round 2, GPU call A: tests + bench sanity runs (results under gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2a_pytest.log
timeout 400 python bench.py --points 2048 --modes 16 --steps 4 --warmup 1 > gpurun_out/r2a_bench_small.json 2> gpurun_out/r2a_bench_small.err; echo "bench small rc=$?"
timeout 500 python bench.py --steps 4 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2a_bench_8192_k4.json 2> gpurun_out/r2a_bench_8192_k4.err; echo "bench 8192 rc=$?"
tail -c 1500 gpurun_out/r2a_bench_8192_k4.json
