#!/bin/bash
# round 2, GPU call L (2 GPUs): the driver's N=2 command (strong scaling of the sharded fit) + configs[4]-shaped sample
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2l_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2l_bench_n2.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','steps','fits_per_s','gpu_launches','n_gpus')}, d['e2e']['value'])
print(d['roofline']['whole_eval_frac'], d['roofline']['peak'], d['fit']['seconds_rounds'], d['fit']['seconds_moments'])
print(json.dumps(d.get('config4_sample')))
PY
