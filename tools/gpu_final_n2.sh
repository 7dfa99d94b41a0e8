#!/bin/bash
# Final multi-GPU sanity run of round 2 (2 GPUs): the driver's torchrun command with fewer rounds, configs[4] sample skipped
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 4 --warmup 2 --no-config4 > $O/final_bench_n2_short.json 2> $O/final_bench_n2.err; echo "bench n2 rc=$?"
tail -3 $O/final_bench_n2.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/final_bench_n2_short.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'steps', 'fits_per_s', 'gpu_launches', 'n_gpus', 'seconds_total')}, d['e2e']['value'])
print(d['roofline']['whole_eval_frac'], d['config']['live_pairs_per_round'], d['fit'])
PY
