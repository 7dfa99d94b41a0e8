#!/bin/bash
# Final evidence of round 2 (1 GPU): GPU test suite, smoke, the driver's bench command, a short reference-arm run, and the
# ncu launch list of a short bench run (each ncu pass only after the same command has run plainly).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/final_smoke.log
timeout 1000 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/final_bench_n1.json 2> $O/final_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/final_bench_n1.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'steps', 'fits_per_s', 'gpu_launches', 'seconds_total')}, d['e2e']['value'])
print(d['roofline']['frac'], d['roofline']['whole_eval_frac'], d['roofline']['class_frac_algorithmic'], d['roofline']['traffic'])
for k in ('evals_configs3', 'roofline_prediction', 'posterior_grid', 'roofline_assembly', 'fit_reference_configs', 'fit_sample', 'cpu_baseline'):
    print(k, json.dumps(d.get(k))[:700])
PY
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_ref_short.json 2> $O/final_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/final_bench_ref_short.json
timeout 300 python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/final_plain_short.json 2>&1; echo "plain short rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/final_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/final_ncu_list.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(l for l in open('gpurun_out/final_launches.csv') if not l.startswith('==')) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    name = re.sub(r'<.*', '', r[ki])
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
unit = rows[1][hdr.index('Metric Unit')] if len(rows) > 1 else '?'
scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(unit, 1e-6)
tot = sum(a[1] for a in agg.values()) * scale
with open('gpurun_out/final_launches_summary.csv', 'w') as f:
    f.write(f"# {sum(a[0] for a in agg.values())} launches, {tot:.1f} ms total device time under ncu\nkernel,launches,total_ms,share\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{a[0]},{a[1] * scale:.3f},{a[1] * scale / tot:.4f}\n")
print(open('gpurun_out/final_launches_summary.csv').read()[:1500])
PY
rm -f $O/final_launches.csv
