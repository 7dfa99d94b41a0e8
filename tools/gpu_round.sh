#!/bin/bash
# One GPU session: tests, bench, ncu launch list, ncu full captures of the top kernels (each only after its plain run).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
SMALL="--steps 1 --warmup 1 --no-cpu-baseline --no-fit-sample"
python bench.py $SMALL > $O/plain_small.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fit-sample > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"
for K in lauum_grad_kernel chol_panel_kernel trtri_row_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 1 -f -o $O/prof_$K \
      python bench.py --modes 28 --starts 37 $SMALL > $O/ncu_$K.log 2>&1; echo "ncu $K rc=$?"
done
python tools/asm_bench.py > $O/asm_bench.log 2>&1; echo "asm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:assemble_sym_kernel -c 1 -f -o $O/prof_assemble_sym \
    python tools/asm_bench.py > $O/ncu_asm.log 2>&1; echo "ncu asm rc=$?"
ls -la $O | head -40
