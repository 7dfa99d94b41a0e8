#!/bin/bash
# Final refresh of the round's evidence for the kernels that changed after the full round (tools/gpu_round.sh):
# tests, bench (both arms), ncu launch list, ncu captures of chol_diag and the assembly kernels.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fit-sample > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"
SMALL="--steps 1 --warmup 1 --no-cpu-baseline --no-fit-sample"
ncu --set full --clock-control none --import-source on -k regex:chol_diag_kernel -s 40 -c 1 -f -o $O/prof_chol_diag_kernel \
    python bench.py --modes 8 --starts 37 $SMALL > $O/ncu_chol_diag_kernel.log 2>&1; echo "ncu diag rc=$?"
python tools/ncu_summary.py $O/prof_chol_diag_kernel.ncu-rep > $O/ncu_summary_chol_diag_kernel.txt 2>&1; rm -f $O/prof_chol_diag_kernel.ncu-rep
python tools/asm_bench.py > $O/asm_bench.log 2>&1; echo "asm rc=$?"
for K in assemble_sym_kernel assemble_general_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -f -o $O/prof_$K python tools/asm_bench.py > $O/ncu_$K.log 2>&1; echo "ncu $K rc=$?"
  python tools/ncu_summary.py $O/prof_$K.ncu-rep > $O/ncu_summary_$K.txt 2>&1; rm -f $O/prof_$K.ncu-rep
done
python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
