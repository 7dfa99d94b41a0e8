#!/bin/bash
# One GPU session (outputs kept small: gpurun merges back at most 64 MiB, so .ncu-rep files are summarised and removed): tests, bench, ncu launch list, ncu full captures of the top kernels (each only after its plain run).
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
# one wave of 296 pairs (8 x 37): small enough for ncu's save/restore of the written buffers, large enough to fill the GPU
for KS in "chol_panel_kernel 20" "trtri_row_kernel 20" "chol_diag_kernel 40"; do
  set -- $KS
  ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o $O/prof_$1 \
      python bench.py --modes 8 --starts 37 $SMALL > $O/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
  python tools/ncu_summary.py $O/prof_$1.ncu-rep > $O/ncu_summary_$1.txt 2>&1; rm -f $O/prof_$1.ncu-rep
done
# the dominant kernel at the bench's own launch size (one full wave of 1036 pairs) for roofline.traffic
ncu --set full --clock-control none --import-source on -k regex:lauum_grad_kernel -s 1 -c 1 -f -o $O/prof_lauum_grad_kernel \
    python bench.py --modes 28 --starts 37 $SMALL > $O/ncu_lauum_grad_kernel.log 2>&1; echo "ncu lauum rc=$?"
python tools/ncu_summary.py $O/prof_lauum_grad_kernel.ncu-rep > $O/ncu_summary_lauum_grad_kernel.txt 2>&1; rm -f $O/prof_lauum_grad_kernel.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:ns_gemm_kernel -s 30 -c 1 -f -o $O/prof_ns_gemm_kernel \
    python bench.py --modes 8 --starts 4 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_ns_gemm.log 2>&1; echo "ncu ns rc=$?"
python tools/ncu_summary.py $O/prof_ns_gemm_kernel.ncu-rep > $O/ncu_summary_ns_gemm_kernel.txt 2>&1; rm -f $O/prof_ns_gemm_kernel.ncu-rep
python tools/asm_bench.py > $O/asm_bench.log 2>&1; echo "asm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:assemble_sym_kernel -c 1 -f -o $O/prof_assemble_sym \
    python tools/asm_bench.py > $O/ncu_asm.log 2>&1; echo "ncu asm rc=$?"
python tools/ncu_summary.py $O/prof_assemble_sym.ncu-rep > $O/ncu_summary_assemble_sym.txt 2>&1; rm -f $O/prof_assemble_sym.ncu-rep
ls -la $O | head -40
