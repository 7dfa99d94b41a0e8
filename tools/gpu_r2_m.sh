#!/bin/bash
# round 2, GPU call M: assembly with the table-driven exp / once-per-CTA constants / second-generation symmetric kernel:
# parity tests, A/B against the previous build (interleaved), write-only memset rate, ncu of both kernels
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "assemble or rbf_eval" > $O/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2m_pytest.log
for rep in 1 2; do
  unset GPBO_LIB GPBO_ASM_SYM_V1
  timeout 200 python tools/asm_quick.py new 2>/dev/null
  GPBO_ASM_SYM_V1=1 timeout 200 python tools/asm_quick.py new_symv1 2>/dev/null | grep -v "3200"
  GPBO_LIB=$PWD/tools/ab/libgpbo_head.so timeout 200 python tools/asm_quick.py head 2>/dev/null
done 2>&1 | tee $O/r2m_asm_ab.log
unset GPBO_LIB GPBO_ASM_SYM_V1
for k in assemble_sym2_kernel assemble_general_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $O/r2m_$k -f python tools/asm_quick.py ncu 8192 > $O/r2m_ncu_$k.log 2>&1
  python tools/ncu_summary.py $O/r2m_$k.ncu-rep > $O/r2m_ncu_summary_$k.txt 2>&1; head -30 $O/r2m_ncu_summary_$k.txt
done
