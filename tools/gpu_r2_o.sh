#!/bin/bash
# round 2, GPU call O: assembly with the pre-pass kernel (constants + scaled abscissae), register caps; A/B of occupancy
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q -k "assemble or rbf_eval or call" > $O/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2o_pytest.log
for rep in 1 2; do
  unset GPBO_LIB
  timeout 200 python tools/asm_quick.py new 2>/dev/null
  GPBO_LIB=$PWD/tools/ab/libgpbo_alt.so timeout 200 python tools/asm_quick.py alt 2>/dev/null
  GPBO_LIB=$PWD/tools/ab/libgpbo_head.so timeout 200 python tools/asm_quick.py head 2>/dev/null
done 2>&1 | tee $O/r2o_asm_ab.log
unset GPBO_LIB
for k in assemble_general_kernel assemble_sym2_kernel; do
  timeout 300 ncu --set full --clock-control none -k regex:$k -s 1 -c 1 -o $O/r2o_$k -f python tools/asm_quick.py ncu 8192 > $O/r2o_ncu_$k.log 2>&1
  python tools/ncu_summary.py $O/r2o_$k.ncu-rep > $O/r2o_ncu_summary_$k.txt 2>&1; head -30 $O/r2o_ncu_summary_$k.txt
  rm -f $O/r2o_$k.ncu-rep
done
