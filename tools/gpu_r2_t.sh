#!/bin/bash
# round 2, GPU call T: posterior-grid tests, full GPU suite
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r2t_pytest.log; grep -E "^E |FAILED" $O/r2t_pytest.log | head -20
