#!/bin/bash
# One GPU call: kernel-stage checks, then the whole `-m gpu` suite.  Logs only (small) under gpurun_out/<tag>_*.
set -u
tag=${1:-t}
out=gpurun_out
mkdir -p $out
python tools/gpu_check.py gemm > $out/${tag}_gemm.log 2>&1; echo "gemm stage rc=$?"; grep -c OK $out/${tag}_gemm.log; grep FAIL $out/${tag}_gemm.log | head -40
python -m pytest tests -m gpu -x -q -s > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log; grep -E "^\[parity|FAIL|Error" $out/${tag}_pytest.log | head -60
