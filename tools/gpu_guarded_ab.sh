#!/bin/bash
# GPU-box visit for a risky kernel change: first ONLY that kernel's tests under a short timeout (a hang must not eat
# the visit), then the usual round + a B arm with a switch flipped.
#   bash tools/gpu_guarded_ab.sh <tag> "<pytest -k expr>" <FLAG=value>
set -u
mkdir -p gpurun_out
tag=${1:-x}; expr=${2:-field_chain}; flag=${3:-EMER_CHAIN_STAGE=0}
timeout 180 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$expr" > gpurun_out/pytest_${tag}_guard.log 2>&1
rc=$?
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert " gpurun_out/pytest_${tag}_guard.log | cut -c1-250 | tail -12
if [ $rc -ne 0 ]; then echo "guard tests failed (rc=$rc): stopping"; tail -30 gpurun_out/pytest_${tag}_guard.log | cut -c1-200; exit 1; fi
bash tools/gpu_round_ab.sh $tag $flag
