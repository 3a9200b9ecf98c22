#!/bin/bash
# One GPU-box visit: the GPU suite under the candidate defaults, smoke, bench.
set -u
mkdir -p gpurun_out
tag=${1:-x}
export EMER_CHAIN_BWD=fused EMER_LINEAR_WGRAD=mn
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_$tag.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/pytest_$tag.log | cut -c1-250 | tail -30
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | cut -c1-400
echo "== bench: fused backward + FusedAdam + MN-major weight gradients + side stream"
EMER_WGRAD_STREAM=1 timeout 900 python bench.py --steps 120 --warmup 12 --profile-all --optimizer fused > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
tail -c 1800 gpurun_out/bench_$tag.log
grep "^# emer\|^# library\|^# graph" gpurun_out/bench_$tag.err | head -24
