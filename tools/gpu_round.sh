#!/bin/bash
# One GPU-box visit: the whole GPU suite (default build + the switches under evaluation), smoke, kernel timers, bench.
set -u
mkdir -p gpurun_out
tag=${1:-x}
echo "== default switches"
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_$tag.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_$tag.log | cut -c1-250 | tail -30
echo "== EMER_CHAIN_BWD=fused EMER_LINEAR_WGRAD=mn"
EMER_CHAIN_BWD=fused EMER_LINEAR_WGRAD=mn timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -m gpu -q > gpurun_out/pytest_new_$tag.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/pytest_new_$tag.log | cut -c1-250 | tail -20
EMER_CHAIN_BWD=fused EMER_LINEAR_WGRAD=mn timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 300 python tools/run_wgrad.py 2>&1 | tail -10
timeout 300 python tools/run_chain.py 524288 3 2>&1 | tail -6
echo "== bench: fused backward + FusedAdam (transposing weight gradients)"
EMER_CHAIN_BWD=fused timeout 900 python bench.py --steps 120 --warmup 12 --profile-all --optimizer fused --no-cpu-baseline --no-full-step > gpurun_out/bench_tc_$tag.log 2> gpurun_out/bench_tc_$tag.err
head -c 300 gpurun_out/bench_tc_$tag.log; echo
grep "^# emer\|^# library\|^# graph" gpurun_out/bench_tc_$tag.err | head -16
echo "== bench: + MN-major weight gradients"
EMER_CHAIN_BWD=fused EMER_LINEAR_WGRAD=mn timeout 900 python bench.py --steps 120 --warmup 12 --profile-all --optimizer fused > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
tail -c 1500 gpurun_out/bench_$tag.log
grep "^#" gpurun_out/bench_$tag.err | head -75
echo "== bench: + side-stream weight gradients"
EMER_WGRAD_STREAM=1 EMER_CHAIN_BWD=fused EMER_LINEAR_WGRAD=mn timeout 900 python bench.py --steps 120 --warmup 12 --optimizer fused --no-cpu-baseline --no-full-step > gpurun_out/bench_side_$tag.log 2> gpurun_out/bench_side_$tag.err
head -c 300 gpurun_out/bench_side_$tag.log; echo
