#!/bin/bash
# Short GPU-box visit while iterating on one kernel: its tests, its stand-alone timing, one bench line.
#   gpurun --timeout 900 -- 'bash tools/gpu_kernel_iter.sh <tag> "<pytest -k expression>"'
set -u
mkdir -p gpurun_out
tag=${1:-x}; expr=${2:-wgrad}
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "$expr" > gpurun_out/pytest_$tag.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/pytest_$tag.log | cut -c1-250 | tail -12
timeout 200 python tools/run_wgrad.py 524288 2>&1 | tail -8
timeout 600 python bench.py --steps 120 --warmup 12 --profile-all --no-cpu-baseline --no-full-step > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
tail -c 1500 gpurun_out/bench_$tag.log
grep "^# emer\|^# library\|^# graph" gpurun_out/bench_$tag.err | head -24
