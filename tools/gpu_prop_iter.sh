#!/bin/bash
# GPU-box visit for the fused proposal-level backward: its kernel tests, the full-size parity file, one bench line.
set -u
mkdir -p gpurun_out
tag=${1:-x}
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "proposal or prop_level" > gpurun_out/pytest_$tag.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/pytest_$tag.log | cut -c1-300 | tail -14
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest_full_$tag.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/pytest_full_$tag.log | cut -c1-300 | tail -10
timeout 600 python bench.py --steps 120 --warmup 12 --profile-all --no-cpu-baseline --no-full-step > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
tail -c 600 gpurun_out/bench_$tag.log
grep "^# emer\|^# library\|^# graph" gpurun_out/bench_$tag.err | head -30
