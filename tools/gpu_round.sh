#!/bin/bash
# One GPU-box visit: probes, the whole GPU suite, smoke, bench with the per-kernel table.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
tag=${1:-x}
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/mn_probe tools/mn_probe.cu && /tmp/mn_probe > gpurun_out/mn_probe_$tag.log 2>&1
cat gpurun_out/mn_probe_$tag.log
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_$tag.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|pinned samples" gpurun_out/pytest_$tag.log | cut -c1-250 | tail -40
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 120 --warmup 12 --profile-all > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
tail -c 2500 gpurun_out/bench_$tag.log
grep "^#" gpurun_out/bench_$tag.err | head -80
