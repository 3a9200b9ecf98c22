#!/bin/bash
# One GPU-box visit: probes, the whole GPU suite (default build + the switches under evaluation), smoke, bench with the
# per-kernel table.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
tag=${1:-x}
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/mn_probe tools/mn_probe.cu && /tmp/mn_probe > gpurun_out/mn_probe_$tag.log 2>&1
cat gpurun_out/mn_probe_$tag.log
echo "== default switches"
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_$tag.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|pinned samples" gpurun_out/pytest_$tag.log | cut -c1-250 | tail -40
echo "== EMER_CHAIN_BWD=fused"
EMER_CHAIN_BWD=fused timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -m gpu -q -k "field_chain or full_size" > gpurun_out/pytest_fusedbwd_$tag.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_fusedbwd_$tag.log | cut -c1-250 | tail -20
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== bench: torch Adam, layer-wise chain backward"
timeout 900 python bench.py --steps 120 --warmup 12 --profile-all --no-full-step > gpurun_out/bench_base_$tag.log 2> gpurun_out/bench_base_$tag.err
tail -c 600 gpurun_out/bench_base_$tag.log | head -c 400; echo
grep "^#" gpurun_out/bench_base_$tag.err | head -45
echo "== bench: fused backward + FusedAdam"
EMER_CHAIN_BWD=fused timeout 900 python bench.py --steps 120 --warmup 12 --profile-all --optimizer fused > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
tail -c 3000 gpurun_out/bench_$tag.log
grep "^#" gpurun_out/bench_$tag.err | head -80
echo "== bench: + side-stream weight gradients"
EMER_WGRAD_STREAM=1 EMER_CHAIN_BWD=fused timeout 900 python bench.py --steps 120 --warmup 12 --optimizer fused --no-cpu-baseline --no-full-step > gpurun_out/bench_side_$tag.log 2> gpurun_out/bench_side_$tag.err
head -c 300 gpurun_out/bench_side_$tag.log; echo
