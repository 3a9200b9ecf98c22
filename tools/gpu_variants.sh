#!/bin/bash
# One GPU: the other BASELINE configurations (dynamic / flow / flow + feature head) and the single-GPU points of the
# strong-scaling configurations.
set -u
mkdir -p gpurun_out
tag=${1:-v}
for v in dynamic flow flow_feat; do
  timeout 600 python bench.py --variant $v --steps 40 --warmup 6 --no-cpu-baseline --no-full-step --profile-all > gpurun_out/bench_${v}_$tag.log 2> gpurun_out/bench_${v}_$tag.err
  echo "$v: $(tail -1 gpurun_out/bench_${v}_$tag.log | head -c 200)"; grep "^# emer\|^# library" gpurun_out/bench_${v}_$tag.err | head -12
done
timeout 600 python bench.py --variant flow --rays 16384 --steps 30 --warmup 6 --no-cpu-baseline --no-full-step > gpurun_out/bench_flow16k_$tag.log 2> gpurun_out/bench_flow16k_$tag.err
echo "flow 16384: $(tail -1 gpurun_out/bench_flow16k_$tag.log | head -c 200)"
