#!/bin/bash
# One GPU-box visit: GPU suite, smoke, bench -- then the same bench with a compile-time switch flipped (A/B on one box).
#   bash tools/gpu_round_ab.sh <tag> <FLAG=value>
set -u
mkdir -p gpurun_out
tag=${1:-x}; flag=${2:-EMER_WARP_ARRIVE=0}
bash tools/gpu_round.sh $tag
echo "== B arm: $flag"
env $flag timeout 900 python bench.py --steps 120 --warmup 12 --profile-all --no-cpu-baseline --no-full-step > gpurun_out/bench_${tag}_B.log 2> gpurun_out/bench_${tag}_B.err
python - <<PY
import json
for t in ("$tag", "${tag}_B"):
    d = json.loads(open(f"gpurun_out/bench_{t}.log").read().strip().splitlines()[-1])
    print(t, "ms/step", round(d["ms_per_step"], 4), "rays/s", round(d["value"]))
PY
grep "^# emer_field\|^# emer_linear_tc_bwd_weight_mn\|^# graph" gpurun_out/bench_${tag}_B.err | head -8
