#!/bin/bash
# Short GPU-box visit: selected kernel tests + one bench line with the per-kernel tables.
#   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag> "<pytest -k expression>" [bench flags]'
set -u
mkdir -p gpurun_out
tag=${1:-x}; expr=${2:-interlevel}; shift 2 || true
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "$expr" > gpurun_out/pytest_$tag.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert " gpurun_out/pytest_$tag.log | cut -c1-250 | tail -12
timeout 600 python bench.py --steps 120 --warmup 12 --profile-all --no-cpu-baseline --no-full-step "$@" > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$tag.log").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"], 4), "rays/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["clocks"])
PY
grep "^# graph\|^# library\|^# ---" gpurun_out/bench_$tag.err
