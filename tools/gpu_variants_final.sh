#!/bin/bash
# One GPU, final build: the other BASELINE configurations (dynamic / flow / flow + feature head).
set -u
mkdir -p gpurun_out
for v in dynamic flow flow_feat; do
  timeout 400 python bench.py --variant $v --steps 40 --warmup 6 --no-cpu-baseline --no-full-step > gpurun_out/bench_${v}_final.log 2> gpurun_out/bench_${v}_final.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${v}_final.log").read().strip().splitlines()[-1])
    print("$v", "ms/step", round(d["ms_per_step"], 3), "rays/s", round(d["value"]), "psnr", d.get("psnr_vs_reference"), "max_rel_err", d.get("max_rel_err"))
except Exception as e:
    print("$v failed", e)
PY
  tail -2 gpurun_out/bench_${v}_final.err | cut -c1-200
done
