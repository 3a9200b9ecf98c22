#!/bin/bash
# A/B of the optimizer's group order (EMER_ADAM_EARLY): the training-trajectory and optimizer tests first, then two bench lines.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_kernels.py -m gpu -q -x -k "training_steps or fused_adam" > gpurun_out/pytest_adam.log 2>&1
rc=$?
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert " gpurun_out/pytest_adam.log | cut -c1-250 | tail -8
if [ $rc -ne 0 ]; then echo "guard failed"; tail -30 gpurun_out/pytest_adam.log | cut -c1-200; exit 1; fi
for arm in 1 0; do
  EMER_ADAM_EARLY=$arm timeout 400 python bench.py --steps 240 --warmup 12 --no-cpu-baseline --no-full-step > gpurun_out/bench_adam$arm.log 2> gpurun_out/bench_adam$arm.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_adam$arm.log").read().strip().splitlines()[-1])
print("EMER_ADAM_EARLY=$arm ms/step", round(d["ms_per_step"], 4), "rays/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "psnr", d.get("psnr_vs_reference"))
PY
done
