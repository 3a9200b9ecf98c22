#!/bin/bash
# ncu --set full of the proposal-update kernels (first update step of an eager run).
set -u
mkdir -p gpurun_out
export EMER_WGRAD_STREAM=0
timeout 200 ncu --set full --clock-control none -k regex:"prop_level_bwd_kernel|interlevel_loss_kernel" -c 4 \
    -o gpurun_out/r2b_prop_update python bench.py --steps 3 --warmup 0 --no-graph --no-e2e --no-cpu-baseline --no-full-step > gpurun_out/r2b_prop_update.log 2>&1
ls -la gpurun_out/r2b_prop_update.ncu-rep
