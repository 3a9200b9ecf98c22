#!/bin/bash
# ncu captures of round 2 (run ON THE GPU BOX, one GPU):  gpurun --timeout 1800 -- 'bash tools/profile_round2.sh'
#   1. launch list of one eager training step (every kernel with its device time: compare SHARES)
#   2. --set full of the fused chain kernels and the weight-gradient kernels (tools/run_chain.py), and of the optimizer,
#      grid and proposal kernels inside one bench step
# Reports land in gpurun_out/; summaries are extracted on the build machine with tools/ncu_summary.py -> profiles/.
set -u
mkdir -p gpurun_out
export EMER_WGRAD_STREAM=0          # one stream: every kernel timed alone
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-full-step > gpurun_out/r2_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches.csv "Round 2: ncu launch list of bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-full-step (FusedAdam, fused chain, EMER_WGRAD_STREAM=0)" > gpurun_out/r2_launches_summary.md
head -45 gpurun_out/r2_launches_summary.md
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"field_fwd_kernel|field_bwd_kernel|wgrad_mn_kernel" \
    -s 7 -c 7 -o gpurun_out/r2_chain python tools/run_chain.py 524288 3 > gpurun_out/r2_chain.log 2>&1
tail -4 gpurun_out/r2_chain.log
timeout 900 ncu --set full --clock-control none -k regex:"adam_step_kernel|grid_fwd_kernel|grid_bwd_kernel|prop_level_kernel" -c 8 \
    -o gpurun_out/r2_misc python bench.py --steps 1 --warmup 0 --no-graph --no-e2e --no-cpu-baseline --no-full-step > gpurun_out/r2_misc.log 2>&1
ls -la gpurun_out/*.ncu-rep
