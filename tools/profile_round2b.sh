#!/bin/bash
# ncu captures of the final round-2 build (run ON THE GPU BOX, one GPU):  gpurun --timeout 1800 -- 'bash tools/profile_round2b.sh'
#   1. launch list of two eager training steps (one without, one with proposal update): compare SHARES
#   2. --set full of the fused chain kernels and the MN-major weight-gradient kernels (tools/run_chain.py)
#   3. --set full of the proposal-update kernels, the narrow weight gradient, the optimizer, grid and proposal kernels
# Reports land in gpurun_out/; summaries are extracted on the build machine with tools/ncu_summary.py -> profiles/.
set -u
mkdir -p gpurun_out
export EMER_WGRAD_STREAM=0          # one stream: every kernel timed alone
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches.csv \
    python bench.py --steps 2 --warmup 0 --no-graph --no-e2e --no-cpu-baseline --no-full-step > gpurun_out/r2b_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r2b_launches.csv "Round 2, final build: ncu launch list of bench.py --steps 2 --warmup 0 --no-graph --no-e2e --no-cpu-baseline --no-full-step (two eager steps, the second with proposal update; EMER_WGRAD_STREAM=0; includes the parity leg's launches)" > gpurun_out/r2b_launches_summary.md
head -40 gpurun_out/r2b_launches_summary.md
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"field_fwd_kernel|field_bwd_kernel|wgrad_mn_kernel" \
    -s 7 -c 7 -o gpurun_out/r2b_chain python tools/run_chain.py 524288 3 > gpurun_out/r2b_chain.log 2>&1
tail -4 gpurun_out/r2b_chain.log
timeout 900 ncu --set full --clock-control none -k regex:"prop_level_bwd_kernel|interlevel_loss_kernel|narrow_wgrad_vec4_kernel|prop_level_kernel|grid_bwd_kernel" -c 12 \
    -o gpurun_out/r2b_misc python bench.py --steps 2 --warmup 0 --no-graph --no-e2e --no-cpu-baseline --no-full-step > gpurun_out/r2b_misc.log 2>&1
ls -la gpurun_out/r2b*.ncu-rep
