#!/bin/bash
# ncu captures of round 2 (run ON THE GPU BOX, one GPU):  gpurun --timeout 1500 -- 'bash tools/profile_round2.sh'
#   1. launch list of one eager training step (every kernel with its device time: compare SHARES)
#   2. --set full of the fused chain kernels, the weight-gradient kernel, the optimizer and the grid kernels
# Reports land in gpurun_out/; summaries are extracted here with tools/ncu_summary.py and copied to profiles/.
set -u
mkdir -p gpurun_out
export EMER_CHAIN_BWD=${EMER_CHAIN_BWD:-fused}
OPT=${OPT:-fused}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-full-step --optimizer $OPT > gpurun_out/r2_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches.csv "Round 2: ncu launch list of bench.py --steps 1 --warmup 1 --no-graph (optimizer $OPT, EMER_CHAIN_BWD=$EMER_CHAIN_BWD)" > gpurun_out/r2_launches_summary.md
head -40 gpurun_out/r2_launches_summary.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"field_fwd_kernel|field_bwd_kernel|tc_wgrad_kernel|wgrad_mn_kernel" \
    -s 10 -c 8 -o gpurun_out/r2_chain python tools/run_chain.py 524288 3 > gpurun_out/r2_chain.log 2>&1
tail -8 gpurun_out/r2_chain.log
timeout 600 ncu --set full --clock-control none -k regex:"adam_step_kernel|grid_fwd_kernel|grid_bwd_kernel|prop_level_kernel" -s 30 -c 8 \
    -o gpurun_out/r2_misc python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-full-step --optimizer $OPT > gpurun_out/r2_misc.log 2>&1
ls -la gpurun_out/*.ncu-rep
