#!/bin/bash
# Final visit: smoke + the driver's default bench command on the final build.
set -u
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err
tail -c 2500 gpurun_out/bench_final.log
tail -3 gpurun_out/bench_final.err | cut -c1-200
