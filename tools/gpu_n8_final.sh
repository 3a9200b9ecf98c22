#!/bin/bash
# Final multi-GPU visit (gpurun --gpus 8): the 2-rank NCCL tests, then the weak-scaling bench at N = 8 (sharded exchange).
set -u
mkdir -p gpurun_out
tag=${1:-f}
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 8 --steps 120 --warmup 12 --no-cpu-baseline --no-full-step > gpurun_out/bench_n8_sharded_$tag.log 2> gpurun_out/bench_n8_sharded_$tag.err
tail -1 gpurun_out/bench_n8_sharded_$tag.log | head -c 400; echo
grep -i "error\|Traceback" gpurun_out/bench_n8_sharded_$tag.err | head -5
