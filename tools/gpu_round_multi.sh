#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL test of the data-parallel step, bench in both exchange modes and the torch arm.
set -u
mkdir -p gpurun_out
N=${1:-2}; tag=${2:-m}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 60 --warmup 8 --no-cpu-baseline "${@:2}"; }
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q 2>&1 | tail -5
fi
for mode in sharded allreduce; do
  timeout 600 bash -c "$(declare -f run); N=$N run 295$((RANDOM % 90 + 10)) --dp-mode $mode" > gpurun_out/bench_n${N}_${mode}_$tag.log 2> gpurun_out/bench_n${N}_${mode}_$tag.err
  echo "$mode: $(head -c 260 gpurun_out/bench_n${N}_${mode}_$tag.log)"; tail -3 gpurun_out/bench_n${N}_${mode}_$tag.err | cut -c1-200
done
timeout 600 bash -c "$(declare -f run); N=$N run 296$((RANDOM % 90 + 10)) --optimizer torch" > gpurun_out/bench_n${N}_torch_$tag.log 2> gpurun_out/bench_n${N}_torch_$tag.err
echo "torch arm: $(head -c 260 gpurun_out/bench_n${N}_torch_$tag.log)"
timeout 600 bash -c "$(declare -f run); N=$N run 297$((RANDOM % 90 + 10)) --dp-mode sharded --no-defer-gather" > gpurun_out/bench_n${N}_nodefer_$tag.log 2> gpurun_out/bench_n${N}_nodefer_$tag.err
echo "sharded, gather not deferred: $(head -c 260 gpurun_out/bench_n${N}_nodefer_$tag.log)"
if [ "$N" = "8" ]; then
  timeout 600 bash -c "$(declare -f run); N=$N run 298$((RANDOM % 90 + 10)) --scaling strong --rays 16384 --variant flow" > gpurun_out/bench_n${N}_strong_flow_$tag.log 2> gpurun_out/bench_n${N}_strong_flow_$tag.err
  echo "strong flow 16384: $(head -c 300 gpurun_out/bench_n${N}_strong_flow_$tag.log)"
fi
