#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL test of the data-parallel step (N = 2), bench in both exchange modes, the torch arm,
# and (N = 8) the strong-scaling flow configuration of BASELINE.json configs[3].
set -u
mkdir -p gpurun_out
N=${1:-2}; tag=${2:-m}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 60 --warmup 8 --no-cpu-baseline "${@:2}"; }
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q 2>&1 | tail -5
fi
for mode in sharded allreduce; do
  timeout 600 bash -c "$(declare -f run); N=$N run 295$((RANDOM % 90 + 10)) --dp-mode $mode" > gpurun_out/bench_n${N}_${mode}_$tag.log 2> gpurun_out/bench_n${N}_${mode}_$tag.err
  echo "$mode: $(tail -1 gpurun_out/bench_n${N}_${mode}_$tag.log | head -c 230)"
done
timeout 600 bash -c "$(declare -f run); N=$N run 296$((RANDOM % 90 + 10)) --optimizer torch" > gpurun_out/bench_n${N}_torch_$tag.log 2> gpurun_out/bench_n${N}_torch_$tag.err
echo "torch arm: $(tail -1 gpurun_out/bench_n${N}_torch_$tag.log | head -c 230)"
if [ "$N" = "8" ]; then
  timeout 600 bash -c "$(declare -f run); N=$N run 298$((RANDOM % 90 + 10)) --scaling strong --rays 16384 --variant flow" > gpurun_out/bench_n${N}_strong_flow_$tag.log 2> gpurun_out/bench_n${N}_strong_flow_$tag.err
  echo "strong flow 16384: $(tail -1 gpurun_out/bench_n${N}_strong_flow_$tag.log | head -c 300)"; tail -3 gpurun_out/bench_n${N}_strong_flow_$tag.err | cut -c1-200
  timeout 600 bash -c "$(declare -f run); N=$N run 299$((RANDOM % 90 + 10)) --scaling strong --rays 8192 --variant flow_feat" > gpurun_out/bench_n${N}_strong_flowfeat_$tag.log 2> gpurun_out/bench_n${N}_strong_flowfeat_$tag.err
  echo "strong flow_feat 8192: $(tail -1 gpurun_out/bench_n${N}_strong_flowfeat_$tag.log | head -c 300)"
fi
