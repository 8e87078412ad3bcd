#!/usr/bin/env bash
# Multi-GPU measurements owed from round 1 (none was taken: the budget went into the kernels):
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/next_gpu_call_2gpu.sh 2'      (then 4, 8)
set -u
N=${1:-2}
mkdir -p gpurun_out
# weak scaling of the bench: every rank owns its own 100 instances, no data-path collective
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus "$N" --steps 5 --warmup 3 > "gpurun_out/n2_bench_${N}gpu.json" 2> "gpurun_out/n2_bench_${N}gpu.err"
cut -c1-500 "gpurun_out/n2_bench_${N}gpu.json"
# one MIQP's frontier (+ look-ahead) split over the GPUs: all-gather of node results + all-reduce(MIN) of the incumbent per B&B step
for K in 0 64; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29532 \
      examples/frontier_split.py --vars 500 --rows 1000 --ints 50 --density 0.7 --speculation "$K" --max-nodes 200 \
      > "gpurun_out/n2_frontier_${N}gpu_k${K}.json" 2> "gpurun_out/n2_frontier_${N}gpu_k${K}.err"
  cut -c1-400 "gpurun_out/n2_frontier_${N}gpu_k${K}.json"
done
