#!/usr/bin/env bash
# Round 2, multi-GPU call: config 4 (one n=2000 MIQP, frontier + look-ahead split over the ranks, NCCL) and the weak-scaling bench
set -u
N=${1:-2}
mkdir -p gpurun_out
for K in 0 64; do
  timeout 400 python examples/frontier_split.py --vars 2000 --rows 4000 --ints 200 --density 0.05 --speculation $K --max-nodes 40 \
      > "gpurun_out/r2_cfg4_1gpu_k${K}.json" 2> "gpurun_out/r2_cfg4_1gpu_k${K}.err"; cut -c1-900 "gpurun_out/r2_cfg4_1gpu_k${K}.json"; tail -2 "gpurun_out/r2_cfg4_1gpu_k${K}.err"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29532 \
      examples/frontier_split.py --vars 2000 --rows 4000 --ints 200 --density 0.05 --speculation $K --max-nodes 40 --dist-backend nccl \
      > "gpurun_out/r2_cfg4_${N}gpu_k${K}.json" 2> "gpurun_out/r2_cfg4_${N}gpu_k${K}.err"; cut -c1-900 "gpurun_out/r2_cfg4_${N}gpu_k${K}.json"; grep -v "^W\|^\*\*\*\|OMP_NUM" "gpurun_out/r2_cfg4_${N}gpu_k${K}.err" | tail -3
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus "$N" --steps 5 --warmup 3 --no-cpu-baseline > "gpurun_out/r2_bench_${N}gpu.json" 2> "gpurun_out/r2_bench_${N}gpu.err"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','e2e')}, d['roofline']['frac'])
b=d.get('bnb') or {}
print({k:b.get(k) for k in ('value',)}, b.get('rolling'), b.get('lockstep'))
PY
