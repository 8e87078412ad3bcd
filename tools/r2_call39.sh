#!/usr/bin/env bash
# Sanitizer passes over the final shared-memory-resident kernel; 2-GPU run of the default bench through torchrun.
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_small_kernel.py -q -m gpu -k "mpc_program or mixed or infeasible or tile_widths" 2>&1 | tail -5 | tee gpurun_out/s50_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_small_kernel.py -q -m gpu -k "mpc_program or mixed" 2>&1 | tail -5 | tee gpurun_out/s50_racecheck.log
timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_small_kernel.py -q -m gpu -k "mpc_program or mixed" 2>&1 | tail -5 | tee gpurun_out/s50_synccheck.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/s50_bench_2gpu.json 2> gpurun_out/s50_bench_2gpu.err; tail -c 400 gpurun_out/s50_bench_2gpu.json
