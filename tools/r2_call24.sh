#!/usr/bin/env bash
# Shared-memory-resident kernel for small problems (bqp_small.cu): parity tests, sanitizer passes, per-iteration timing against
# the direct-load kernel it replaces on a config-3-shaped problem, whole GPU suite, config 3 bench (1000 MPC steps).
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
timeout 600 python -m pytest tests/test_gpu_small_kernel.py -q -m gpu -x 2>&1 | tail -25 | tee gpurun_out/s35_small_tests.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_small_kernel.py -q -m gpu -k "mpc_program or mixed" 2>&1 | tail -8 | tee gpurun_out/s35_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_small_kernel.py -q -m gpu -k "mpc_program" 2>&1 | tail -8 | tee gpurun_out/s35_racecheck.log
for k in small direct; do
  if [ $k = direct ]; then export BQP_KERNEL=direct; else unset BQP_KERNEL; fi
  timeout 300 python tools/iter_bench.py --n 60 --m 90 --p 60 --density 0.05 --instances 16 --iters 2000 2>&1 | tail -2 | sed "s/^/$k: /" | tee -a gpurun_out/s35_iter_bench.log
  timeout 300 python tools/iter_bench.py --n 20 --m 50 --p 10 --density 1.0 --instances 49 --leaves 1 --iters 2000 2>&1 | tail -2 | sed "s/^/$k pickle-shape: /" | tee -a gpurun_out/s35_iter_bench.log
done
unset BQP_KERNEL
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/s35_tests.log
timeout 900 python bench.py --workload mpc > gpurun_out/s35_mpc.json 2> gpurun_out/s35_mpc.err; tail -c 1500 gpurun_out/s35_mpc.json
BQP_SMALL=0 timeout 900 python bench.py --workload mpc --no-cpu-baseline --mpc-steps 200 > gpurun_out/s35_mpc_direct_200.json 2>> gpurun_out/s35_mpc.err; tail -c 600 gpurun_out/s35_mpc_direct_200.json
timeout 900 python bench.py --workload mpc --no-cpu-baseline --mpc-steps 200 > gpurun_out/s35_mpc_small_200.json 2>> gpurun_out/s35_mpc.err; tail -c 600 gpurun_out/s35_mpc_small_200.json
