#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_small_kernel.py tests/test_pickle_replay.py tests/test_mpc_power_converter.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/s43_small_tests.log
timeout 300 python tools/iter_bench.py --mpc --instances 16 --iters 2000 2>&1 | tail -2 | tee gpurun_out/s43_iter_bench.log
timeout 300 python tools/iter_bench.py --mpc --instances 16 --iters 25 2>&1 | tail -2 | tee -a gpurun_out/s43_iter_bench.log
BQP_BNB_TIMERS=1 BQP_API_TIMERS=1 timeout 900 python bench.py --workload mpc --no-cpu-baseline --mpc-steps 200 > gpurun_out/s43_mpc_small_200.json 2> gpurun_out/s43_mpc_timers.err
tail -c 300 gpurun_out/s43_mpc_small_200.json; tail -3 gpurun_out/s43_mpc_timers.err
