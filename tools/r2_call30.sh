#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_small_kernel.py tests/test_pickle_replay.py tests/test_mpc_power_converter.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/s41_small_tests.log
BQP_LIB_SUFFIX=_smalldbg BQP_BUILD_DEFS="-DBQP_SMALL_DEBUG" timeout 300 python tools/iter_bench.py --mpc --instances 1 --iters 2000 2>&1 | grep -v "^$" | tail -3 | tee -a gpurun_out/s41_small_phase_timers.log
timeout 300 python tools/iter_bench.py --mpc --instances 16 --iters 2000 2>&1 | tail -1 | tee -a gpurun_out/s41_small_phase_timers.log
timeout 300 python tools/iter_bench.py --n 20 --m 50 --p 10 --density 1.0 --instances 49 --leaves 1 --iters 2000 2>&1 | tail -1 | tee -a gpurun_out/s41_small_phase_timers.log
timeout 300 python tools/iter_bench.py --n 60 --m 90 --p 60 --density 0.05 --instances 16 --iters 2000 2>&1 | tail -1 | tee -a gpurun_out/s41_small_phase_timers.log
BQP_BNB_TIMERS=1 BQP_API_TIMERS=1 timeout 900 python bench.py --workload mpc --no-cpu-baseline --mpc-steps 200 > gpurun_out/s41_mpc_small_200.json 2> gpurun_out/s41_mpc_timers.err
tail -c 300 gpurun_out/s41_mpc_small_200.json; tail -3 gpurun_out/s41_mpc_timers.err
