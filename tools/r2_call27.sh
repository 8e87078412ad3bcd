#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_small_kernel.py tests/test_pickle_replay.py tests/test_mpc_power_converter.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/s38_small_tests.log
BQP_LIB_SUFFIX=_smalldbg BQP_BUILD_DEFS="-DBQP_SMALL_DEBUG" timeout 300 python tools/iter_bench.py --mpc --instances 1 --iters 2000 2>&1 | grep -v "^$" | tail -3 | tee gpurun_out/s38_small_phase_timers.log
timeout 300 python tools/iter_bench.py --mpc --instances 16 --iters 2000 2>&1 | tail -2 | tee -a gpurun_out/s38_small_phase_timers.log
timeout 300 python tools/iter_bench.py --n 20 --m 50 --p 10 --density 1.0 --instances 49 --leaves 1 --iters 2000 2>&1 | tail -1 | tee -a gpurun_out/s38_small_phase_timers.log
