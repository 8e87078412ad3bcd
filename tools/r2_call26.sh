#!/usr/bin/env bash
# Phase timers of the shared-memory-resident kernel on the config-3 problem (debug build), per-iteration time of the product build.
set -u
mkdir -p gpurun_out
BQP_LIB_SUFFIX=_smalldbg BQP_BUILD_DEFS="-DBQP_SMALL_DEBUG" timeout 300 python tools/iter_bench.py --mpc --instances 16 --iters 2000 2>&1 | grep -v "^$" | tail -12 | tee gpurun_out/s37_small_phase_timers.log
timeout 300 python tools/iter_bench.py --mpc --instances 16 --iters 2000 2>&1 | tail -2 | tee -a gpurun_out/s37_small_phase_timers.log
BQP_KERNEL=direct timeout 300 python tools/iter_bench.py --mpc --instances 16 --iters 2000 2>&1 | tail -2 | sed "s/^/direct: /" | tee -a gpurun_out/s37_small_phase_timers.log
