#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/iter_bench.py --mpc --instances 16 --iters 2000 2>&1 | tail -2 | tee gpurun_out/s42_iter_bench.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/s42_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -14 | tee gpurun_out/s42_smoke.log
BQP_BNB_TIMERS=1 BQP_API_TIMERS=1 timeout 900 python bench.py --workload mpc > gpurun_out/s42_mpc.json 2> gpurun_out/s42_mpc_timers.err
tail -c 700 gpurun_out/s42_mpc.json; tail -3 gpurun_out/s42_mpc_timers.err
