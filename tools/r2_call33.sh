#!/usr/bin/env bash
# Single-wait path of bqp_solve_multi + shared-memory-resident kernel: whole GPU suite, smoke, config 3 / config 4 benches with
# and without the fast path, default bench line.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/s44_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -14 | tee gpurun_out/s44_smoke.log
BQP_BNB_TIMERS=1 BQP_API_TIMERS=1 timeout 900 python bench.py --workload mpc > gpurun_out/s44_mpc.json 2> gpurun_out/s44_mpc_timers.err
tail -c 500 gpurun_out/s44_mpc.json; tail -3 gpurun_out/s44_mpc_timers.err
BQP_FAST_PATH=0 timeout 900 python bench.py --workload mpc --no-cpu-baseline > gpurun_out/s44_mpc_stepwise.json 2>> gpurun_out/s44_mpc_timers.err; python -c "import json;d=json.loads(open('gpurun_out/s44_mpc_stepwise.json').read().strip().splitlines()[-1]);print('stepwise path', d['value'])"
timeout 900 python bench.py --workload cfg4 > gpurun_out/s44_cfg4.json 2> gpurun_out/s44_cfg4.err; tail -c 400 gpurun_out/s44_cfg4.json
timeout 900 python bench.py > gpurun_out/s44_bench.json 2> gpurun_out/s44_bench.err; tail -c 300 gpurun_out/s44_bench.json
