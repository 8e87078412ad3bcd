#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_small_kernel.py tests/test_mpc_power_converter.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/s47_small_tests.log
timeout 300 python tools/iter_bench.py --mpc --instances 16 --iters 2000 2>&1 | tail -1 | tee gpurun_out/s47_iter_bench.log
timeout 300 python tools/iter_bench.py --mpc --instances 16 --iters 2000 2>&1 | tail -1 | tee -a gpurun_out/s47_iter_bench.log
BQP_BNB_TIMERS=1 BQP_API_TIMERS=1 timeout 900 python bench.py --workload mpc --no-cpu-baseline > gpurun_out/s47_mpc.json 2> gpurun_out/s47_mpc_timers.err
python -c "import json;d=json.loads(open('gpurun_out/s47_mpc.json').read().strip().splitlines()[-1]);print('mpc 1000 steps', d['value'], d['lookahead_32_first_steps']['ms_per_mpc_step'])"; tail -3 gpurun_out/s47_mpc_timers.err
