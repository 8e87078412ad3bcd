#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
timeout 900 python -m pytest tests/test_mpc_power_converter.py tests/test_zz_native_replay.py tests/test_zzz_cfg4_bnb.py tests/test_bnb_parity.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/s49_tests.log
BQP_BNB_TIMERS=1 BQP_API_TIMERS=1 timeout 900 python bench.py --workload mpc > gpurun_out/s49_mpc.json 2> gpurun_out/s49_mpc_timers.err
python -c "import json;d=json.loads(open('gpurun_out/s49_mpc.json').read().strip().splitlines()[-1]);print('mpc 1000 steps', d['value'], d['lookahead_32_first_steps']['ms_per_mpc_step'], d['gpu_over_cpu_on_the_same_steps'], d['cpu_baseline']['same_inputs_and_node_counts_as_gpu'])"; tail -3 gpurun_out/s49_mpc_timers.err
