#!/usr/bin/env bash
# Last validation of the round at HEAD (small kernel without padding loads): GPU suite, smoke, config 3 bench, ncu of the small kernel.
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/s63_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2 | tee gpurun_out/s63_smoke.log
timeout 900 python bench.py --workload mpc > gpurun_out/s63_mpc.json 2> gpurun_out/s63_mpc.err
python -c "import json;d=json.loads(open('gpurun_out/s63_mpc.json').read().strip().splitlines()[-1]);print('mpc 1000 steps', d['value'], d['lookahead_128']['qp_per_s_consumed'], d['lookahead_32_first_steps']['ms_per_mpc_step'], d['cpu_baseline']['value'], d['gpu_over_cpu_on_the_same_steps'], d['cpu_baseline']['same_inputs_and_node_counts_as_gpu'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:admm_small -s 5 -c 1 -f -o gpurun_out/r02k_small_mpc_final python tools/iter_bench.py --mpc --instances 16 --iters 2000 > gpurun_out/r02k_ncu_small.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_small_kernel.py -q -m gpu -k "mpc_program or mixed or tile_widths" 2>&1 | tail -2 | tee gpurun_out/s63_memcheck.log
