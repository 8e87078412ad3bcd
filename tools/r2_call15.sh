#!/usr/bin/env bash
# Work queue in the rows kernel + the whole-GPU kernel (bqp_grid.cu): parity suites, frontier step variants, config-4 timings.
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
timeout 900 python -m pytest tests -q -m gpu -x -k "not grid and not cfg4" 2>&1 | tail -8 | tee gpurun_out/s27_tests.log
for inst in 74 100; do timeout 120 python tools/iter_bench.py --instances $inst --iters 200 2>&1 | tail -1 | cut -c1-250; done
run() { echo "--- $*"; env "$@" timeout 300 python bench.py --mode frontier --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/s27_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({k: d.get(k) for k in ('value', 'ms_per_step', 'admm_node_iters_per_s', 'gpu_launches')}), json.dumps({k: d['roofline'].get(k) for k in ('frac', 'launches_per_step', 'streamed_gbs')}), 'e2e', d['e2e']['value'])" || tail -3 gpurun_out/s27_bench.err; }
run BQP_ROWS_QUEUE=1
run BQP_ROWS_QUEUE=0
run BQP_ROWS_PROBES=1
run BQP_ROWS_AUTO_CLUSTER=1
echo "=== grid kernel"
timeout 600 python -m pytest tests -q -m gpu -x -k "grid or cfg4" 2>&1 | tail -25 | tee gpurun_out/s27_grid_tests.log
timeout 200 python tools/iter_bench.py --instances 1 --n 2000 --m 4000 --p 200 --density 0.05 --iters 200 2>&1 | tail -2 | cut -c1-250
BQP_GRID=0 timeout 200 python tools/iter_bench.py --instances 1 --n 2000 --m 4000 --p 200 --density 0.05 --iters 100 2>&1 | tail -2 | cut -c1-250
timeout 400 python bench.py --workload cfg4 --no-cpu-baseline 2>gpurun_out/s27_cfg4.err | tee gpurun_out/s27_cfg4.json | cut -c1-1500
