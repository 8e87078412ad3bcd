#!/usr/bin/env bash
# Work queue in the rows kernel: parity suite, then the frontier step with / without the queue, probes and automatic cluster size.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/s26_tests.log
for inst in 74 100; do timeout 120 python tools/iter_bench.py --instances $inst --iters 200 2>&1 | tail -1 | cut -c1-250; done
run() { echo "--- $*"; env "$@" timeout 300 python bench.py --mode frontier --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({k: d.get(k) for k in ('value', 'ms_per_step', 'admm_node_iters_per_s', 'gpu_launches')}), json.dumps({k: d['roofline'].get(k) for k in ('frac', 'launches_per_step', 'streamed_gbs')}), 'e2e', d['e2e']['value'])"; }
run BQP_ROWS_QUEUE=1
run BQP_ROWS_QUEUE=0
run BQP_ROWS_PROBES=1
run BQP_ROWS_PROBES=3
run BQP_ROWS_AUTO_CLUSTER=1
run BQP_ROWS_AUTO_CLUSTER=1 BQP_ROWS_QUEUE=0
