#!/usr/bin/env bash
set -u
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_zy_launch_invariance.py -q -m gpu 2>&1 | tail -9
timeout 200 python tools/iter_bench.py --instances 74 --iters 200 2>&1 | tail -2
BQP_ROWS_CLUSTER=4 timeout 200 python tools/iter_bench.py --instances 33 --iters 200 2>&1 | tail -2
BQP_ROWS_CLUSTER=1 timeout 200 python tools/iter_bench.py --instances 1 --iters 200 2>&1 | tail -1
BQP_LIB_SUFFIX=_rd BQP_BUILD_DEFS="-DBQP_ROWS_DEBUG" timeout 200 python tools/iter_bench.py --instances 1 --iters 100 2>&1 | grep -E "^PHASES|^ROWS" | sort | uniq -c | sort -rn | awk '$1>=1' | head -8
