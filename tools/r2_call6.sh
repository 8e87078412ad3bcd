#!/usr/bin/env bash
set -u
BQP_LIB_SUFFIX=_rd BQP_BUILD_DEFS="-DBQP_ROWS_DEBUG" timeout 200 python tools/iter_bench.py --instances 1 --iters 100 2>&1 | grep -E "^PHASES" | sort | uniq -c | sort -rn | head -12
