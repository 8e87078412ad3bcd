#!/usr/bin/env bash
# Round 2, GPU call 4: first run of the row-split cluster kernel (bqp_rows.cu): parity tests, then per-iteration latency
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -25
echo "--- all parity (no -x)"
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_zy_launch_invariance.py -q -m gpu 2>&1 | tail -15
echo "--- iter bench rows cs=2 / cs=1 / cs=4, panel"
timeout 200 python tools/iter_bench.py --instances 74 --iters 200 2>&1 | tail -2
BQP_ROWS_CLUSTER=1 timeout 200 python tools/iter_bench.py --instances 74 --iters 200 2>&1 | tail -2
BQP_ROWS_CLUSTER=4 timeout 200 python tools/iter_bench.py --instances 37 --iters 200 2>&1 | tail -2
