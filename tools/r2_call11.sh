#!/usr/bin/env bash
set -u
timeout 300 python -m pytest tests/test_mpc_power_converter.py -q -m gpu 2>&1 | tail -3
for s in 3 2 5; do echo "--- slots $s"; BQP_ROWS_SLOTS=$s timeout 120 python tools/iter_bench.py --instances 74 --iters 200 2>&1 | tail -2 | cut -c1-200; done
echo "--- slots 3 parity"; BQP_ROWS_SLOTS=3 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rows or cfg2_size_leaves or tile_widths" 2>&1 | tail -3
