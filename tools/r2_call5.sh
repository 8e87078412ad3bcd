#!/usr/bin/env bash
# Round 2, GPU call 5: rows kernel ring / prefetch sensitivity
set -u
for pf in 0 2 8; do echo "--- prefetch $pf"; BQP_ROWS_PREFETCH=$pf timeout 200 python tools/iter_bench.py --instances 74 --iters 200 2>&1 | tail -2; done
for sl in 2 3; do echo "--- slots $sl"; BQP_ROWS_SLOTS=$sl timeout 200 python tools/iter_bench.py --instances 74 --iters 200 2>&1 | tail -2; done
