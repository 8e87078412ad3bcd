#!/usr/bin/env bash
# Adaptive rho on the whole-GPU kernel: parity against the oracle, single-tree timings.
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
timeout 900 python -m pytest tests -q -m gpu -x -k "grid or cfg4 or adaptive" 2>&1 | tail -25 | tee gpurun_out/s30_tests.log
timeout 200 python tools/iter_bench.py --instances 1 --n 2000 --m 4000 --p 200 --density 0.05 --iters 200 2>&1 | tail -1 | cut -c1-250
timeout 600 python tools/adaptive_bench.py --instances 3 2>&1 | tee gpurun_out/s30_adaptive_cfg2.jsonl | cut -c1-400
timeout 600 python tools/adaptive_bench.py --cfg4 2>&1 | tee gpurun_out/s30_adaptive_cfg4.jsonl | cut -c1-400
