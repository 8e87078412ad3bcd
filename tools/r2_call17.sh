#!/usr/bin/env bash
# Whole-GPU kernel after the latency work: parity, phase timers, config-4 timings.
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
timeout 600 python -m pytest tests -q -m gpu -x -k "grid or cfg4" 2>&1 | tail -8 | tee gpurun_out/s29_grid_tests.log
echo "=== phase timers (debug build)"
BQP_BUILD_DEFS="-DBQP_GRID_DEBUG" BQP_LIB_SUFFIX=_gd timeout 200 python tools/iter_bench.py --instances 1 --n 2000 --m 4000 --p 200 --density 0.05 --iters 200 2>&1 | grep -E "GRID|instances" | tail -4 | cut -c1-250
echo "=== release build"
timeout 200 python tools/iter_bench.py --instances 1 --n 2000 --m 4000 --p 200 --density 0.05 --iters 200 2>&1 | tail -2 | cut -c1-250
timeout 200 python tools/iter_bench.py --instances 1 --iters 200 2>&1 | tail -2 | cut -c1-250
BQP_KERNEL=grid BQP_GRID_ALL=1 timeout 200 python tools/iter_bench.py --instances 1 --iters 200 2>&1 | tail -2 | cut -c1-250
timeout 400 python bench.py --workload cfg4 --no-cpu-baseline 2>gpurun_out/s29_cfg4.err | tee gpurun_out/s29_cfg4.json | cut -c1-900
