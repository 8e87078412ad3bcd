#!/usr/bin/env bash
# Round 2, GPU call 2: A/B of pass-1 warp sets / tiles per pass-1 warp in the panel kernel (per-iteration latency, 1 tile and a full wave)
set -u
mkdir -p gpurun_out
run() { # suffix sets tiles [extra env]
  BQP_LIB_SUFFIX=_v$1 BQP_BUILD_DEFS="-DBQP_P1_SETS=$2 -DBQP_P1_TILES=$3" timeout 300 python tools/iter_bench.py --instances 74 --iters 200 2>&1 | grep -v "^$" | tail -3
}
run 0 1 1
run 1 2 2
run 2 1 2
run 3 2 1
echo "--- prefetch 4 on v0 / v1"
BQP_PANEL_PREFETCH=4 run 0 1 1
BQP_PANEL_PREFETCH=4 run 1 2 2
echo "--- parity v1"
BQP_LIB_SUFFIX=_v1 BQP_BUILD_DEFS="-DBQP_P1_SETS=2 -DBQP_P1_TILES=2" timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_zy_launch_invariance.py -q -m gpu 2>&1 | tail -5
