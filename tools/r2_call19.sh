#!/usr/bin/env bash
# Adaptive rho on the rows kernel (second barrier for the V pass) + whole suite + default bench line with the adaptive extra.
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee gpurun_out/s31_tests.log
for k in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "adaptive_rho_rounds" 2>&1 | tail -1; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/s31_bench.json 2> gpurun_out/s31_bench.err; tail -c 2500 gpurun_out/s31_bench.json
