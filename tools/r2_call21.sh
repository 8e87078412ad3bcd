#!/usr/bin/env bash
# Canonical partial slots (bit-identity of rounds at every cluster size, adaptive on / off) + whole GPU suite.
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "adaptive_rho_rounds" > gpurun_out/s33_rounds.log 2>&1; grep -E "cluster [0-9]|passed|failed" gpurun_out/s33_rounds.log | cut -c1-200
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee gpurun_out/s33_tests.log
