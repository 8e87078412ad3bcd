#!/usr/bin/env bash
# Bit-identity of rounds vs one launch per cluster size (adaptive on / off), adaptive B&B goldens on the engine.
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "adaptive_rho_rounds" 2>&1 | grep -E "^cluster|passed|failed" | tee gpurun_out/s32_rounds.log
timeout 900 python -m pytest tests/test_bnb_parity.py tests/test_zzb_contexts_sessions.py tests/test_zz_native_replay.py -q -m gpu 2>&1 | tail -8 | tee gpurun_out/s32_bnb.log
