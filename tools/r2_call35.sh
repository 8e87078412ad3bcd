#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for v in "" _mfsmem; do
  if [ -n "$v" ]; then export BQP_LIB_SUFFIX=$v BQP_BUILD_DEFS="-DBQP_SMALL_MF_SMEM"; fi
  timeout 300 python tools/iter_bench.py --mpc --instances 16 --iters 2000 2>&1 | tail -1 | sed "s/^/variant [$v]: /" | tee -a gpurun_out/s46_small_variants.log
  timeout 300 python tools/iter_bench.py --mpc --instances 16 --iters 2000 2>&1 | tail -1 | sed "s/^/variant [$v]: /" | tee -a gpurun_out/s46_small_variants.log
done
