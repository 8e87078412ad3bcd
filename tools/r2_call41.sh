#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for v in "" _sel _sellu; do
  unset BQP_LIB_SUFFIX BQP_BUILD_DEFS
  if [ "$v" = "_sel" ]; then export BQP_LIB_SUFFIX=_sel BQP_BUILD_DEFS="-DBQP_SMALL_SEL"; fi
  if [ "$v" = "_sellu" ]; then export BQP_LIB_SUFFIX=_sellu BQP_BUILD_DEFS="-DBQP_SMALL_SEL -DBQP_SMALL_LU_REG"; fi
  timeout 300 python tools/iter_bench.py --mpc --instances 16 --iters 2000 2>&1 | tail -1 | sed "s/^/variant [$v]: /" | tee -a gpurun_out/s54_small_variants.log
  timeout 300 python tools/iter_bench.py --mpc --instances 16 --iters 2000 2>&1 | tail -1 | sed "s/^/variant [$v]: /" | tee -a gpurun_out/s54_small_variants.log
  timeout 600 python -m pytest tests/test_gpu_small_kernel.py tests/test_mpc_power_converter.py -q -m gpu -x 2>&1 | tail -2 | sed "s/^/variant [$v]: /" | tee -a gpurun_out/s54_small_variants.log
done
