#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:admm_small -s 2 -c 1 -f -o gpurun_out/r02h_small_mpc python tools/iter_bench.py --mpc --instances 16 --iters 2000 > gpurun_out/r02h_ncu_small.log 2>&1
tail -2 gpurun_out/r02h_ncu_small.log | cut -c1-200
ls -la gpurun_out/r02h_*
