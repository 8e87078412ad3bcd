#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -12
timeout 400 python bench.py --parallel-setup --no-cpu-baseline > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err; cut -c1-1500 gpurun_out/r2c8_bench.json; tail -3 gpurun_out/r2c8_bench.err
BQP_ROUND_ITERS=0 timeout 400 python bench.py --parallel-setup --no-cpu-baseline > gpurun_out/r2c8_bench_noround.json 2>> gpurun_out/r2c8_bench.err; cut -c1-400 gpurun_out/r2c8_bench_noround.json
