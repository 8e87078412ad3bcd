#!/usr/bin/env bash
# Round 2, first GPU call: the tests owed from round 1 (marker gpu_next), config 3 Python vs native replay, bench with parallel setup.
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m "gpu or gpu_next" -q > gpurun_out/r2c1_tests.log 2>&1; echo "gpu tests rc=$?"; tail -15 gpurun_out/r2c1_tests.log
timeout 120 python tools/mpc_bench.py --steps 20 --horizon 10 --budgets 0,32,128 --warm > gpurun_out/r2c1_mpc_python.jsonl 2>&1
timeout 120 python tools/mpc_bench.py --steps 20 --horizon 10 --budgets 0,32,128 --warm --replay native > gpurun_out/r2c1_mpc_native.jsonl 2>&1
cat gpurun_out/r2c1_mpc_python.jsonl gpurun_out/r2c1_mpc_native.jsonl | cut -c1-400
timeout 400 python bench.py --parallel-setup > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; cut -c1-900 gpurun_out/r2c1_bench.json; tail -3 gpurun_out/r2c1_bench.err
