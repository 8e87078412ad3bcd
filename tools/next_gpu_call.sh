#!/usr/bin/env bash
# First GPU call of the next round (everything below was written or changed after round 1's GPU budget was spent):
#   gpurun --timeout 900 -- 'bash tools/next_gpu_call.sh'
# Writes logs under gpurun_out/; copy what is worth keeping into profiles/.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m "gpu or gpu_next" -q > gpurun_out/n1_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/n1_tests.log
# config 3: Python vs native replay, look-ahead 0 / 32 / 128, warmed up
timeout 120 python tools/mpc_bench.py --steps 20 --horizon 10 --budgets 0,32,128 --warm > gpurun_out/n1_mpc_python.jsonl 2>&1
timeout 120 python tools/mpc_bench.py --steps 20 --horizon 10 --budgets 0,32,128 --warm --replay native > gpurun_out/n1_mpc_native.jsonl 2>&1
cat gpurun_out/n1_mpc_python.jsonl gpurun_out/n1_mpc_native.jsonl | cut -c1-400
# launch invariance of all three kernels on a small problem (the race fixed in round 1)
timeout 180 python tools/batch_invariance.py --sweep > gpurun_out/n1_invariance.log 2>&1; grep -E "^(====|panel|stream|direct|   node)" gpurun_out/n1_invariance.log | cut -c1-300
# the bench line after the barrier added to the panel kernel (expected: unchanged, 431 ms per step)
timeout 400 python bench.py > gpurun_out/n1_bench.json 2> gpurun_out/n1_bench.err; cut -c1-600 gpurun_out/n1_bench.json
# parallel host setup path (bqp_setup_many): same bench line, shorter untimed setup
timeout 400 python bench.py --parallel-setup --no-cpu-baseline > gpurun_out/n1_bench_parallel_setup.json 2>> gpurun_out/n1_bench.err; cut -c1-300 gpurun_out/n1_bench_parallel_setup.json
