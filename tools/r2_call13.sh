#!/usr/bin/env bash
# Re-entry check of HEAD: whole GPU suite, smoke, default bench line.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/s25_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/s25_bench.json 2> gpurun_out/s25_bench.err; tail -c 3000 gpurun_out/s25_bench.json
