#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/s51_tests.log
timeout 900 python bench.py > gpurun_out/s51_bench.json 2> gpurun_out/s51_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/s51_bench.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['adaptive_rho']['value'], d['bnb']['value'], d['clocks'])"
