#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err; tail -3 gpurun_out/r2c10_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c10_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')})
print(d['roofline']['frac'], d['cpu_baseline'])
print(json.dumps(d.get('bnb'))[:3000])
PY
timeout 300 python bench.py --impl reference 2>&1 | cut -c1-600
