#!/usr/bin/env bash
# Rows kernel split into a plain and an extended (eq_rho == 2 / adaptive rho) instantiation: GPU suite, default bench line,
# full-occupancy ncu capture of both instantiations, launch lists of a bench step under both contracts.
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/s34_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -12 | tee gpurun_out/s34_smoke.log
timeout 900 python bench.py > gpurun_out/s34_bench.json 2> gpurun_out/s34_bench.err; tail -c 600 gpurun_out/s34_bench.json
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --mode frontier --no-adaptive-extra"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:admm_rows -s 52 -c 1 -f -o gpurun_out/r02g_rows_full74 $B > gpurun_out/r02g_ncu_rows.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:admm_rows -s 7 -c 1 -f -o gpurun_out/r02g_rows_adaptive_full74 $B --adaptive-rho-interval 50 > gpurun_out/r02g_ncu_rows_adaptive.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02g_launches_fixed.csv $B > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02g_launches_adaptive.csv $B --adaptive-rho-interval 50 > /dev/null 2>&1
ls -la gpurun_out/r02g_* gpurun_out/s34_*
