#!/usr/bin/env bash
# ncu evidence at the final kernel sources: rows kernel (fixed rho, full-occupancy launch), rows kernel under adaptive rho,
# whole-GPU kernel (config 4); launch lists of a bench step under both contracts.
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --mode frontier --no-adaptive-extra"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:admm_rows -s 52 -c 1 -f -o gpurun_out/r02f_rows_full74 $B > gpurun_out/r02f_ncu_rows.log 2>&1
tail -1 gpurun_out/r02f_ncu_rows.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:admm_rows -s 52 -c 1 -f -o gpurun_out/r02f_rows_adaptive_full74 $B --adaptive-rho-interval 50 > gpurun_out/r02f_ncu_rows_adaptive.log 2>&1
tail -1 gpurun_out/r02f_ncu_rows_adaptive.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:admm_grid -s 1 -c 1 -f -o gpurun_out/r02f_grid_cfg4 \
    python tools/iter_bench.py --instances 1 --n 2000 --m 4000 --p 200 --density 0.05 --iters 200 > gpurun_out/r02f_ncu_grid.log 2>&1
tail -1 gpurun_out/r02f_ncu_grid.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02f_launches_fixed.csv $B > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02f_launches_adaptive.csv $B --adaptive-rho-interval 50 > /dev/null 2>&1
ls -la gpurun_out/r02f_*
