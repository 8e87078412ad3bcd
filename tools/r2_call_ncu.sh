#!/usr/bin/env bash
# Round 2: ncu evidence for the rows kernel -- launch list of one bench step, one full capture of a full-occupancy launch
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:admm_rows -s 52 -c 1 -o gpurun_out/r02_rows_full74 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --mode frontier > gpurun_out/r02_ncu_full.log 2>&1
tail -2 gpurun_out/r02_ncu_full.log | cut -c1-300
ls -la gpurun_out/r02_rows_full74.ncu-rep
