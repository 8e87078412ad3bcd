#!/usr/bin/env bash
# Final state of the round: whole GPU suite, smoke, config 3 bench with its CPU arm, default bench line, ncu of the
# shared-memory-resident kernel (full set, one launch of 16 tiles) and its launch list inside an MPC run.
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/s48_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -14 | tee gpurun_out/s48_smoke.log
timeout 900 python bench.py --workload mpc > gpurun_out/s48_mpc.json 2> gpurun_out/s48_mpc.err; tail -c 400 gpurun_out/s48_mpc.json
timeout 900 python bench.py > gpurun_out/s48_bench.json 2> gpurun_out/s48_bench.err; tail -c 300 gpurun_out/s48_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:admm_small -s 5 -c 1 -f -o gpurun_out/r02i_small_mpc_final python tools/iter_bench.py --mpc --instances 16 --iters 2000 > gpurun_out/r02i_ncu_small.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02i_launches_mpc.csv python bench.py --workload mpc --no-cpu-baseline --mpc-steps 20 > /dev/null 2>&1
ls -la gpurun_out/r02i_*
