#!/usr/bin/env bash
# bqp_small.cu with chunked ELL products; where the wall time of a config-3 step goes (host timers).
set -u
mkdir -p gpurun_out
python -c "import miosqp_b200.build as b; assert not b._stale(), 'libbqp.so is stale'" || exit 1
timeout 600 python -m pytest tests/test_gpu_small_kernel.py tests/test_pickle_replay.py tests/test_mpc_power_converter.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/s36_small_tests.log
for d in 0.05 0.034; do
  timeout 300 python tools/iter_bench.py --n 60 --m 90 --p 60 --density $d --instances 16 --iters 2000 2>&1 | tail -1 | sed "s/^/small d=$d: /" | tee -a gpurun_out/s36_iter_bench.log
done
BQP_BNB_TIMERS=1 BQP_API_TIMERS=1 timeout 900 python bench.py --workload mpc --no-cpu-baseline --mpc-steps 200 > gpurun_out/s36_mpc_small_200.json 2> gpurun_out/s36_mpc_timers.err
tail -c 400 gpurun_out/s36_mpc_small_200.json; tail -4 gpurun_out/s36_mpc_timers.err
timeout 600 python -X importtime -c "pass" 2>/dev/null; timeout 900 python - <<'PY' 2>&1 | tail -30 | tee gpurun_out/s36_mpc_pyprofile.log
import cProfile, pstats, sys, os
sys.path.insert(0, os.getcwd())
from miosqp_b200 import power_converter as pc
pc.closed_loop(2, N=10, speculation=0, replay='native').solver.work.solver.free()
pr = cProfile.Profile(); pr.enable()
r = pc.closed_loop(200, N=10, speculation=128, replay='native')
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
PY
