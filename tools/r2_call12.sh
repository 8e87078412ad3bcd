#!/usr/bin/env bash
set -u
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rows_kernel_cluster or inverse_guard" 2>&1 | tail -3
BQP_ROWS_CLUSTER=8 timeout 120 python tools/iter_bench.py --instances 16 --iters 200 2>&1 | tail -2 | cut -c1-200
python - <<'PY'
import time, json
import miosqp_b200
from miosqp_b200 import problems, miqp
prs = problems.random_miqp(500, 1000, 50, 0.7, seed=1, count=3)
for auto in (False, True):
    ss = miqp.setup_many(prs, dict(problems.RANDOM_MIQP_SETTINGS, replay='native', cluster_auto=auto), dict(problems.RANDOM_MIQP_QP_SETTINGS))
    for k, s in enumerate(ss):
        t0 = time.perf_counter(); r = s.solve(); dt = time.perf_counter() - t0
        print("single tree inst %d cluster_auto=%s: %.2f s, %d nodes, %d ADMM iterations, status %s, upper %.9f, %.1f us per iteration on the critical path" % (
            k, auto, dt, s.work.iter_num - 1, s.work.osqp_iter, r.status, r.upper_glob, 1e6 * dt / max(1, s.work.osqp_iter) * 2), flush=True)
PY
timeout 600 python tools/bnb_bench.py --instances 100 --runs rolling:0 2>&1 | cut -c1-700
