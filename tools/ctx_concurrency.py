#!/usr/bin/env python
"""
Do tiles launched from different solve contexts (CUDA streams, host threads) overlap on the GPU?  K threads, each with its
own bqp_ctx, each solve one tile of 8 leaves with a fixed iteration count; wall time as a function of K.
"""
import os
import sys
import threading
import time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from miosqp_b200 import engine, problems

ITERS = 400
base = problems.extend(problems.random_miqp(500, 1000, 50, 0.7, seed=1, count=1)[0])
P, q, A, l, u, i_idx = base
st = dict(eps_abs=1e-12, eps_rel=1e-12, eps_prim_inf=1e-12, eps_dual_inf=1e-12, max_iter=ITERS, check_termination=ITERS)
KMAX = int(sys.argv[1]) if len(sys.argv) > 1 else 32
es = engine.setup_many([(P, q * (1 + 0.01 * k), A, l, u, i_idx) for k in range(KMAX)], **st)
rng = np.random.default_rng(0)
ls, us = problems.branched_nodes(l, u, len(i_idx), 8, rng)
ctxs = [engine.Context(0, True) for _ in range(KMAX)]
n, m = A.shape[1], A.shape[0]


def work(k, reps):
    for _ in range(reps):
        ctxs[k].solve_multi([es[k]] * 8, list(ls), list(us), [np.zeros(n)] * 8, [np.zeros(m)] * 8)


work(0, 1)
for K in [1, 2, 4, 8, 16, 32, 64]:
    if K > KMAX:
        break
    th = [threading.Thread(target=work, args=(k, 2)) for k in range(K)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    print("contexts=%2d: %.1f ms per solve round (one tile of 8 leaves x %d iterations each; kernel_ms of ctx 0: %.1f)" % (
        K, 1e3 * dt / 2, ITERS, ctxs[0].last_timing()["kernel_ms"]), flush=True)
