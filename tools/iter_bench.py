#!/usr/bin/env python
"""
Per-iteration timing of the ADMM kernels with convergence switched off (fixed iteration count, so no
tile-to-tile imbalance): separates kernel efficiency from the workload's iteration-count spread.

    python tools/iter_bench.py [--instances I] [--leaves L] [--iters K] [--tile-nodes T] [--threads N]
"""
import argparse
import os
import sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from miosqp_b200 import engine, problems

ap = argparse.ArgumentParser()
ap.add_argument("--instances", type=int, default=100)
ap.add_argument("--leaves", type=int, default=8)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--tile-nodes", type=int, default=0)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--n", type=int, default=500)
ap.add_argument("--m", type=int, default=1000)
ap.add_argument("--p", type=int, default=50)
ap.add_argument("--density", type=float, default=0.7)
ap.add_argument("--mpc", action="store_true", help="BASELINE config 3: the power-converter MPC program (n = 60, m = 150) instead of a random MIQP")
a = ap.parse_args()
if a.mpc:
    import scipy.sparse as spa
    from miosqp_b200 import power_converter as pc
    drive = pc.Drive(); system = pc.System(drive, 300, 5.5)
    prog = pc.MpcProgram(system, 10, pc.TailCost(system, 0.95, "delta_550"))
    q, l, u = prog.vectors(drive.initial_state())
    A = spa.vstack([prog.A, spa.identity(prog.P.shape[0], format="csc")[prog.i_idx, :]]).tocsc()
    base = (prog.P, q, A, np.append(l, prog.i_l), np.append(u, prog.i_u), np.asarray(prog.i_idx))
    a.n = 60
else:
    base = problems.extend(problems.random_miqp(a.n, a.m, a.p, a.density, seed=1, count=1)[0])
P, q, A, l, u, i_idx = base
rng = np.random.default_rng(0)
st = dict(eps_abs=1e-12, eps_rel=1e-12, eps_prim_inf=1e-12, eps_dual_inf=1e-12, max_iter=a.iters, check_termination=a.iters)
# distinct device copies of the same matrices (distinct HBM addresses), host halves of the setup on all host threads
es = engine.setup_many([(P, q * (1 + 0.01 * k), A, l, u, i_idx) for k in range(a.instances)], **st)
engine.set_tuning(a.tile_nodes, a.threads)
for count in sorted(set([1, a.instances])):
    qps, L, U, X0, Y0 = [], [], [], [], []
    for e in es[:count]:
        ls, us = problems.branched_nodes(l, u, len(i_idx), a.leaves, rng)
        for b in range(a.leaves):
            qps.append(e); L.append(ls[b]); U.append(us[b]); X0.append(np.zeros(a.n)); Y0.append(np.zeros(A.shape[0]))
    rb = engine.ResidentBatch(qps, L, U, X0, Y0)
    ms = []
    for _ in range(4):
        rb.run()
        ms.append(engine.last_timing()["kernel_ms"])
    rb.download()
    t = engine.last_timing()
    best = min(ms[1:])
    per_iter_us = 1e3 * best / a.iters
    print("%s instances=%d tiles=%d T=%d threads=%d smem=%d slots=%d kernel %.2f ms  -> %.1f us/iteration/tile-wave, streamed %.0f GB/s, node-iters/s %.3e" % (
        os.environ.get("BQP_LIB_SUFFIX", ""), count, t["tiles"], t["tile_nodes"], t["threads"], t["smem_bytes"], t["ring_slots"], best, per_iter_us,
        t["stream_bytes"] / best / 1e6, t["node_iters"] / best * 1e3), flush=True)
