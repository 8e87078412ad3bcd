#!/usr/bin/env python
"""
Does a node's result depend on the launch it is solved in?  Records every node of one B&B run (random MIQP
n=40 m=40 p=20, the problem of tests/test_examples.py), then re-solves the same node inputs (a) one by one,
(b) all in one launch, (c) all in one launch again, for each kernel, and compares status / iteration count /
iterates with the first answers and with the CPU oracle (checker only).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miosqp_b200                                  # noqa: E402
from miosqp_b200 import engine, problems            # noqa: E402
from oracle import oracle                           # noqa: E402


SWEEP = [(34, 30, 8, 0.7, 1), (40, 40, 20, 0.7, 3), (64, 40, 10, 0.7, 2), (50, 100, 5, 0.7, 1), (96, 60, 12, 0.7, 4),
         (70, 150, 10, 0.1, 5), (130, 60, 10, 0.7, 6), (33, 96, 6, 0.5, 7)]      # (n, m, p, density, seed): A passes of 5..20 row panels


def main():
    if "--sweep" in sys.argv:
        for n, m, p, d, seed in SWEEP:
            print("==== n=%d m=%d p=%d density=%g seed=%d" % (n, m, p, d, seed), flush=True)
            one(n, m, p, d, seed)
        return
    n, m, p, seed = (int(v) for v in (sys.argv[1:5] + [40, 40, 20, 3][len(sys.argv) - 1:]))
    one(n, m, p, 0.7, seed)


def one(n, m, p, density, seed):
    pr = problems.random_miqp(n, m, p, density, seed=seed)[0]
    rec = []
    real = engine.solve_multi

    def spy(qps, l, u, x0, y0):
        xs, ys, sc = real(qps, l, u, x0, y0)
        for k in range(len(qps)):
            rec.append((np.array(l[k]), np.array(u[k]), np.array(x0[k]), np.array(y0[k]), int(sc.status[k]), int(sc.iters[k]), np.array(xs[k])))
        return xs, ys, sc
    engine.solve_multi = spy
    s = miosqp_b200.MIOSQP()
    s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
            dict(problems.RANDOM_MIQP_SETTINGS), dict(problems.RANDOM_MIQP_QP_SETTINGS))
    s.solve()
    engine.solve_multi = real
    qp = s.work.solver
    print("recorded", len(rec), "nodes; timing of last launch:", engine.last_timing())
    P, q, A, l, u, i_idx = problems.extend(pr)
    o = oracle.OSQP(); o.setup(P, q, A, l, u, **problems.RANDOM_MIQP_QP_SETTINGS)
    L = np.array([r[0] for r in rec]); U = np.array([r[1] for r in rec]); X0 = np.array([r[2] for r in rec]); Y0 = np.array([r[3] for r in rec])
    xo, yo, so, io, extra = o.solve_batch(L, U, X0, Y0, threads=8)
    it_rec = np.array([r[5] for r in rec]); st_rec = np.array([r[4] for r in rec])
    print("as recorded (pairs) vs oracle: iters differ at", np.where(it_rec != io)[0].tolist(), "status differ at", np.where(st_rec != so)[0].tolist())
    for kern in ("panel", "stream", "direct"):
        os.environ["BQP_KERNEL"] = kern
        try:
            one = [qp.solve_batch(L[k:k + 1], U[k:k + 1], X0[k:k + 1], Y0[k:k + 1]) for k in range(len(rec))]
            it1 = np.array([int(r.iters[0]) for r in one])
            allb = qp.solve_batch(L, U, X0, Y0); t = engine.last_timing()
            allb2 = qp.solve_batch(L, U, X0, Y0)
            print("%-6s kernel=%d tiles=%d tile_nodes=%d launches=%d | one-by-one vs oracle: %s | all-in-one vs oracle: %s | all-in-one twice identical: %s | bitwise x one-by-one == all-in-one: %s"
                  % (kern, t["kernel"], t["tiles"], t["tile_nodes"], t["launches"], np.where(it1 != io)[0].tolist(),
                     np.where(allb.iters != io)[0].tolist(), bool(np.array_equal(allb.iters, allb2.iters) and np.array_equal(allb.x, allb2.x, equal_nan=True)),
                     bool(all(np.array_equal(one[k].x[0], allb.x[k], equal_nan=True) for k in range(len(rec))))))
            for k in np.where((it1 != io) | (allb.iters != io) | (it_rec != io))[0]:
                print("   node %d: oracle it=%d st=%d pri=%.3e dua=%.3e | recorded it=%d | alone it=%d pri=%.3e dua=%.3e | batch it=%d pri=%.3e dua=%.3e"
                      % (k, io[k], so[k], extra["pri_res"][k], extra["dua_res"][k], it_rec[k], it1[k], one[k].pri_res[0], one[k].dua_res[0],
                         allb.iters[k], allb.pri_res[k], allb.dua_res[k]))
        except Exception as e:                      # a kernel that cannot run this size
            print(kern, "->", type(e).__name__, e)


if __name__ == "__main__":
    main()
