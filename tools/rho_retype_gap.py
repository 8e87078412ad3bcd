#!/usr/bin/env python
"""
What does the engine's parity contract (rho typed per row ONCE at setup) cost against osqp >= 0.4's
behaviour (rows re-typed at every update_bounds, equality rows get rho x 1e3 and the KKT matrix is refactored;
SURVEY App. A.3, section 8f2)?  Runs the B&B replay on the CPU oracle under both rules (eq_rho = 1 / 2) and prints
nodes, ADMM iterations and the optimum per workload.  CPU only; measures the ALGORITHMIC gap, not a speed.

    python tools/rho_retype_gap.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fake_engine                                   # noqa: E402  (oracle-backed stand-in: this tool is a checker, not a product path)
import miosqp_b200                                   # noqa: E402
from miosqp_b200 import engine, problems, power_converter as pc   # noqa: E402

engine.BatchedQP = fake_engine.FakeBatchedQP
engine.solve_multi = fake_engine.solve_multi


def bnb(pr, eq_rho):
    s = miosqp_b200.MIOSQP()
    s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
            dict(problems.RANDOM_MIQP_SETTINGS), dict(problems.RANDOM_MIQP_QP_SETTINGS, eq_rho=eq_rho))
    r = s.solve()
    return dict(status=r.status, nodes=s.work.iter_num - 1, admm_iters=int(s.work.osqp_iter), upper_glob=float(r.upper_glob))


def main():
    rows = []
    for (n, m, p, seed) in [(50, 100, 5, 1), (30, 60, 10, 1), (40, 40, 20, 3), (100, 200, 15, 2)]:
        pr = problems.random_miqp(n, m, p, 0.7, seed=seed)[0]
        a, b = bnb(pr, 1), bnb(pr, 2)
        rows.append(dict(workload="random_miqp n=%d m=%d p=%d seed=%d" % (n, m, p, seed), setup_typing=a, per_node_typing=b))
    for N in (3, 10):
        out = {}
        for eq in (1, 2):
            r = pc.closed_loop(10, N=N, speculation=32, qp_settings=dict(eq_rho=eq))
            out[eq] = dict(nodes=int(r.nodes.sum()), admm_iters=int(r.admm_iters.sum()), node_limit_steps=sum(s != 'Solved' for s in r.status),
                           obj_sum=float(r.obj.sum()), inputs_equal_to_setup_typing=None)
            out[eq]["U"] = r.U
        out[2]["inputs_equal_to_setup_typing"] = bool(np.array_equal(out[1].pop("U"), out[2].pop("U")))
        rows.append(dict(workload="power_converter MPC N=%d, first 10 steps" % N, setup_typing=out[1], per_node_typing=out[2]))
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
