#!/usr/bin/env python
"""
Random MIQP timing harness on the B200 engine: the reference's examples/random_miqp/run_example.py without its
Gurobi arm and LaTeX output (/root/reference/examples/random_miqp/run_example.py:28-224).  For every (n, m, p) it
draws `--repeat` instances with the reference generator (miosqp_b200/problems.py), solves each with MIOSQP at the
reference's settings and prints the reference's timing table (t_min/t_max/t_avg/t_std in ms, OSQP share in %),
plus nodes and ADMM iterations.  `--together` solves the instances of one size in lock-step (`solve_many`: one launch
per B&B step covers every instance's frontier), which is how BASELINE config 2 is meant to be run.

    python examples/random_miqp.py                       # problem set 1 of the reference
    python examples/random_miqp.py --sizes 500,1000,50 --repeat 100 --together     # BASELINE config 2
"""
import argparse
import time

import numpy as np
import pandas as pd

import _common

SETS = {1: ([10, 10, 50, 50, 100, 100, 150, 150], [5, 100, 25, 200, 50, 200, 100, 300], [2, 2, 5, 10, 2, 15, 5, 20]),
        2: ([2, 4, 8, 12, 20, 26, 30, 36], None, None)}         # run_example.py:181-190


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--problem-set", type=int, default=1, choices=(1, 2))
    ap.add_argument("--sizes", default=None, help="n,m,p[;n,m,p...] instead of a problem set")
    ap.add_argument("--repeat", type=int, default=10)
    ap.add_argument("--density", type=float, default=0.7)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--together", action="store_true", help="solve the instances of one size in lock-step (solve_many)")
    ap.add_argument("--speculation", type=int, default=0, help="nodes solved ahead of the replay per launch")
    ap.add_argument("--replay", default="python", choices=("python", "native"), help="B&B loop in Python (tree.py) or in C++ (bqp_bnb_solve / bqp_bnb_solve_many)")
    ap.add_argument("--csv", default=None)
    args = ap.parse_args(argv)
    backend = _common.BACKEND
    import miosqp_b200
    from miosqp_b200 import problems

    if args.sizes:
        trip = [tuple(int(v) for v in s.split(",")) for s in args.sizes.split(";")]
        n_arr, m_arr, p_arr = zip(*trip)
    else:
        n_arr, m_arr, p_arr = SETS[args.problem_set]
        if m_arr is None:
            m_arr = [5 * n for n in n_arr]; p_arr = [n // 2 for n in n_arr]
    np.random.seed(args.seed)                                   # one seed for the whole sweep (run_example.py:43)
    rows = []
    for n, m, p in zip(n_arr, m_arr, p_arr):
        solvers = []
        t_setup = time.perf_counter()
        for _ in range(args.repeat):
            pr = problems.random_miqp_draw(n, m, p, args.density)
            s = miosqp_b200.MIOSQP()
            s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
                    dict(problems.RANDOM_MIQP_SETTINGS, speculation=args.speculation,
                         replay='native' if args.replay == 'native' else None),
                    dict(problems.RANDOM_MIQP_QP_SETTINGS))
            solvers.append(s)
        t_setup = time.perf_counter() - t_setup
        t0 = time.perf_counter()
        results = miosqp_b200.solve_many(solvers) if args.together else [s.solve() for s in solvers]
        wall = time.perf_counter() - t0
        bad = [r.status for r in results if r.status != miosqp_b200.MI_SOLVED]
        ms = 1e3 * np.array([r.run_time for r in results])
        share = 100 * np.array([r.osqp_solve_time / r.run_time for r in results])
        nodes = np.array([s.work.iter_num - 1 for s in solvers]); admm = np.array([s.work.osqp_iter for s in solvers])
        rows.append(dict(n=n, m=m, p=p, t_min=ms.min(), t_max=ms.max(), t_avg=ms.mean(), t_std=ms.std(),
                         t_osqp_avg=share.mean(), nodes_avg=nodes.mean(), admm_iters_avg=admm.mean(),
                         launches=sum(s.work.batches for s in solvers) if not args.together else max(s.work.batches for s in solvers),
                         wall_ms=1e3 * wall, setup_ms=1e3 * t_setup, not_solved=len(bad)))
        for s in solvers:
            s.work.solver.free()
    table = pd.DataFrame(rows)
    print("backend:", backend, "| repeat", args.repeat, "| lock-step" if args.together else "| one instance at a time")
    print(table.to_string(index=False, float_format=lambda v: "%.2f" % v))
    if args.csv:
        table.to_csv(args.csv, index=False)


if __name__ == "__main__":
    main()
