#!/usr/bin/env python
"""
BASELINE config 5: replay the reference's "max_iter" QP relaxations on the B200 engine, all of them in ONE launch
(every problem is its own tile).  Replaces /root/reference/extra/run_maxiter_problem.py:15-30, which loads one pickle
and solves it with OSQP; `--rho` / `--max-iter` override the stored settings the way that script does.

    python examples/replay_maxiter.py                                   # the 49 fixtures shipped with the tests
    python examples/replay_maxiter.py --pickles /path/to/max_iter_examples --only 76 --rho 0.01
"""
import argparse
import os

import numpy as np

import _common

STATUS = {1: "solved", 2: "solved inaccurate", 3: "primal infeasible inaccurate", 4: "dual infeasible inaccurate",
          -2: "maximum iterations reached", -3: "primal infeasible", -4: "dual infeasible", -7: "non convex", -10: "unsolved"}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--pickles", default=None, help="directory of the reference's *.pickle files (default: the converted test bundle)")
    ap.add_argument("--only", default=None, help="comma-separated problem numbers")
    ap.add_argument("--rho", type=float, default=None)
    ap.add_argument("--max-iter", type=int, default=None)
    args = ap.parse_args(argv)
    backend = _common.BACKEND
    from miosqp_b200 import engine, maxiter_problems

    if args.pickles:
        names = sorted(int(f.split(".")[0]) for f in os.listdir(args.pickles) if f.endswith(".pickle"))
        probs = [dict(maxiter_problems.load_pickle(os.path.join(args.pickles, "%d.pickle" % k)), name=k) for k in names]
    else:
        probs = maxiter_problems.load_npz(os.path.join(_common.ROOT, "tests", "golden", "max_iter_examples.npz"))
    if args.only:
        keep = set(int(v) for v in args.only.split(","))
        probs = [p for p in probs if p["name"] in keep]
    qps, L, U, X0, Y0 = [], [], [], [], []
    for p in probs:
        s = dict(p["settings"])
        if args.rho is not None:
            s["rho"] = args.rho
        if args.max_iter is not None:
            s["max_iter"] = args.max_iter
        qps.append(engine.BatchedQP().setup(p["P"], p["q"], p["A"], p["l"], p["u"], **s))
        n, m = p["A"].shape[1], p["A"].shape[0]
        L.append(np.asarray(p["l"], float)); U.append(np.asarray(p["u"], float)); X0.append(np.zeros(n)); Y0.append(np.zeros(m))
    xs, ys, sc = engine.solve_multi(qps, L, U, X0, Y0)
    print("backend:", backend, "|", len(probs), "problems in one launch")
    print("%5s  %-30s %6s  %12s" % ("name", "status", "iter", "objective"))
    for k, p in enumerate(probs):
        st = int(sc.status[k])
        obj = 0.5 * xs[k].dot(p["P"].dot(xs[k])) + p["q"].dot(xs[k]) if st in (1, 2, -2) else float("nan")
        print("%5d  %-30s %6d  %12.5e" % (p["name"], STATUS.get(st, str(st)), int(sc.iters[k]), obj))
    counts = {}
    for st in sc.status:
        counts[STATUS.get(int(st), str(int(st)))] = counts.get(STATUS.get(int(st), str(int(st))), 0) + 1
    print("summary:", counts)


if __name__ == "__main__":
    main()
