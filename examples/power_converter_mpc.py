#!/usr/bin/env python
"""
BASELINE config 3: closed-loop MPC of the three-level inverter (horizon N, warm-started re-solves every 25 us
sampling instant) on the B200 engine -- the reference's examples/power_converter/run_example.py without its Gurobi
arm, THD figures and plots (/root/reference/examples/power_converter/run_example.py:21-150,
power_converter.py:589-675).  Prints the reference's timing table (avg/std/min/max solve time per MPC step, OSQP share,
average ADMM iterations) per horizon, plus B&B nodes, engine launches and the switching frequency, and one JSON line.

    python examples/power_converter_mpc.py --horizons 10 --steps 1000 --speculation 32       # BASELINE config 3
    python examples/power_converter_mpc.py --horizons 1,2,3,4,5                              # the reference's sweep
"""
import argparse
import json
import time

import numpy as np
import pandas as pd

import _common


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--horizons", default="1,2,3,4,5")
    ap.add_argument("--steps", type=int, default=None, help="sampling instants (default: the reference's 3 periods = 2400)")
    ap.add_argument("--speculation", type=int, default=32, help="nodes solved ahead of the replay per launch (0 = off)")
    ap.add_argument("--replay", default="python", choices=("python", "native"),
                    help="B&B loop in miosqp_b200/tree.py (python) or in csrc/bqp_bnb.cpp (native): same decisions")
    ap.add_argument("--tail", default="delta_550")
    ap.add_argument("--csv", default=None)
    args = ap.parse_args(argv)
    backend = _common.BACKEND
    from miosqp_b200 import power_converter as pc

    drive = pc.Drive()
    steps = args.steps if args.steps is not None else 3 * drive.steps_per_period
    rows = []
    for N in [int(v) for v in args.horizons.split(",")]:
        t0 = time.perf_counter()
        r = pc.closed_loop(steps, N=N, drive=drive, tail=args.tail, speculation=args.speculation,
                           replay='native' if args.replay == 'native' else None)
        wall = time.perf_counter() - t0
        w = r.solver.work
        t = r.run_time
        rows.append(dict(T=N, steps=steps, miosqp_avg=t.mean(), miosqp_std=t.std(), miosqp_min=t.min(), miosqp_max=t.max(),
                         miosqp_osqp_avg_time=float(np.mean(100 * r.osqp_solve_time / r.run_time)),
                         miosqp_avg_osqp_iter=float(r.osqp_iter_avg.mean()), nodes_per_step=float(r.nodes.mean()),
                         admm_iters_per_step=float(r.admm_iters.mean()), launches_per_step=w.batches / float(steps),
                         solved_nodes_per_step=w.batched_nodes / float(steps), spec_hit_rate=w.spec_hits / float(max(1, w.spec_nodes)),
                         node_limit_steps=sum(s != 'Solved' for s in r.status), fsw_hz=r.switching_frequency,
                         qp_per_s=float(r.nodes.sum()) / wall, wall_s=wall))
        r.solver.work.solver.free()
    table = pd.DataFrame(rows)
    print("backend:", backend, "| speculation", args.speculation, "| replay", args.replay)
    print(table.to_string(index=False, float_format=lambda v: "%.4g" % v))
    if args.csv:
        table.to_csv(args.csv, index=False)
    print(json.dumps({"workload": "power_converter MPC closed loop", "backend": backend, "speculation": args.speculation, "replay": args.replay, "rows": rows}))


if __name__ == "__main__":
    main()
