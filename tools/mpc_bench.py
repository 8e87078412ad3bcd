#!/usr/bin/env python
"""
Config 3 (power-converter MPC closed loop, horizon N) on the CUDA engine with different look-ahead budgets
(`settings['speculation']`, miosqp_b200/tree.py): wall time per MPC step, launches per step, consumed and solved QP
relaxations per second.  One JSON line per budget, flushed as it goes.

    python tools/mpc_bench.py --steps 10 --horizon 10 --budgets 0,32,256
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--horizon", type=int, default=10)
    ap.add_argument("--budgets", default="0,32,256")
    ap.add_argument("--replay", default="python", choices=("python", "native"))
    ap.add_argument("--warm", action="store_true", help="run 2 untimed steps first (CUDA context, first-launch costs)")
    args = ap.parse_args()
    from miosqp_b200 import power_converter as pc
    ref = None
    replay = 'native' if args.replay == 'native' else None
    if args.warm:
        pc.closed_loop(2, N=args.horizon, speculation=0, replay=replay).solver.work.solver.free()
    for b in [int(v) for v in args.budgets.split(",")]:
        t0 = time.perf_counter()
        r = pc.closed_loop(args.steps, N=args.horizon, speculation=b, replay=replay)
        wall = time.perf_counter() - t0
        w = r.solver.work
        same = None if ref is None else bool((ref.U == r.U).all() and (ref.nodes == r.nodes).all() and (ref.admm_iters == r.admm_iters).all())
        if ref is None:
            ref = r
        print(json.dumps({"workload": "power_converter MPC N=%d, %d steps" % (args.horizon, args.steps), "speculation": b, "replay": args.replay,
                          "ms_per_mpc_step": 1e3 * wall / args.steps, "launches_per_step": w.batches / float(args.steps),
                          "nodes_per_step": float(r.nodes.mean()), "solved_nodes_per_step": w.batched_nodes / float(args.steps),
                          "qp_per_s_consumed": float(r.nodes.sum()) / wall, "qp_per_s_solved": w.batched_nodes / wall,
                          "spec_hit_rate": w.spec_hits / float(max(1, w.spec_nodes)), "same_inputs_and_counts_as_first": same,
                          "node_limit_steps": sum(s != 'Solved' for s in r.status)}), flush=True)
        w.solver.free()


if __name__ == "__main__":
    main()
