#!/usr/bin/env python
"""
BASELINE config 2 as the reference runs it: branch-and-bound TO COMPLETION over the random_miqp instances
(n=500, m=1000, |i_idx|=50; /root/reference/examples/random_miqp/run_example.py:28-154, settings :98-116), every
relaxation on the CUDA engine.  Compares the drivers: lock-step (one launch per B&B step over all trees) in Python and in
C++, and the asynchronous native driver (every tree on its own stream), each at several look-ahead budgets.
One JSON line per run; decisions are checked against tests/golden/bnb_cfg2.json and across runs.

    python tools/bnb_bench.py --instances 100 --runs lockstep:0,async:0,async:6
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=100)
    ap.add_argument("--runs", default="lockstep:0,async:0,async:6")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    import miosqp_b200
    from miosqp_b200 import problems, miqp, engine
    prs = problems.random_miqp(500, 1000, 50, 0.7, seed=a.seed, count=a.instances)
    golden = {}
    gp = os.path.join(ROOT, "tests", "golden", "bnb_cfg2.json")
    if a.seed == 1 and os.path.exists(gp):
        golden = json.load(open(gp))
    t0 = time.perf_counter()
    solvers = miqp.setup_many(prs, dict(problems.RANDOM_MIQP_SETTINGS, replay='native'), dict(problems.RANDOM_MIQP_QP_SETTINGS))
    setup_s = time.perf_counter() - t0
    first = None
    for spec in a.runs.split(","):
        mode, k = spec.split(":")
        k = int(k)
        for s in solvers:
            s.work.settings['speculation'] = k
            s.work.settings['replay'] = None if mode == "python" else 'native'
            s.work.reset(); s.work.first_run = 0
            s.work.batches = s.work.batched_nodes = s.work.spec_nodes = s.work.spec_hits = 0
        t0 = time.perf_counter()
        roll = False
        if mode.startswith("rolling"):              # rolling = automatic number of sessions, rolling1 / rolling2 / ... = that many
            roll = int(mode[7:]) if mode[7:] else True
        res = miosqp_b200.solve_many(solvers, async_threads=(a.threads if mode == "async" else None), rolling=roll)
        wall = time.perf_counter() - t0
        works = [s.work for s in solvers]
        consumed = sum(w.iter_num - 1 for w in works)
        decisions = [list(map(tuple, w.decisions)) for w in works]
        sig = [(r.status, round(float(r.upper_glob), 9), w.iter_num, int(w.osqp_iter)) for r, w in zip(res, works)]
        same = None if first is None else bool(sig == first[0] and decisions == first[1])
        if first is None:
            first = (sig, decisions)
        gold_ok = None
        if golden:
            gold_ok = True
            for i in range(min(a.instances, len(golden))):
                g = golden["cfg2_inst%d" % i]["result"]
                ok = (decisions[i] == [tuple(d) for d in g["decisions"]] and res[i].status == g["status"] and works[i].iter_num == g["iter_num"]
                      and int(works[i].osqp_iter) == g["osqp_iter"] and abs(res[i].upper_glob - g["upper_glob"]) <= 1e-9 * (1 + abs(g["upper_glob"])))
                gold_ok = gold_ok and ok
        print(json.dumps({"workload": "random_miqp n=500 m=1000 |i_idx|=50, %d instances, B&B to completion" % a.instances,
                          "driver": mode, "speculation": k, "wall_s": wall, "setup_s": setup_s,
                          "qp_consumed": consumed, "qp_solved": int(sum(w.batched_nodes for w in works)),
                          "qp_per_s_consumed": consumed / wall, "admm_iters_consumed": int(sum(w.osqp_iter for w in works)),
                          "admm_node_iters_per_s_consumed": sum(w.osqp_iter for w in works) / wall,
                          "launches": int(sum(w.batches for w in works)) if mode == "async" else int(max(w.batches for w in works)),
                          "spec_hit_rate": sum(w.spec_hits for w in works) / float(max(1, sum(w.spec_nodes for w in works))),
                          "status": {st: [r.status for r in res].count(st) for st in set(r.status for r in res)},
                          "engine_timing_default_ctx": {k: engine.last_timing()[k] for k in ("launches", "kernel_ms", "tile_iters", "tiles", "tile_nodes")},
                          "same_as_first_run": same, "golden_instances_ok": gold_ok,
                          "nodes_per_instance_max": max(w.iter_num - 1 for w in works)}), flush=True)


if __name__ == "__main__":
    main()
