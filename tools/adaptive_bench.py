#!/usr/bin/env python
"""
Adaptive rho (osqp adaptive_rho, fixed interval) against the fixed-rho contract on single B&B trees:
config 2 instances (random_miqp n=500 m=1000 |i_idx|=50) and, with --cfg4, config 4 (n=2000, cut at 40 nodes).

    python tools/adaptive_bench.py [--instances 3] [--interval 50] [--cfg4]
"""
import argparse
import json
import os
import sys
import time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from miosqp_b200 import engine, problems, miqp

ap = argparse.ArgumentParser()
ap.add_argument("--instances", type=int, default=3)
ap.add_argument("--interval", type=int, default=50)
ap.add_argument("--cfg4", action="store_true")
a = ap.parse_args()
if a.cfg4:
    prs = problems.random_miqp(2000, 4000, 200, 0.05, seed=1, count=1)
    extra = dict(max_iter_bb=40)
else:
    prs = problems.random_miqp(500, 1000, 50, 0.7, seed=1, count=a.instances)
    extra = {}
for name, qp_extra, st_extra in (("fixed rho (rows kernel, automatic cluster size)" if not a.cfg4 else "fixed rho (whole-GPU kernel)", {}, dict(cluster_auto=True)),
                                 ("adaptive rho, interval %d (whole-GPU kernel, spectral inverse)" % a.interval,
                                  dict(adaptive_rho=True, adaptive_rho_interval=a.interval), {})):
    t0 = time.perf_counter()
    ss = miqp.setup_many(prs, dict(problems.RANDOM_MIQP_SETTINGS, replay='native', **st_extra, **extra),
                         dict(problems.RANDOM_MIQP_QP_SETTINGS, **qp_extra))
    t_setup = time.perf_counter() - t0
    for k, s in enumerate(ss):
        for rep in range(2):                      # second run: warm context
            s.work.reset(); s.work.first_run = 0
            t0 = time.perf_counter(); r = s.solve(); dt = time.perf_counter() - t0
        print(json.dumps({"mode": name, "instance": k, "wall_s": dt, "nodes": s.work.iter_num - 1, "admm_iters": int(s.work.osqp_iter),
                          "status": r.status, "upper_glob": r.upper_glob, "kernel": engine.last_timing()["kernel"], "setup_s": t_setup}), flush=True)
