#!/usr/bin/env python
"""
Generates tests/golden/bnb_*.json: the UNMODIFIED reference package (/root/reference/miosqp) run in THIS
container on top of the CPU oracle through tests/osqp_shim (the reference's own `osqp` dependency is not
installable here), recording per problem the sequence of branching decisions (constr_idx, nextvar_idx), the
final status, upper_glob, x and the node / ADMM iteration counters.

    python tests/golden/make_bnb_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "osqp_shim"))
sys.path.insert(0, "/root/reference")

import osqp  # noqa: E402  (the shim)
osqp.set_backend("oracle")
import miosqp as ref  # noqa: E402
from miosqp import workspace as ref_ws  # noqa: E402
from miosqp_b200 import problems  # noqa: E402

CASES = {
    "cfg1_seed1": dict(n=50, m=100, p=5, density=0.7, seed=1),
    "cfg1_seed2": dict(n=50, m=100, p=5, density=0.7, seed=2),
    "cfg1_seed3": dict(n=50, m=100, p=5, density=0.7, seed=3),
    "small_seed5": dict(n=30, m=60, p=8, density=0.5, seed=5),
    "mid_seed7": dict(n=120, m=200, p=10, density=0.7, seed=7),
    # osqp adaptive_rho with a fixed interval (the reproducible variant of osqp's default): "qp" is merged into the QP settings
    "cfg1_seed1_adaptive": dict(n=50, m=100, p=5, density=0.7, seed=1, qp=dict(adaptive_rho=True, adaptive_rho_interval=25)),
    "mid_seed7_adaptive": dict(n=120, m=200, p=10, density=0.7, seed=7, qp=dict(adaptive_rho=True, adaptive_rho_interval=25)),
    "mid_seed9_adaptive50": dict(n=300, m=500, p=20, density=0.7, seed=9, qp=dict(adaptive_rho=True, adaptive_rho_interval=50)),
}


def run_reference(pr, settings, qp_settings, x0=None):
    decisions = []
    orig = ref_ws.Workspace.branch

    def branch(self, leaf):
        self.pick_nextvar(leaf)
        decisions.append((int(leaf.constr_idx), int(leaf.nextvar_idx)))
        self.add_left(leaf)
        self.add_right(leaf)
    ref_ws.Workspace.branch = branch
    try:
        m = ref.MIOSQP()
        m.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'], settings, qp_settings)
        if x0 is not None:
            m.set_x0(x0)
        res = m.solve()
    finally:
        ref_ws.Workspace.branch = orig
    w = m.work
    return dict(decisions=decisions, status=res.status, upper_glob=float(res.upper_glob),
                x=[float(v) for v in res.x], iter_num=int(w.iter_num), osqp_iter=int(w.osqp_iter),
                osqp_iter_avg=float(res.osqp_iter_avg), lower_glob=float(w.lower_glob))


# BASELINE config 4 (n=2000, m=4000, |i_idx|=200, 5 % dense): the full tree is far too large for a fixture (91 fractional
# integers at the root), so the golden stops at the reference's own node limit, max_iter_bb
CFG4 = dict(n=2000, m=4000, p=200, density=0.05, seed=1, max_iter_bb=10)


def main_cfg4():
    c = CFG4
    pr = problems.random_miqp(c["n"], c["m"], c["p"], c["density"], seed=c["seed"])[0]
    r = run_reference(pr, dict(problems.RANDOM_MIQP_SETTINGS, max_iter_bb=c["max_iter_bb"]), dict(problems.RANDOM_MIQP_QP_SETTINGS))
    r.pop("x")                                  # 2000 doubles: not needed, the incumbent (if any) is pinned by upper_glob
    with open(os.path.join(HERE, "bnb_cfg4_nodelimit.json"), "w") as f:
        json.dump(dict(case=c, result=r), f, indent=1)
    print("cfg4", r["status"], r["upper_glob"], "nodes", r["iter_num"] - 1, "admm", r["osqp_iter"], "branchings", len(r["decisions"]))


# BASELINE config 2 (n=500, m=1000, |i_idx|=50, density 0.7): the first draws of the 100-instance workload, to completion.
# Also records, per consumed node in consumption order, (status, ADMM iterations): the workload shape bench.py --mode bnb
# is judged on.  x is kept (500 doubles per instance).
CFG2 = dict(n=500, m=1000, p=50, density=0.7, seed=1)


def main_cfg2(count, adaptive_interval=0):
    c = dict(CFG2)
    qp_extra = {}
    if adaptive_interval:       # osqp adaptive_rho with a fixed interval (python make_bnb_golden.py --cfg2 3 --adaptive 50)
        qp_extra = dict(adaptive_rho=True, adaptive_rho_interval=adaptive_interval)
        c["qp"] = qp_extra
    prs = problems.random_miqp(c["n"], c["m"], c["p"], c["density"], seed=c["seed"], count=count)
    path = os.path.join(HERE, "bnb_cfg2_adaptive%d.json" % adaptive_interval if adaptive_interval else "bnb_cfg2.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    from miosqp import node as ref_node
    for k, pr in enumerate(prs):
        name = "cfg2_inst%d" % k
        if name in out:
            continue
        trace = []
        orig_solve = ref_node.Node.solve

        def solve(self, _o=orig_solve):
            _o(self)
            trace.append((int(self.status), int(self.num_iter), int(self.depth)))
        ref_node.Node.solve = solve
        try:
            r = run_reference(pr, dict(problems.RANDOM_MIQP_SETTINGS), dict(problems.RANDOM_MIQP_QP_SETTINGS, **qp_extra))
        finally:
            ref_node.Node.solve = orig_solve
        r["node_trace"] = trace
        out[name] = dict(case=dict(c, instance=k), result=r)
        print(name, r["status"], r["upper_glob"], "nodes", r["iter_num"] - 1, "admm", r["osqp_iter"], "branchings", len(r["decisions"]), flush=True)
        with open(path, "w") as f:
            json.dump(out, f, indent=1)


def main():
    if "--cfg4" in sys.argv:
        return main_cfg4()
    if "--cfg2" in sys.argv:
        return main_cfg2(int(sys.argv[sys.argv.index("--cfg2") + 1]),
                         int(sys.argv[sys.argv.index("--adaptive") + 1]) if "--adaptive" in sys.argv else 0)
    out, out_adaptive = {}, {}
    for name, c in CASES.items():
        pr = problems.random_miqp(c["n"], c["m"], c["p"], c["density"], seed=c["seed"])[0]
        dst = out_adaptive if "qp" in c else out       # the adaptive-rho cases live in a file of their own
        dst[name] = dict(case=c, result=run_reference(pr, dict(problems.RANDOM_MIQP_SETTINGS),
                                                       dict(problems.RANDOM_MIQP_QP_SETTINGS, **c.get("qp", {}))))
        r = dst[name]["result"]
        print(name, r["status"], r["upper_glob"], "nodes", r["iter_num"] - 1, "admm", r["osqp_iter"], "branchings", len(r["decisions"]))
    with open(os.path.join(HERE, "bnb_random_miqp.json"), "w") as f:
        json.dump(out, f, indent=1)
    with open(os.path.join(HERE, "bnb_random_miqp_adaptive.json"), "w") as f:
        json.dump(out_adaptive, f, indent=1)


if __name__ == "__main__":
    main()
