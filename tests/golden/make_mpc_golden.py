#!/usr/bin/env python
"""
Generates tests/golden/mpc_power_converter.npz: BASELINE config 3.  The reference's power-converter example
(/root/reference/examples/power_converter, horizon N=10, parameters of run_example.py:25-77) is imported IN THIS
CONTAINER (matplotlib / mathprogbasepy / ipdb are absent and only used for plotting / Gurobi, so they are
stubbed), and its closed loop (power_converter.py:589-649) is driven for the first STEPS time steps with the
UNMODIFIED reference miosqp package on the CPU oracle (tests/osqp_shim).  Per step the fixture stores the QP
vectors handed to MIOSQP (q, l, u, x0 = shifted previous input) and the reference's answers (x, upper_glob,
status, node count, total ADMM iterations, branching decisions).

    python tests/golden/make_mpc_golden.py
"""
import hashlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
STEPS = 120          # the first 8 with their full branching sequences (round 1), the rest with a sha256 of it
FULL_DECISIONS = 8
N_HORIZON = 10

class _Stub(types.ModuleType):
    """Absorbs any attribute access / call made at import time by the plotting helpers."""

    def __getattr__(self, item):
        return lambda *a, **k: None


for name in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot", "mathprogbasepy", "ipdb"):
    sys.modules.setdefault(name, _Stub(name))
sys.modules["matplotlib"].pylab = sys.modules["matplotlib.pylab"]
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "osqp_shim"))
sys.path.insert(0, "/root/reference")
os.chdir("/root/reference")

import osqp  # noqa: E402
osqp.set_backend("oracle")
import miosqp  # noqa: E402
from miosqp import workspace as ref_ws  # noqa: E402
from examples.power_converter.power_converter import Model  # noqa: E402
from examples.power_converter.quadratic_program import MIQP  # noqa: E402


def main():
    model = Model()
    model.set_params(25.0e-06, 50., 0.8e03, 0.8e03, 1.)
    model.set_time(0.0, 1, 2)
    model.set_initial_conditions()
    model.gen_dynamical_system(300, 5.5)
    model.gen_tail_cost(50, 0.95, name='delta_550.mat')
    model.solver = None
    model.qp_matrices = MIQP(model.dyn_system, N_HORIZON, model.tail_cost)
    qp = model.qp_matrices
    nu = model.dyn_system.B.shape[1]

    decisions = []
    orig = ref_ws.Workspace.branch

    def branch(self, leaf):
        self.pick_nextvar(leaf)
        decisions.append((int(leaf.constr_idx), int(leaf.nextvar_idx)))
        self.add_left(leaf)
        self.add_right(leaf)
    ref_ws.Workspace.branch = branch

    x = np.array(model.init_conditions.x0, dtype=float)
    u_prev = np.zeros(nu * N_HORIZON)
    out = dict(P=qp.P.toarray(), A=qp.A.toarray(), i_idx=np.asarray(qp.i_idx), i_l=np.asarray(qp.i_l, float),
               i_u=np.asarray(qp.i_u, float), steps=np.array(STEPS))
    for k in range(STEPS):
        del decisions[:]
        q = 2. * (qp.q_x.dot(x) + qp.q_u)
        u0, obj, _, u_full, _, _ = model.compute_mpc_input(x, u_prev, solver='miosqp')
        w = model.solver.work
        out["q_%d" % k] = q
        out["l_%d" % k] = np.array(qp.l, dtype=float)
        out["u_%d" % k] = np.array(qp.u, dtype=float)
        out["x0_%d" % k] = np.array(u_prev, dtype=float)
        out["sol_%d" % k] = np.array(u_full, dtype=float)
        out["obj_%d" % k] = np.array(float(obj))
        out["stats_%d" % k] = np.array([w.iter_num, w.osqp_iter])
        dec = np.array(decisions, dtype=np.int64).reshape(-1, 2)
        if k < FULL_DECISIONS:
            out["dec_%d" % k] = dec
        out["dech_%d" % k] = np.frombuffer(hashlib.sha256(dec.tobytes()).digest(), dtype=np.uint8)
        out["status_%d" % k] = np.array(str(w.status))
        print("step", k, "status", w.status, "obj %.6f" % obj, "nodes", w.iter_num - 1, "admm", w.osqp_iter, "branchings", len(decisions))
        x = np.asarray(model.simulate_one_step(x, u0)[0], dtype=float).flatten()
        u_prev = np.append(u_full[nu:], u_full[-nu:])
    ref_ws.Workspace.branch = orig
    path = os.path.join(HERE, "mpc_power_converter.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; n =", out["P"].shape[0], "m =", out["A"].shape[0], "n_int =", len(out["i_idx"]))


if __name__ == "__main__":
    main()
