#!/usr/bin/env python
"""
Generates the fixtures the native power-converter workload (miosqp_b200/power_converter.py) is checked against:

  miosqp_b200/data/tail_delta_{4,510,550}.npz    the ADP tail costs P0 (12x12), q0 (12), r0 of
                                                 /root/reference/examples/power_converter/tail_backups/*.mat
                                                 (DATA of the reference example, tail_cost.py:15-19)
  tests/golden/power_converter_model.npz         what the reference's own model builder produces for the parameters of
                                                 run_example.py:25-77: system matrices, initial state, and the MIQP
                                                 matrices of quadratic_program.py for horizons N = 1, 3, 10

The reference example is imported IN THIS CONTAINER only (plot / Gurobi modules stubbed); the GPU box never reads
/root/reference.

    python tests/golden/make_power_converter_model_golden.py
"""
import os
import sys
import types

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


class _Stub(types.ModuleType):
    def __getattr__(self, item):
        return lambda *a, **k: None


for name in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot", "mathprogbasepy", "ipdb"):
    sys.modules.setdefault(name, _Stub(name))
sys.modules["matplotlib"].pylab = sys.modules["matplotlib.pylab"]
sys.path.insert(0, os.path.join(ROOT, "tests", "osqp_shim"))
sys.path.insert(0, REF)
os.chdir(REF)

from examples.power_converter.power_converter import Model  # noqa: E402
from examples.power_converter.quadratic_program import MIQP  # noqa: E402


def main():
    data_dir = os.path.join(ROOT, "miosqp_b200", "data")
    os.makedirs(data_dir, exist_ok=True)
    for tag in ("4", "510", "550"):
        mat = sio.loadmat(os.path.join(REF, "examples/power_converter/tail_backups/delta_%s.mat" % tag))
        np.savez(os.path.join(data_dir, "tail_delta_%s.npz" % tag), P0=np.asarray(mat["P0"], float),
                 q0=np.asarray(mat["q0"], float).reshape(-1), r0=np.asarray(mat["r0"], float).reshape(()))

    model = Model()
    model.set_params(25.0e-06, 50., 0.8e03, 0.8e03, 1.)
    model.set_time(0.0, 1, 2)
    model.set_initial_conditions()
    model.gen_dynamical_system(300, 5.5)
    model.gen_tail_cost(50, 0.95, name='delta_550.mat')
    ds = model.dyn_system
    out = dict(sys_A=np.asarray(ds.A), sys_B=np.asarray(ds.B), sys_C=np.asarray(ds.C), x0=np.asarray(model.init_conditions.x0),
               cur_step_torque=np.asarray(model.init_conditions.cur_step_torque), Nstpp=np.array(model.params.Nstpp),
               T_final=np.array(model.time.T_final), t=np.asarray(model.time.t))
    for N in (1, 3, 10):
        qp = MIQP(ds, N, model.tail_cost)
        out["P_%d" % N] = qp.P.toarray(); out["A_%d" % N] = qp.A.toarray()
        out["q_x_%d" % N] = np.asarray(qp.q_x); out["q_u_%d" % N] = np.asarray(qp.q_u)
        out["SA_%d" % N] = np.asarray(qp.SA_tilde); out["l_%d" % N] = qp.l; out["u_%d" % N] = qp.u
    path = os.path.join(HERE, "power_converter_model.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
