"""
Sequential `osqp`-shaped adapter over the CUDA engine: the six calls the reference makes on its solver object
(/root/reference/miosqp/workspace.py:63-68, node.py:102-125, solver.py:185) served by batches of ONE node.
It exists so the UNMODIFIED reference package can be pointed at the engine (tests/osqp_shim);
the product path is miosqp_b200.MIOSQP, which batches.
"""
from time import perf_counter

import numpy as np

from . import engine


def constant(name):
    return engine.CONSTANTS[name]


class _Info(object):
    pass


class _Results(object):
    pass


class OSQP(object):
    def __init__(self):
        self._qp = None

    def setup(self, P=None, q=None, A=None, l=None, u=None, **settings):
        self._qp = engine.BatchedQP().setup(P, q, A, l, u, **settings)
        self._l = np.array(l, dtype=np.float64); self._u = np.array(u, dtype=np.float64)
        self._x0 = np.zeros(self._qp.n); self._y0 = np.zeros(self._qp.m)

    def update(self, q=None, l=None, u=None, **kw):
        if kw:
            raise NotImplementedError("only update(q=, l=, u=) is supported")
        if q is not None:
            self._qp.update_q(q)
        if l is not None:
            self._l = np.array(l, dtype=np.float64)
        if u is not None:
            self._u = np.array(u, dtype=np.float64)
        if np.any(self._l > self._u):
            raise ValueError("Lower bound must be lower than or equal to upper bound")

    def warm_start(self, x=None, y=None):
        if x is not None:
            self._x0 = np.array(x, dtype=np.float64)
        if y is not None:
            self._y0 = np.array(y, dtype=np.float64)

    def solve(self):
        t0 = perf_counter()
        r = self._qp.solve_batch(self._l[None], self._u[None], self._x0[None], self._y0[None])
        out = _Results(); out.info = _Info()
        out.x, out.y = r.x[0], r.y[0]
        out.info.status_val = int(r.status[0]); out.info.iter = int(r.iters[0])
        out.info.obj_val = float(r.obj[0]); out.info.pri_res = float(r.pri_res[0]); out.info.dua_res = float(r.dua_res[0])
        out.info.run_time = out.info.solve_time = perf_counter() - t0
        # osqp keeps the last iterates as the next warm start; the reference always warm-starts explicitly
        if out.info.status_val in (1, 2, -2):
            self._x0, self._y0 = out.x.copy(), out.y.copy()
        return out
