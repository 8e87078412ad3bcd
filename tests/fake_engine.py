"""
TEST-ONLY stand-in for the CUDA engine, backed by the CPU oracle, so the host-side B&B replay
(miosqp_b200/tree.py, miqp.py) can be exercised where there is no GPU.  It is installed by the
`cpu_engine` fixture (monkeypatching miosqp_b200.engine); nothing in the product imports it.
"""
import numpy as np

from oracle import oracle


class Scalars(object):
    pass


class FakeBatchedQP(object):
    def __init__(self):
        self.o = None

    def setup(self, P, q, A, l, u, i_idx=None, **settings):
        settings.pop("device", None)
        self.o = oracle.OSQP()
        self.o.setup(P, q, A, l, u, **settings)
        self.n, self.m = self.o.n, self.o.m
        return self

    def update_q(self, q):
        self.o.update(q=q)

    def free(self):
        self.o = None

    def solve_batch(self, l, u, x0, y0):
        x, y, st, it, extra = self.o.solve_batch(np.atleast_2d(l), np.atleast_2d(u), np.atleast_2d(x0), np.atleast_2d(y0))
        r = Scalars()
        r.x, r.y, r.status, r.iters = x, y, st, it
        r.obj, r.pri_res, r.dua_res = extra["obj"], extra["pri_res"], extra["dua_res"]
        return r


def solve_multi(qps, l, u, x0, y0):
    xs, ys, st, it, _ = oracle.solve_multi([q.o for q in qps], l, u, x0, y0, threads=4)
    sc = Scalars()
    sc.status = np.array(st); sc.iters = np.array(it)
    return xs, ys, sc
