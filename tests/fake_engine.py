"""
TEST-ONLY stand-in for the CUDA engine, backed by the CPU oracle, so the host-side B&B replay
(miosqp_b200/tree.py, miqp.py) can be exercised where there is no GPU.  It is installed by the
`cpu_engine` fixture (monkeypatching miosqp_b200.engine); nothing in the product imports it.
"""
import os

import numpy as np

from oracle import oracle

THREADS = int(os.environ.get("FAKE_ENGINE_THREADS", "4"))      # bench.py's CPU arm sets 1: one relaxation at a time


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

    def native_solve_fn(self):
        """bqp_solve_fn (include/bqp.h) backed by the oracle: lets the C++ B&B replay run where there is no GPU."""
        import ctypes as C
        from miosqp_b200 import engine
        n, m, o = self.n, self.m, self.o

        def fn(ctx, B, l, u, x0, y0, x, y, status, iters):
            try:
                arr = lambda pp, k, size: np.ctypeslib.as_array(pp[k], shape=(size,))
                L = np.array([arr(l, b, m) for b in range(B)]); U = np.array([arr(u, b, m) for b in range(B)])
                X0 = np.array([arr(x0, b, n) for b in range(B)]); Y0 = np.array([arr(y0, b, m) for b in range(B)])
                xs, ys, st, it, _ = o.solve_batch(L, U, X0, Y0, threads=THREADS)
                for b in range(B):
                    arr(x, b, n)[:] = xs[b]; arr(y, b, m)[:] = ys[b]; status[b] = int(st[b]); iters[b] = int(it[b])
                return 0
            except Exception:           # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return -1
        self._native_fn = engine.SOLVE_FN(fn)      # keep the callback object alive
        return self._native_fn

    def solve_batch(self, l, u, x0, y0):
        x, y, st, it, extra = self.o.solve_batch(np.atleast_2d(l), np.atleast_2d(u), np.atleast_2d(x0), np.atleast_2d(y0))
        r = Scalars()
        r.x, r.y, r.status, r.iters = x, y, st, it
        r.obj, r.pri_res, r.dua_res = extra["obj"], extra["pri_res"], extra["dua_res"]
        return r


def solve_multi(qps, l, u, x0, y0):
    xs, ys, st, it, _ = oracle.solve_multi([q.o for q in qps], l, u, x0, y0, threads=THREADS)
    sc = Scalars()
    sc.status = np.array(st); sc.iters = np.array(it)
    return xs, ys, sc


def native_solve_many_fn(qps):
    """bqp_solve_many_fn (include/bqp.h) backed by the oracle: node b belongs to problem owner[b]."""
    from miosqp_b200 import engine
    oracles = [q.o for q in qps]

    def fn(ctx, B, owner, l, u, x0, y0, x, y, status, iters):
        try:
            own = [int(owner[b]) for b in range(B)]
            arr = lambda pp, b, size: np.ctypeslib.as_array(pp[b], shape=(size,))
            os_ = [oracles[k] for k in own]
            L = [np.array(arr(l, b, os_[b].m)) for b in range(B)]; U = [np.array(arr(u, b, os_[b].m)) for b in range(B)]
            X0 = [np.array(arr(x0, b, os_[b].n)) for b in range(B)]; Y0 = [np.array(arr(y0, b, os_[b].m)) for b in range(B)]
            xs, ys, st, it, _ = oracle.solve_multi(os_, L, U, X0, Y0, threads=THREADS)
            for b in range(B):
                arr(x, b, os_[b].n)[:] = xs[b]; arr(y, b, os_[b].m)[:] = ys[b]; status[b] = int(st[b]); iters[b] = int(it[b])
            return 0
        except Exception:
            import traceback
            traceback.print_exc()
            return -1
    keep = engine.SOLVE_MANY_FN(fn)
    qps[0]._native_many_fn = keep          # keep the callback object alive
    return keep
