"""
oracle/oracle.py -- ctypes binding of the CPU ORACLE (TEST INFRASTRUCTURE).

Exposes the six-call `osqp` surface that the reference miOSQP uses
(/root/reference/miosqp/workspace.py:63-68, node.py:102-125, solver.py:185)
on top of oracle/osqp_oracle.c, plus batch entry points used as the CPU
baseline.  PARITY UNPINNED (see osqp_oracle.c header): `osqp` itself is not
available in this image.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.  The product package
(miosqp_b200/) never does.
"""
import ctypes as C
import os
import subprocess
from time import perf_counter

import numpy as np
import scipy.sparse as spa

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

OSQP_INFTY = 1e30
_CONSTANTS = {
    "OSQP_SOLVED": 1, "OSQP_SOLVED_INACCURATE": 2,
    "OSQP_PRIMAL_INFEASIBLE_INACCURATE": 3, "OSQP_DUAL_INFEASIBLE_INACCURATE": 4,
    "OSQP_MAX_ITER_REACHED": -2, "OSQP_PRIMAL_INFEASIBLE": -3, "OSQP_DUAL_INFEASIBLE": -4,
    "OSQP_SIGINT": -5, "OSQP_TIME_LIMIT_REACHED": -6, "OSQP_NON_CVX": -7, "OSQP_UNSOLVED": -10,
    "OSQP_INFTY": OSQP_INFTY, "OSQP_NAN": float("nan"),
}


def constant(name):
    return _CONSTANTS[name]


class _Settings(C.Structure):
    _fields_ = [("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
                ("eps_abs", C.c_double), ("eps_rel", C.c_double),
                ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
                ("max_iter", C.c_int), ("scaling", C.c_int), ("check_termination", C.c_int),
                ("scaled_termination", C.c_int), ("eq_rho", C.c_int),
                ("adaptive_rho", C.c_int), ("adaptive_rho_interval", C.c_int), ("adaptive_rho_tolerance", C.c_double)]


class _Info(C.Structure):
    _fields_ = [("status", C.c_int), ("iter", C.c_int), ("obj_val", C.c_double),
                ("pri_res", C.c_double), ("dua_res", C.c_double), ("solve_time", C.c_double),
                ("rho_updates", C.c_int), ("rho_final", C.c_double)]


DEFAULTS = dict(rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3, eps_prim_inf=1e-4,
                eps_dual_inf=1e-4, max_iter=4000, scaling=10, check_termination=25,
                scaled_termination=0, eq_rho=1, adaptive_rho=0, adaptive_rho_interval=0, adaptive_rho_tolerance=5.0)
_ALIASES = {"eps_inf": "eps_prim_inf", "eps_unb": "eps_dual_inf"}
_IGNORED = {"verbose", "polish", "polishing", "warm_start",
            "adaptive_rho_fraction", "time_limit", "linsys_solver",
            "delta", "polish_refine_iter", "pol_refine_iter", "scaling_iter", "scaling_norm",
            "early_terminate", "early_terminate_interval", "auto_rho", "scaled_termination_"}


def normalize_settings(kw):
    s = dict(DEFAULTS)
    for k, v in kw.items():
        k = _ALIASES.get(k, k)
        if k in s:
            s[k] = v
        elif k in _IGNORED:
            continue
        else:
            raise TypeError("unknown OSQP setting %r" % k)
    if isinstance(s["scaling"], bool):
        s["scaling"] = 10 if s["scaling"] else 0
    s["adaptive_rho"] = 1 if s["adaptive_rho"] else 0
    if s["adaptive_rho"] and not int(s["adaptive_rho_interval"]) > 0:
        # osqp's automatic interval (0) is derived from wall-clock setup time: not reproducible, refused as the engine does
        raise ValueError("adaptive_rho needs a fixed adaptive_rho_interval > 0 (the automatic, timing-based interval of osqp "
                         "is not reproducible and is outside the parity contract)")
    s["adaptive_rho_interval"] = int(s["adaptive_rho_interval"])
    return s


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.oracle_setup.restype = C.c_void_p
        L.oracle_setup.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, ip, ip, dp, dp, dp, C.POINTER(_Settings)]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_update_lin_cost.argtypes = [C.c_void_p, dp]
        L.oracle_update_bounds.argtypes = [C.c_void_p, dp, dp]
        L.oracle_warm_start.argtypes = [C.c_void_p, dp, dp]
        L.oracle_solve.argtypes = [C.c_void_p, dp, dp, C.POINTER(_Info)]
        L.oracle_solve_node.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, dp, C.POINTER(_Info)]
        L.oracle_solve_batch.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, dp, dp, dp, C.POINTER(_Info), C.c_int]
        L.oracle_solve_multi.argtypes = [C.c_int] + [C.POINTER(C.c_void_p)] * 7 + [C.POINTER(_Info), C.c_int]
        L.oracle_dims.argtypes = [C.c_void_p, ip, ip, ip, ip]
        L.oracle_get_scaling.argtypes = [C.c_void_p, dp, dp, dp]
        L.oracle_get_factor.argtypes = [C.c_void_p, ip, ip, ip, dp, dp]
        L.oracle_kkt_solve.argtypes = [C.c_void_p, dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class _ResInfo(object):
    pass


class _Results(object):
    pass


class OSQP(object):
    """`osqp.OSQP`-shaped object over the C oracle."""

    def __init__(self):
        self._w = None

    def __del__(self):
        if getattr(self, "_w", None):
            lib().oracle_free(self._w)
            self._w = None

    def setup(self, P=None, q=None, A=None, l=None, u=None, **settings):
        s = normalize_settings(settings)
        self.settings = s
        P = spa.triu(spa.csc_matrix(P), format="csc")
        A = spa.csc_matrix(A)
        P.sort_indices(); A.sort_indices()
        self.n, self.m = A.shape[1], A.shape[0]
        q = _f64(q); l = _f64(l); u = _f64(u)
        if np.any(l > u):
            raise ValueError("Lower bound must be lower than or equal to upper bound")
        st = _Settings(**{k: s[k] for k, _ in _Settings._fields_})
        Pp, Pi, Px = P.indptr.astype(np.int32), P.indices.astype(np.int32), _f64(P.data)
        Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), _f64(A.data)
        t0 = perf_counter()
        self._w = lib().oracle_setup(self.n, self.m, _i(Pp), _i(Pi), _d(Px), _d(q), _i(Ap), _i(Ai), _d(Ax),
                                     _d(l), _d(u), C.byref(st))
        self.setup_time = perf_counter() - t0
        if not self._w:
            raise ValueError("oracle setup failed (non-convex or singular KKT)")
        self._first = True

    def update(self, q=None, l=None, u=None, **kw):
        if kw:
            raise NotImplementedError("oracle supports update(q=, l=, u=) only")
        if q is not None:
            q = _f64(q)
            if q.shape[0] != self.n:
                raise ValueError("q must have length n")
            lib().oracle_update_lin_cost(self._w, _d(q))
        if l is not None or u is not None:
            if l is None or u is None:
                raise NotImplementedError("oracle update needs both l and u")
            l = _f64(l); u = _f64(u)
            if l.shape[0] != self.m or u.shape[0] != self.m:
                raise ValueError("l/u must have length m")
            if lib().oracle_update_bounds(self._w, _d(l), _d(u)):
                raise ValueError("Lower bound must be lower than or equal to upper bound")

    def warm_start(self, x=None, y=None):
        x = _f64(x if x is not None else np.zeros(self.n))
        y = _f64(y if y is not None else np.zeros(self.m))
        lib().oracle_warm_start(self._w, _d(x), _d(y))

    def solve(self):
        x = np.empty(self.n); y = np.empty(self.m); info = _Info()
        lib().oracle_solve(self._w, _d(x), _d(y), C.byref(info))
        return _pack(x, y, info)

    # ---- stateless / batch entry points (CPU baseline + parity checks)
    def solve_node(self, l, u, x0, y0):
        x = np.empty(self.n); y = np.empty(self.m); info = _Info()
        rc = lib().oracle_solve_node(self._w, _d(_f64(l)), _d(_f64(u)), _d(_f64(x0)), _d(_f64(y0)),
                                     _d(x), _d(y), C.byref(info))
        if rc:
            raise ValueError("Lower bound must be lower than or equal to upper bound")
        return _pack(x, y, info)

    def solve_batch(self, l, u, x0, y0, threads=1):
        l = _f64(l); u = _f64(u); x0 = _f64(x0); y0 = _f64(y0)
        B = l.shape[0]
        x = np.empty((B, self.n)); y = np.empty((B, self.m)); infos = (_Info * B)()
        rc = lib().oracle_solve_batch(self._w, B, _d(l), _d(u), _d(x0), _d(y0), _d(x), _d(y), infos, threads)
        if rc:
            raise ValueError("Lower bound must be lower than or equal to upper bound")
        status = np.array([i.status for i in infos]); iters = np.array([i.iter for i in infos])
        extra = dict(obj=np.array([i.obj_val for i in infos]), pri_res=np.array([i.pri_res for i in infos]),
                     dua_res=np.array([i.dua_res for i in infos]), solve_time=np.array([i.solve_time for i in infos]),
                     rho_updates=np.array([i.rho_updates for i in infos]), rho_final=np.array([i.rho_final for i in infos]))
        return x, y, status, iters, extra

    def dims(self):
        n, m, N, nnz = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        lib().oracle_dims(self._w, C.byref(n), C.byref(m), C.byref(N), C.byref(nnz))
        return n.value, m.value, N.value, nnz.value

    def scaling(self):
        D = np.empty(self.n); E = np.empty(self.m); c = C.c_double()
        lib().oracle_get_scaling(self._w, _d(D), _d(E), C.byref(c))
        return D, E, c.value

    def factor(self):
        n, m, N, nnz = self.dims()
        perm = np.empty(N, np.int32); Lp = np.empty(N + 1, np.int32); Li = np.empty(max(nnz, 1), np.int32)
        Lx = np.empty(max(nnz, 1)); Dd = np.empty(N)
        lib().oracle_get_factor(self._w, _i(perm), _i(Lp), _i(Li), _d(Lx), _d(Dd))
        return perm, Lp, Li[:nnz], Lx[:nnz], Dd

    def kkt_solve(self, b):
        b = _f64(b).copy()
        lib().oracle_kkt_solve(self._w, _d(b))
        return b


def solve_multi(solvers, l, u, x0, y0, threads=1):
    """Nodes of different oracle instances in one thread pool.  All args are lists of length B."""
    B = len(solvers)
    l = [_f64(a) for a in l]; u = [_f64(a) for a in u]; x0 = [_f64(a) for a in x0]; y0 = [_f64(a) for a in y0]
    x = [np.empty(s.n) for s in solvers]; y = [np.empty(s.m) for s in solvers]
    infos = (_Info * B)()

    def parr(arrs):
        return (C.c_void_p * B)(*[a.ctypes.data for a in arrs])
    ws = (C.c_void_p * B)(*[s._w for s in solvers])
    rc = lib().oracle_solve_multi(B, ws, parr(l), parr(u), parr(x0), parr(y0), parr(x), parr(y), infos, threads)
    if rc:
        raise ValueError("Lower bound must be lower than or equal to upper bound")
    return x, y, [i.status for i in infos], [i.iter for i in infos], [i.solve_time for i in infos]


def _pack(x, y, info):
    r = _Results(); r.x = x; r.y = y; r.info = _ResInfo()
    r.info.status_val = info.status; r.info.iter = info.iter; r.info.obj_val = info.obj_val
    r.info.pri_res = info.pri_res; r.info.dua_res = info.dua_res
    r.info.run_time = info.solve_time; r.info.solve_time = info.solve_time
    return r
