"""
ctypes binding of libbqp.so (include/bqp.h) -- the B200 batched QP-relaxation engine.

`BatchedQP` is one set-up problem (what `osqp.OSQP().setup(...)` is for the reference,
/root/reference/miosqp/workspace.py:63-68); `solve_batch` / `solve_multi` run the ADMM loops of
many B&B nodes in one kernel launch (what the reference does one node at a time in
/root/reference/miosqp/node.py:96-143).  There is no CPU fallback: if the library cannot be
built or no B200 is visible, calls raise.
"""
import ctypes as C
import os

# More hardware work queues than the default 8: tiles launched from different solve contexts (streams) must not queue
# behind each other (tools/ctx_concurrency.py: 64 contexts take 101 ms per round with 8 queues, 63 ms with 32).  Read by the
# CUDA driver when the process initialises CUDA, so it has to be in the environment before that.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np
import scipy.sparse as spa

from . import build as _build

OSQP_INFTY = 1e30
CONSTANTS = {
    "OSQP_SOLVED": 1, "OSQP_SOLVED_INACCURATE": 2,
    "OSQP_PRIMAL_INFEASIBLE_INACCURATE": 3, "OSQP_DUAL_INFEASIBLE_INACCURATE": 4,
    "OSQP_MAX_ITER_REACHED": -2, "OSQP_PRIMAL_INFEASIBLE": -3, "OSQP_DUAL_INFEASIBLE": -4,
    "OSQP_SIGINT": -5, "OSQP_TIME_LIMIT_REACHED": -6, "OSQP_NON_CVX": -7, "OSQP_UNSOLVED": -10,
    "OSQP_INFTY": OSQP_INFTY, "OSQP_NAN": float("nan"),
}

# OSQP setting names understood by the engine, with OSQP's defaults
DEFAULTS = dict(rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3, eps_prim_inf=1e-4,
                eps_dual_inf=1e-4, max_iter=4000, scaling=10, check_termination=25, eq_rho=1, device=0,
                adaptive_rho=0, adaptive_rho_interval=0, adaptive_rho_tolerance=5.0)
# pre-0.1.3 names found in /root/reference/max_iter_examples/*.pickle
ALIASES = {"eps_inf": "eps_prim_inf", "eps_unb": "eps_dual_inf"}
# accepted and ignored (no effect on the iterates of the parity contract)
IGNORED = {"verbose", "polish", "polishing", "warm_start", "time_limit", "linsys_solver", "delta",
           "polish_refine_iter", "pol_refine_iter", "scaling_iter", "scaling_norm", "early_terminate",
           "early_terminate_interval", "auto_rho", "adaptive_rho_fraction"}


class BqpError(RuntimeError):
    pass


class _Settings(C.Structure):
    _fields_ = [("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
                ("eps_abs", C.c_double), ("eps_rel", C.c_double),
                ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
                ("max_iter", C.c_int), ("scaling", C.c_int), ("check_termination", C.c_int),
                ("eq_rho", C.c_int), ("device", C.c_int),
                ("adaptive_rho", C.c_int), ("adaptive_rho_interval", C.c_int), ("adaptive_rho_tolerance", C.c_double)]


_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


class _Problem(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int),
                ("Pp", _ip), ("Pi", _ip), ("Px", _dp),
                ("Ap", _ip), ("Ai", _ip), ("Ax", _dp),
                ("q", _dp), ("l", _dp), ("u", _dp),
                ("n_int", C.c_int), ("i_idx", _ip)]


class _BnbSettings(C.Structure):
    _fields_ = [("eps_int_feas", C.c_double), ("max_iter_bb", C.c_int), ("tree_explor_rule", C.c_int),
                ("branching_rule", C.c_int), ("speculation", C.c_int), ("eps_abs", C.c_double)]


class _BnbResult(C.Structure):
    _fields_ = [("status", C.c_int), ("iter_num", C.c_int), ("osqp_iter", C.c_longlong),
                ("osqp_solve_time", C.c_double), ("upper_glob", C.c_double), ("lower_glob", C.c_double),
                ("batches", C.c_int), ("batched_nodes", C.c_longlong), ("spec_nodes", C.c_longlong),
                ("spec_hits", C.c_longlong), ("n_decisions", C.c_int), ("open_leaves", C.c_int)]


_pp_d = C.POINTER(C.POINTER(C.c_double))
# bqp_solve_fn of include/bqp.h: a batch solver that stands in for the engine (CPU tests drive the native replay with the oracle)
SOLVE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, _pp_d, _pp_d, _pp_d, _pp_d, _pp_d, _pp_d, C.POINTER(C.c_int), C.POINTER(C.c_int))


SOLVE_MANY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), _pp_d, _pp_d, _pp_d, _pp_d, _pp_d, _pp_d,
                            C.POINTER(C.c_int), C.POINTER(C.c_int))


class _NodeOut(C.Structure):
    _fields_ = [("status", _ip), ("iters", _ip), ("obj", _dp), ("pri_res", _dp), ("dua_res", _dp), ("lower", _dp)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
                ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong),
                ("launches", C.c_int), ("tiles", C.c_int), ("tile_nodes", C.c_int), ("threads", C.c_int),
                ("smem_bytes", C.c_longlong), ("node_iters", C.c_longlong), ("tile_iters", C.c_longlong),
                ("stream_bytes", C.c_longlong), ("kernel", C.c_int), ("ring_slots", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = ["bqp_default_settings", "bqp_setup", "bqp_update_q", "bqp_solve_batch", "bqp_solve_multi",
           "bqp_batch_upload", "bqp_batch_run", "bqp_batch_download", "bqp_last_timing", "bqp_free",
           "bqp_set_tuning", "bqp_get_dims", "bqp_get_scaling", "bqp_device_count", "bqp_strerror",
           "bqp_version", "bqp_debug_host_setup", "bqp_debug_host_kkt_solve", "bqp_debug_host_stream_kkt_solve",
           "bqp_debug_host_panel_kkt_solve", "bqp_debug_host_small_kkt_solve",
           "bqp_debug_host_matvec", "bqp_bnb_solve", "bqp_bnb_solve_many", "bqp_setup_many", "bqp_bnb_solve_async",
           "bqp_ctx_create", "bqp_ctx_free", "bqp_ctx_solve_multi", "bqp_ctx_last_timing", "bqp_handle_device",
           "bqp_get_inverse_guard", "bqp_ctx_set_auto_cluster", "bqp_ctx_set_sm_share", "bqp_bnb_solve_rolling", "bqp_session_begin", "bqp_session_append", "bqp_session_round", "bqp_session_fetch"]

_lib = None


def lib():
    """Load (building first when stale) libbqp.so.  Raises if it cannot be produced."""
    global _lib
    if _lib is None:
        path = _build.build()
        L = C.CDLL(path)
        vp = C.c_void_p
        pp = C.POINTER(C.c_void_p)
        L.bqp_default_settings.argtypes = [C.POINTER(_Settings)]
        L.bqp_setup.argtypes = [C.POINTER(_Problem), C.POINTER(_Settings), pp]
        L.bqp_debug_host_setup.argtypes = [C.POINTER(_Problem), C.POINTER(_Settings), pp]
        L.bqp_update_q.argtypes = [vp, _dp]
        L.bqp_solve_batch.argtypes = [vp, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, C.POINTER(_NodeOut)]
        L.bqp_solve_multi.argtypes = [C.c_int, pp, pp, pp, pp, pp, pp, pp, C.POINTER(_NodeOut)]
        L.bqp_batch_upload.argtypes = [C.c_int, pp, pp, pp, pp, pp]
        L.bqp_batch_run.argtypes = []
        L.bqp_batch_download.argtypes = [pp, pp, C.POINTER(_NodeOut)]
        L.bqp_last_timing.argtypes = [C.POINTER(Timing)]
        L.bqp_free.argtypes = [vp]
        L.bqp_set_tuning.argtypes = [C.c_int, C.c_int]
        L.bqp_get_dims.argtypes = [vp, _ip, _ip, _ip, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        L.bqp_get_scaling.argtypes = [vp, _dp, _dp, _dp]
        L.bqp_strerror.restype = C.c_char_p
        L.bqp_strerror.argtypes = [C.c_int]
        L.bqp_version.restype = C.c_char_p
        L.bqp_debug_host_kkt_solve.argtypes = [vp, _dp]
        L.bqp_debug_host_stream_kkt_solve.argtypes = [vp, _dp]
        L.bqp_debug_host_panel_kkt_solve.argtypes = [vp, _dp]
        L.bqp_debug_host_small_kkt_solve.argtypes = [vp, _dp]
        L.bqp_debug_host_matvec.argtypes = [vp, C.c_int, _dp, _dp]
        L.bqp_setup_many.argtypes = [C.c_int, C.POINTER(C.POINTER(_Problem)), C.POINTER(_Settings), pp, C.c_int, C.c_int]
        L.bqp_bnb_solve_many.argtypes = [C.c_int, pp, C.POINTER(C.POINTER(_Problem)), C.POINTER(_BnbSettings), _pp_d, _dp, C.c_void_p, vp,
                                         _pp_d, C.POINTER(_BnbResult), C.POINTER(_ip), C.c_int]
        L.bqp_bnb_solve_async.argtypes = [C.c_int, pp, C.POINTER(C.POINTER(_Problem)), C.POINTER(_BnbSettings), _pp_d, _dp,
                                          _pp_d, C.POINTER(_BnbResult), C.POINTER(_ip), C.c_int, C.c_int]
        L.bqp_bnb_solve_rolling.argtypes = [C.c_int, pp, C.POINTER(C.POINTER(_Problem)), C.POINTER(_BnbSettings), _pp_d, _dp,
                                            _pp_d, C.POINTER(_BnbResult), C.POINTER(_ip), C.c_int, _ip, C.c_int]
        L.bqp_ctx_set_sm_share.argtypes = [vp, C.c_int]
        L.bqp_session_begin.argtypes = [vp]
        L.bqp_session_append.argtypes = [vp, C.c_int, pp, pp, pp, pp, pp, _ip]
        L.bqp_session_round.argtypes = [vp, _ip, C.c_int, _ip, _ip]
        L.bqp_session_fetch.argtypes = [vp, C.c_int, _dp, _dp, C.POINTER(_NodeOut)]
        L.bqp_ctx_create.argtypes = [C.c_int, C.c_int, pp]
        L.bqp_ctx_free.argtypes = [vp]
        L.bqp_ctx_solve_multi.argtypes = [vp, C.c_int, pp, pp, pp, pp, pp, pp, pp, C.POINTER(_NodeOut)]
        L.bqp_ctx_last_timing.argtypes = [vp, C.POINTER(Timing)]
        L.bqp_handle_device.argtypes = [vp]
        L.bqp_ctx_set_auto_cluster.argtypes = [vp, C.c_int]
        L.bqp_get_inverse_guard.argtypes = [vp, _dp, _ip]
        L.bqp_bnb_solve.argtypes = [vp, C.POINTER(_Problem), C.POINTER(_BnbSettings), _dp, C.c_double, C.c_void_p, vp, _dp,
                                    C.POINTER(_BnbResult), _ip, C.c_int]
        _lib = L
    return _lib


def _check(rc):
    if rc == 0:
        return
    msg = lib().bqp_strerror(rc).decode()
    if rc in (-1, -2, -3, -6, -7, -8):     # osqp / miosqp raise ValueError for bad data, l > u, non-convex, unknown rules
        raise ValueError(msg)
    raise BqpError(msg)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def normalize_settings(kw):
    s = dict(DEFAULTS)
    for k, v in kw.items():
        k = ALIASES.get(k, k)
        if k in s:
            s[k] = v
        elif k == "scaled_termination":
            if v:
                raise ValueError("scaled_termination is not supported")
        elif k in IGNORED:
            continue
        else:
            raise TypeError("unknown OSQP setting %r" % k)
    if s["eq_rho"] not in (0, 1, 2, False, True):
        raise ValueError("eq_rho must be 0, 1 (rho typed per row once at setup: the default contract) or 2 (the integer-bound rows "
                         "re-typed per node, what osqp >= 0.4 does in update_bounds; dense-A problems only)")
    if isinstance(s["scaling"], bool):
        s["scaling"] = 10 if s["scaling"] else 0
    s["adaptive_rho"] = 1 if s["adaptive_rho"] else 0
    s["adaptive_rho_interval"] = int(s["adaptive_rho_interval"])
    if s["adaptive_rho"]:
        # osqp's default (interval 0) derives the interval from wall-clock setup time: not reproducible, so a node would not
        # be a pure function of (l, u, x0, y0) -- refused; a fixed interval on the termination-check grid is the contract
        if s["adaptive_rho_interval"] <= 0 or s["adaptive_rho_interval"] % int(s["check_termination"]) != 0:
            raise ValueError("adaptive_rho needs a fixed adaptive_rho_interval > 0 that is a multiple of check_termination "
                             "(osqp's automatic, timing-based interval is not reproducible)")
        if s["eq_rho"] == 2:
            raise ValueError("adaptive_rho and eq_rho = 2 (per-node re-typing) cannot be combined")
    return s


def set_tuning(tile_nodes=0, threads=0):
    _check(lib().bqp_set_tuning(int(tile_nodes), int(threads)))


def set_auto_cluster(on, ctx=None):
    """Let launches with few tiles spread each tile over 4 or 8 CTAs (include/bqp.h bqp_ctx_set_auto_cluster): lower
    per-iteration latency for single trees, results equal to rounding instead of bit for bit.  ctx None: the default context."""
    _check(lib().bqp_ctx_set_auto_cluster(ctx._c if ctx is not None else None, 1 if on else 0))


def device_count():
    return lib().bqp_device_count()


def last_timing():
    t = Timing()
    _check(lib().bqp_last_timing(C.byref(t)))
    return t.as_dict()


class BatchResult(object):
    """Per-node outputs of one batched solve (arrays of length B along axis 0)."""
    __slots__ = ("x", "y", "status", "iters", "obj", "pri_res", "dua_res", "lower")


class BatchedQP(object):
    """One set-up QP family: P, A fixed; q updatable; l,u,x0,y0 per node."""

    def __init__(self):
        self._h = C.c_void_p()
        self.n = self.m = 0

    def __del__(self):
        self.free()

    def free(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().bqp_free(self._h)
            self._h = C.c_void_p()

    def setup(self, P, q, A, l, u, i_idx=None, host_only=False, **settings):
        s = normalize_settings(settings)
        self.settings = s
        P = spa.triu(spa.csc_matrix(P), format="csc")
        A = spa.csc_matrix(A)
        P.sort_indices()
        A.sort_indices()
        self.n, self.m = A.shape[1], A.shape[0]
        if P.shape != (self.n, self.n):
            raise ValueError("P must be n x n")
        q = _f64(q); l = _f64(l); u = _f64(u)
        if q.shape != (self.n,) or l.shape != (self.m,) or u.shape != (self.m,):
            raise ValueError("q, l, u have inconsistent dimensions")
        idx = np.ascontiguousarray(np.asarray(i_idx if i_idx is not None else [], dtype=np.int32))
        Pp, Pi, Px = P.indptr.astype(np.int32), P.indices.astype(np.int32), _f64(P.data)
        Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), _f64(A.data)
        prob = _Problem(self.n, self.m, _i(Pp), _i(Pi), _d(Px), _i(Ap), _i(Ai), _d(Ax), _d(q), _d(l), _d(u),
                        int(idx.size), _i(idx))
        st = _Settings(**{k: s[k] for k, _ in _Settings._fields_})
        self.free()
        fn = lib().bqp_debug_host_setup if host_only else lib().bqp_setup
        _check(fn(C.byref(prob), C.byref(st), C.byref(self._h)))
        self.n_int = int(idx.size)
        return self

    def native_solve_fn(self):
        """None: the native B&B replay solves its nodes on the device through this handle."""
        return None

    def update_q(self, q):
        q = _f64(q)
        if q.shape != (self.n,):
            raise ValueError("q must have length n")
        _check(lib().bqp_update_q(self._h, _d(q)))

    def dims(self):
        n, m, npad = C.c_int(), C.c_int(), C.c_int()
        fb, cb = C.c_longlong(), C.c_longlong()
        _check(lib().bqp_get_dims(self._h, C.byref(n), C.byref(m), C.byref(npad), C.byref(fb), C.byref(cb)))
        return dict(n=n.value, m=m.value, npad=npad.value, factor_bytes=fb.value, check_bytes=cb.value)

    def inverse_guard(self):
        """(error, in_use): how well KKT solves through the explicit reduced inverse of the dense-A kernels agree with the LDL'
        factor (largest relative difference over 4 probes), and whether the dense layout is in use (include/bqp.h)."""
        err = C.c_double(); use = C.c_int()
        _check(lib().bqp_get_inverse_guard(self._h, C.byref(err), C.byref(use)))
        return err.value, bool(use.value)

    def scaling(self):
        D = np.empty(self.n); E = np.empty(self.m); c = C.c_double()
        _check(lib().bqp_get_scaling(self._h, _d(D), _d(E), C.byref(c)))
        return D, E, c.value

    def solve_batch(self, l, u, x0, y0):
        """l,u,y0: [B][m]; x0: [B][n].  One launch for all B nodes."""
        l = np.atleast_2d(_f64(l)); u = np.atleast_2d(_f64(u))
        x0 = np.atleast_2d(_f64(x0)); y0 = np.atleast_2d(_f64(y0))
        B = l.shape[0]
        if l.shape != (B, self.m) or u.shape != (B, self.m) or y0.shape != (B, self.m) or x0.shape != (B, self.n):
            raise ValueError("batch arrays have inconsistent dimensions")
        r = _alloc_result(B, self.n, self.m)
        out = _node_out(r)
        _check(lib().bqp_solve_batch(self._h, B, _d(l), _d(u), _d(x0), _d(y0), _d(r.x), _d(r.y), C.byref(out)))
        return r

    # host-only debug hooks (layout tests)
    def debug_kkt_solve(self, rhs):
        b = _f64(rhs).copy()
        _check(lib().bqp_debug_host_kkt_solve(self._h, _d(b)))
        return b

    def debug_stream_kkt_solve(self, rhs):
        b = _f64(rhs).copy()
        _check(lib().bqp_debug_host_stream_kkt_solve(self._h, _d(b)))
        return b

    def debug_panel_kkt_solve(self, rhs):
        b = _f64(rhs).copy()
        _check(lib().bqp_debug_host_panel_kkt_solve(self._h, _d(b)))
        return b

    def debug_small_kkt_solve(self, rhs):
        b = _f64(rhs).copy()
        _check(lib().bqp_debug_host_small_kkt_solve(self._h, _d(b)))
        return b

    def debug_matvec(self, which, v):
        v = _f64(v)
        out = np.zeros({0: self.m, 1: self.n, 2: self.n, 3: self.n, 4: self.n, 5: self.n}[which])
        _check(lib().bqp_debug_host_matvec(self._h, which, _d(v), _d(out)))
        return out


def setup_many(items, host_only=False, threads=0, **settings):
    """`BatchedQP().setup(...)` for many problems that share their OSQP settings (include/bqp.h bqp_setup_many): the host
    halves (scaling, factorisation, layouts) run on `threads` host threads (0 = all), the device uploads one by one.
    items: sequence of (P, q, A, l, u, i_idx or None).  Returns the list of set-up BatchedQP objects."""
    s = normalize_settings(settings)
    probs, keep, qps = [], [], []
    for (P, q, A, l, u, i_idx) in items:
        P = spa.triu(spa.csc_matrix(P), format="csc"); A = spa.csc_matrix(A)
        P.sort_indices(); A.sort_indices()
        n, m = A.shape[1], A.shape[0]
        if P.shape != (n, n):
            raise ValueError("P must be n x n")
        q = _f64(q); l = _f64(l); u = _f64(u)
        if q.shape != (n,) or l.shape != (m,) or u.shape != (m,):
            raise ValueError("q, l, u have inconsistent dimensions")
        idx = np.ascontiguousarray(np.asarray(i_idx if i_idx is not None else [], dtype=np.int32))
        arrs = [P.indptr.astype(np.int32), P.indices.astype(np.int32), _f64(P.data),
                A.indptr.astype(np.int32), A.indices.astype(np.int32), _f64(A.data), q, l, u, idx]
        keep.append(arrs)
        probs.append(_Problem(n, m, _i(arrs[0]), _i(arrs[1]), _d(arrs[2]), _i(arrs[3]), _i(arrs[4]), _d(arrs[5]),
                              _d(q), _d(l), _d(u), int(idx.size), _i(idx)))
        qp = BatchedQP(); qp.settings = s; qp.n, qp.m, qp.n_int = n, m, int(idx.size)
        qps.append(qp)
    count = len(probs)
    if count == 0:
        return []
    pptr = (C.POINTER(_Problem) * count)(*[C.pointer(p) for p in probs])
    handles = (C.c_void_p * count)()
    st = _Settings(**{k: s[k] for k, _ in _Settings._fields_})
    _check(lib().bqp_setup_many(count, pptr, C.byref(st), handles, int(threads), 1 if host_only else 0))
    for qp, h in zip(qps, handles):
        qp._h = C.c_void_p(h)
    return qps


def native_solve_many_fn(qps):
    """None: the lock-step native replay solves its nodes on the device through the handles of `qps`
    (the CPU tests replace this hook with an oracle-backed bqp_solve_many_fn)."""
    return None


def _bnb_problem(data):
    """(bqp_problem, keep-alive arrays) of a problem_data.Data: the UNSCALED problem the replay evaluates objectives on."""
    fixed = getattr(data, "_native_csc", None)          # P, A never change after setup: convert once per Data
    if fixed is None:
        P = spa.csc_matrix(data.P); A = spa.csc_matrix(data.A)
        fixed = (A.shape, P.indptr.astype(np.int32), P.indices.astype(np.int32), _f64(P.data),
                 A.indptr.astype(np.int32), A.indices.astype(np.int32), _f64(A.data),
                 np.ascontiguousarray(np.asarray(data.i_idx, dtype=np.int32)))
        data._native_csc = fixed
    m_ext, n = fixed[0]
    keep = list(fixed[1:7]) + [_f64(data.q), _f64(data.l), _f64(data.u), fixed[7]]
    prob = _Problem(n, m_ext, _i(keep[0]), _i(keep[1]), _d(keep[2]), _i(keep[3]), _i(keep[4]), _d(keep[5]),
                    _d(keep[6]), _d(keep[7]), _d(keep[8]), int(keep[9].size), _i(keep[9]))
    return prob, keep


def _bnb_settings(settings, eps_abs):
    return _BnbSettings(float(settings['eps_int_feas']), int(settings['max_iter_bb']), int(settings['tree_explor_rule']),
                        int(settings['branching_rule']), int(settings.get('speculation', 0) or 0), float(eps_abs))


def bnb_solve(qp, data, settings, eps_abs, x_incumbent=None, upper_incumbent=np.inf):
    """The native B&B replay (include/bqp.h bqp_bnb_solve, csrc/bqp_bnb.cpp) on one set-up problem.
    `qp` is the BatchedQP holding the factor (or a stand-in exposing `native_solve_fn()`, CPU tests);
    `data` the MIQP's problem_data.Data; `settings` the reference's MIOSQP settings dict (+ 'speculation').
    Returns (x, result dict, decisions)."""
    prob, keep = _bnb_problem(data)
    st = _bnb_settings(settings, eps_abs)
    fn = qp.native_solve_fn()
    handle = getattr(qp, "_h", None) if fn is None else None
    x = np.empty(prob.n); res = _BnbResult()
    cap = max(1, int(settings['max_iter_bb']))
    dec = np.zeros(2 * cap, dtype=np.int32)
    xin = _f64(x_incumbent) if x_incumbent is not None and np.isfinite(upper_incumbent) else None
    auto = fn is None and bool(settings.get('cluster_auto', True))
    if auto:        # a single tree offers two leaves per step: spread its one tile over more SMs (see set_auto_cluster)
        set_auto_cluster(True)
    try:
        rc = _bnb_solve_call(handle, prob, st, xin, upper_incumbent, fn, x, res, dec, cap)
    finally:
        if auto:
            set_auto_cluster(False)
    _check(rc)
    out = {k: getattr(res, k) for k, _ in _BnbResult._fields_}
    decisions = [(int(dec[2 * k]), int(dec[2 * k + 1])) for k in range(min(res.n_decisions, cap))]
    return x, out, decisions


def _bnb_solve_call(handle, prob, st, xin, upper_incumbent, fn, x, res, dec, cap):
    return lib().bqp_bnb_solve(handle, C.byref(prob), C.byref(st), _d(xin) if xin is not None else None, float(upper_incumbent),
                             C.cast(fn, C.c_void_p) if fn is not None else None, None, _d(x), C.byref(res), _i(dec), cap)


def bnb_solve_many(qps, datas, settings, eps_abs, x_incumbents, upper_incumbents, many_fn=None, async_threads=None, rolling=False):
    """Native replay over several set-up problems.  Lock-step (bqp_bnb_solve_many): one launch per B&B step covers all
    their frontiers.  async_threads is not None (bqp_bnb_solve_async, 0 = automatic): every problem advances at its own
    pace on its own CUDA stream, driven by a pool of host threads.  Arguments are sequences of equal length; `many_fn`
    (tests) replaces the engine.  Returns a list of (x, result dict, decisions)."""
    count = len(qps)
    built = [_bnb_problem(d) for d in datas]
    pptr = (C.POINTER(_Problem) * count)(*[C.pointer(b[0]) for b in built])
    sts = (_BnbSettings * count)(*[_bnb_settings(s, e) for s, e in zip(settings, eps_abs)])
    cap = max(1, max(int(s['max_iter_bb']) for s in settings))
    xs = [np.empty(b[0].n) for b in built]
    decs = [np.zeros(2 * cap, dtype=np.int32) for _ in range(count)]
    res = (_BnbResult * count)()
    xins = [(_f64(x) if x is not None and np.isfinite(u) else None) for x, u in zip(x_incumbents, upper_incumbents)]
    xin_ptrs = (_dp * count)(*[(_d(x) if x is not None else None) for x in xins])
    uppers = _f64(np.asarray(upper_incumbents, dtype=np.float64))
    handles = None if many_fn is not None else (C.c_void_p * count)(*[q._h.value for q in qps])
    x_ptrs = (_dp * count)(*[_d(x) for x in xs])
    dec_ptrs = (_ip * count)(*[_i(d) for d in decs])
    if rolling and many_fn is None:
        # one engine session shared by all trees: a tree whose leaves have terminated rejoins the next round (bqp_bnb_solve_rolling)
        nr = C.c_int(0)
        rc = lib().bqp_bnb_solve_rolling(count, handles, pptr, sts, xin_ptrs, _d(uppers), x_ptrs, res, dec_ptrs, cap, C.byref(nr),
                                         int(rolling) if rolling is not True else 0)
    elif async_threads is not None and many_fn is None:
        rc = lib().bqp_bnb_solve_async(count, handles, pptr, sts, xin_ptrs, _d(uppers), x_ptrs, res, dec_ptrs, cap, int(async_threads))
    else:
        rc = lib().bqp_bnb_solve_many(count, handles, pptr, sts, xin_ptrs, _d(uppers),
                                      C.cast(many_fn, C.c_void_p) if many_fn is not None else None, None, x_ptrs, res, dec_ptrs, cap)
    _check(rc)
    out = []
    for k in range(count):
        r = {name: getattr(res[k], name) for name, _ in _BnbResult._fields_}
        out.append((xs[k], r, [(int(decs[k][2 * j]), int(decs[k][2 * j + 1])) for j in range(min(r["n_decisions"], cap))]))
    return out


class Context(object):
    """One solve context (include/bqp.h bqp_ctx): its own CUDA stream and staging buffers, so several host threads can
    each run `solve_multi` on the same device at the same time (ctypes releases the GIL during the call)."""

    def __init__(self, device=0, run_to_completion=True):
        self._c = C.c_void_p()
        _check(lib().bqp_ctx_create(int(device), 1 if run_to_completion else 0, C.byref(self._c)))

    def free(self):
        if getattr(self, "_c", None) is not None and self._c.value:
            lib().bqp_ctx_free(self._c)
            self._c = None

    def __del__(self):
        self.free()

    def solve_multi(self, qps, l, u, x0, y0):
        B = len(qps)
        res = [_alloc_result(1, q.n, q.m) for q in qps]
        sc = _Scalars()
        sc.status = np.empty(B, dtype=np.int32); sc.iters = np.empty(B, dtype=np.int32)
        sc.obj = np.empty(B); sc.pri_res = np.empty(B); sc.dua_res = np.empty(B); sc.lower = np.empty(B)
        ls = [_f64(v) for v in l]; us = [_f64(v) for v in u]; xs = [_f64(v) for v in x0]; ys = [_f64(v) for v in y0]
        out = _NodeOut(_i(sc.status), _i(sc.iters), _d(sc.obj), _d(sc.pri_res), _d(sc.dua_res), _d(sc.lower))
        hs = (C.c_void_p * B)(*[q._h.value for q in qps])
        _check(lib().bqp_ctx_solve_multi(self._c, B, hs, _ptr_array(ls), _ptr_array(us), _ptr_array(xs), _ptr_array(ys),
                                         _ptr_array([r.x[0] for r in res]), _ptr_array([r.y[0] for r in res]), C.byref(out)))
        return [r.x[0] for r in res], [r.y[0] for r in res], sc

    def last_timing(self):
        t = Timing()
        _check(lib().bqp_ctx_last_timing(self._c, C.byref(t)))
        return t.as_dict()


def _alloc_result(B, n, m):
    r = BatchResult()
    r.x = np.empty((B, n)); r.y = np.empty((B, m))
    r.status = np.empty(B, np.int32); r.iters = np.empty(B, np.int32)
    r.obj = np.empty(B); r.pri_res = np.empty(B); r.dua_res = np.empty(B); r.lower = np.empty(B)
    return r


def _node_out(r):
    return _NodeOut(_i(r.status), _i(r.iters), _d(r.obj), _d(r.pri_res), _d(r.dua_res), _d(r.lower))


class _Scalars(object):
    __slots__ = ("status", "iters", "obj", "pri_res", "dua_res", "lower")


def _ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


class ResidentBatch(object):
    """Staged variant: upload once, run many times from HBM-resident inputs, download once."""

    def __init__(self, qps, l, u, x0, y0):
        self.qps = list(qps)
        B = len(self.qps)
        self._keep = [[_f64(a) for a in seq] for seq in (l, u, x0, y0)]
        hs = (C.c_void_p * B)(*[q._h.value for q in self.qps])
        _check(lib().bqp_batch_upload(B, hs, *[_ptr_array(k) for k in self._keep]))

    def run(self):
        _check(lib().bqp_batch_run())

    def download(self):
        B = len(self.qps)
        xs = [np.empty(q.n) for q in self.qps]
        ys = [np.empty(q.m) for q in self.qps]
        sc = _Scalars()
        sc.status = np.empty(B, np.int32); sc.iters = np.empty(B, np.int32)
        sc.obj = np.empty(B); sc.pri_res = np.empty(B); sc.dua_res = np.empty(B); sc.lower = np.empty(B)
        out = _node_out(sc)
        _check(lib().bqp_batch_download(_ptr_array(xs), _ptr_array(ys), C.byref(out)))
        return xs, ys, sc


def solve_multi(qps, l, u, x0, y0):
    """Nodes of several set-up problems in ONE submission (bqp_solve_multi: upload, launches and read-back behind one C call; a
    kernel that needs a single launch is waited for once).  All arguments are sequences of length B."""
    qps = list(qps)
    B = len(qps)
    keep = [[_f64(a) for a in seq] for seq in (l, u, x0, y0)]
    hs = (C.c_void_p * B)(*[q._h.value for q in qps])
    xs = [np.empty(q.n) for q in qps]
    ys = [np.empty(q.m) for q in qps]
    sc = _Scalars()
    sc.status = np.empty(B, np.int32); sc.iters = np.empty(B, np.int32)
    sc.obj = np.empty(B); sc.pri_res = np.empty(B); sc.dua_res = np.empty(B); sc.lower = np.empty(B)
    out = _node_out(sc)
    _check(lib().bqp_solve_multi(B, hs, *[_ptr_array(k) for k in keep], _ptr_array(xs), _ptr_array(ys), C.byref(out)))
    return xs, ys, sc
