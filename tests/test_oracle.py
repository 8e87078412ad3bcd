"""
The CPU oracle (oracle/osqp_oracle.c) against reference-independent known answers.  The reference's own
arithmetic lives in PyPI `osqp`, which is not installable here and ships no golden vectors in the reference
tree (PARITY UNPINNED, see oracle header), so the oracle is pinned by: closed-form QPs, KKT optimality of its
output, LDL' reconstruction, OSQP's documented status semantics, and the exhaustive-enumeration optimum of
BASELINE config 1.
"""
import numpy as np
import pytest
import scipy.sparse as spa

from miosqp_b200 import problems
from oracle import oracle

QP = dict(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000)


def _solve(P, q, A, l, u, **kw):
    o = oracle.OSQP()
    s = dict(QP); s.update(kw)
    o.setup(spa.csc_matrix(P), np.asarray(q, float), spa.csc_matrix(A), np.asarray(l, float), np.asarray(u, float), **s)
    return o, o.solve()


def test_box_qp_closed_form():
    # min 1/2 x'x - c'x  s.t. 0 <= x <= 1  ->  x = clip(c, 0, 1)
    c = np.array([-0.5, 0.3, 1.7, 0.999])
    o, r = _solve(np.eye(4), -c, np.eye(4), np.zeros(4), np.ones(4))
    assert r.info.status_val == 1
    assert np.abs(r.x - np.clip(c, 0, 1)).max() < 1e-6
    # multipliers: y = c - x on active bounds (sign: positive at upper bound)
    assert np.abs(r.y - (c - np.clip(c, 0, 1))).max() < 1e-5


def test_equality_constrained_closed_form():
    # min 1/2 x'Px + q'x s.t. a'x = b  ->  KKT system
    rng = np.random.default_rng(0)
    M = rng.standard_normal((5, 5)); P = M @ M.T + np.eye(5); q = rng.standard_normal(5)
    a = rng.standard_normal((1, 5)); b = np.array([0.7])
    K = np.block([[P, a.T], [a, np.zeros((1, 1))]])
    sol = np.linalg.solve(K, np.r_[-q, b])
    o, r = _solve(P, q, a, b, b)
    assert r.info.status_val == 1
    assert np.abs(r.x - sol[:5]).max() < 1e-6 and abs(r.y[0] - sol[5]) < 1e-5


def test_primal_infeasible_and_dual_infeasible_status():
    # x <= -1 and x >= 1 cannot hold together
    A = np.array([[1.0], [1.0]])
    o, r = _solve(np.eye(1), [0.0], A, [-np.inf, 1.0], [-1.0, np.inf], eps_abs=1e-4, eps_rel=1e-4)
    assert r.info.status_val == oracle.constant("OSQP_PRIMAL_INFEASIBLE")
    assert np.isnan(r.x).all()
    # unbounded below along x2 (P singular there, q2 < 0, no upper bound)
    P = np.diag([1.0, 0.0]); A = np.eye(2)
    o, r = _solve(P, [0.0, -1.0], A, [-1.0, 0.0], [1.0, np.inf], eps_abs=1e-4, eps_rel=1e-4)
    assert r.info.status_val == oracle.constant("OSQP_DUAL_INFEASIBLE")


def test_max_iter_status_and_iteration_granularity():
    pr = problems.random_miqp(50, 100, 5, 0.7, seed=1)[0]
    P, q, A, l, u, _ = problems.extend(pr)
    o = oracle.OSQP(); o.setup(P, q, A, l, u, eps_abs=1e-9, eps_rel=1e-9, max_iter=60)
    r = o.solve()
    assert r.info.status_val in (oracle.constant("OSQP_MAX_ITER_REACHED"), oracle.constant("OSQP_SOLVED_INACCURATE"))
    assert r.info.iter == 60
    o2 = oracle.OSQP(); o2.setup(P, q, A, l, u, eps_abs=1e-3, eps_rel=1e-3)
    r2 = o2.solve()
    assert r2.info.status_val == 1 and r2.info.iter % 25 == 0      # termination is only tested every 25 iterations


def test_kkt_optimality_random_qp():
    pr = problems.random_miqp(50, 100, 5, 0.7, seed=1)[0]
    P, q, A, l, u, _ = problems.extend(pr)
    o = oracle.OSQP(); o.setup(P, q, A, l, u, **QP)
    r = o.solve()
    assert r.info.status_val == 1
    Pd = P.toarray(); Ad = A.toarray()
    assert np.abs(Pd @ r.x + q + Ad.T @ r.y).max() < 1e-6                 # stationarity
    z = Ad @ r.x
    assert (z >= l - 1e-6).all() and (z <= u + 1e-6).all()               # primal feasibility
    assert (r.y[(z > l + 1e-5) & (z < u - 1e-5)] ** 2).max() < 1e-10     # complementarity
    assert (r.y[np.abs(z - u) < 1e-7] >= -1e-8).all() and (r.y[np.abs(z - l) < 1e-7] <= 1e-8).all()


def test_cfg1_fingerprints_and_root_relaxation():
    """BASELINE.md known answer: root relaxation objective -12.706728044."""
    pr = problems.random_miqp(50, 100, 5, 0.7, seed=1)[0]
    assert list(pr["i_idx"]) == [27, 35, 40, 38, 2]
    assert abs(pr["P"].data.sum() - 15878.928065531603) < 1e-6 and abs(pr["q"].sum() - 1.245233404579) < 1e-9
    P, q, A, l, u, _ = problems.extend(pr)
    o = oracle.OSQP(); o.setup(P, q, A, l, u, **QP)
    r = o.solve()
    assert abs(r.info.obj_val - (-12.706728044)) < 1e-6
    assert np.abs(r.x[pr["i_idx"]] - np.array([0, 0, 0, 0.392675, 0.013816])).max() < 1e-4


def test_cfg1_exhaustive_enumeration_optimum():
    """All 2^5 binary assignments solved as QPs: optimum -12.420811293 at all-zero."""
    pr = problems.random_miqp(50, 100, 5, 0.7, seed=1)[0]
    P, q, A, l, u, i_idx = problems.extend(pr)
    o = oracle.OSQP(); o.setup(P, q, A, l, u, **QP)
    L, U = [], []
    for code in range(32):
        ll, uu = l.copy(), u.copy()
        for k in range(5):
            ll[100 + k] = uu[100 + k] = (code >> k) & 1
        L.append(ll); U.append(uu)
    x, y, st, it, extra = o.solve_batch(np.array(L), np.array(U), np.zeros((32, 50)), np.zeros((32, 105)), threads=8)
    assert (st == 1).all()
    objs = extra["obj"]
    assert int(np.argmin(objs)) == 0 and abs(objs.min() - (-12.420811293)) < 1e-6
    assert abs(np.sort(objs)[1] - (-12.035724881)) < 1e-6


def test_ldl_reconstruction_and_scaling():
    pr = problems.random_miqp(30, 60, 4, 0.5, seed=2)[0]
    P, q, A, l, u, _ = problems.extend(pr)
    o = oracle.OSQP(); o.setup(P, q, A, l, u, eps_abs=1e-3, eps_rel=1e-3, sigma=1e-6, rho=0.1)
    n, m, N, nnzL = o.dims()
    perm, Lp, Li, Lx, Dd = o.factor()
    L = spa.csc_matrix((Lx, Li, Lp), shape=(N, N)).toarray() + np.eye(N)
    K = L @ np.diag(Dd) @ L.T
    b = np.random.default_rng(1).standard_normal(N)
    sol = o.kkt_solve(b)
    Kfull = np.zeros((N, N)); Kfull[np.ix_(perm, perm)] = K
    assert np.abs(Kfull @ sol - b).max() < 1e-8 * (1 + np.abs(b).max()) * 1e3
    D, E, c = o.scaling()
    assert (D > 0).all() and (E > 0).all() and c > 0
    assert (Dd[:0] if False else True)
    # quasi-definite: n positive and m negative pivots
    assert (Dd > 0).sum() == n and (Dd < 0).sum() == m


def test_pure_function_of_node_inputs():
    """Parity contract: a node's result depends on (l,u,x0,y0) only -- not on what was solved before."""
    pr = problems.random_miqp(50, 100, 5, 0.7, seed=1)[0]
    P, q, A, l, u, i_idx = problems.extend(pr)
    o = oracle.OSQP(); o.setup(P, q, A, l, u, eps_abs=1e-3, eps_rel=1e-3)
    ls, us = problems.branched_nodes(l, u, 5, 4, np.random.default_rng(0))
    a = [o.solve_node(ls[k], us[k], np.zeros(50), np.zeros(105)) for k in range(4)]
    b = [o.solve_node(ls[k], us[k], np.zeros(50), np.zeros(105)) for k in (3, 2, 1, 0)][::-1]
    for ra, rb in zip(a, b):
        assert ra.info.iter == rb.info.iter and np.array_equal(ra.x, rb.x)


def test_bounds_validation():
    pr = problems.random_miqp(20, 30, 3, 0.5, seed=3)[0]
    P, q, A, l, u, _ = problems.extend(pr)
    o = oracle.OSQP(); o.setup(P, q, A, l, u)
    bad = l.copy(); bad[0] = u[0] + 1
    with pytest.raises(ValueError):
        o.update(l=bad, u=u)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_against_an_independent_solver(seed):
    """Cross-solver check: the oracle at tight tolerance against scipy's SLSQP (a sequential-QP active-set method that
    shares nothing with ADMM) on random strictly convex QPs with two-sided, one-sided and equality rows."""
    from scipy.optimize import minimize
    rng = np.random.default_rng(seed)
    n, m = 8, 12
    M = rng.standard_normal((n, n)); P = M @ M.T + 0.1 * np.eye(n); q = rng.standard_normal(n)
    A = rng.standard_normal((m, n))
    xf = rng.standard_normal(n); z = A @ xf                       # a feasible point keeps the problem feasible
    l = z - rng.uniform(0.1, 1.0, m); u = z + rng.uniform(0.1, 1.0, m)
    l[:3] = -np.inf; u[3:5] = np.inf; l[5] = u[5] = z[5]
    o, r = _solve(P, q, A, l, u)
    assert r.info.status_val == 1
    cons = []
    for i in range(m):
        if l[i] == u[i]:
            cons.append({"type": "eq", "fun": lambda x, i=i: A[i] @ x - u[i], "jac": lambda x, i=i: A[i]})
            continue
        if np.isfinite(u[i]):
            cons.append({"type": "ineq", "fun": lambda x, i=i: u[i] - A[i] @ x, "jac": lambda x, i=i: -A[i]})
        if np.isfinite(l[i]):
            cons.append({"type": "ineq", "fun": lambda x, i=i: A[i] @ x - l[i], "jac": lambda x, i=i: A[i]})
    ref = minimize(lambda x: 0.5 * x @ P @ x + q @ x, xf, jac=lambda x: P @ x + q, constraints=cons, method="SLSQP",
                   options={"ftol": 1e-14, "maxiter": 500})
    assert ref.success
    f_or = 0.5 * r.x @ P @ r.x + q @ r.x
    assert abs(f_or - ref.fun) <= 1e-7 * (1 + abs(ref.fun))
    assert np.abs(r.x - ref.x).max() <= 1e-5 * (1 + np.abs(ref.x).max())
    # dual sign convention (OSQP): y_i > 0 only at the upper bound, y_i < 0 only at the lower bound
    Ax = A @ r.x
    assert np.all(r.y[Ax < u - 1e-5] <= 1e-6) and np.all(r.y[Ax > l + 1e-5] >= -1e-6)


def test_miqp_optimum_by_enumeration_small():
    """A second reference-independent MIQP answer (besides BASELINE config 1): every 0/1 assignment of 6 integer
    variables solved as a QP by SLSQP; the B&B on the oracle must return that assignment and objective (1e-3, the
    example tolerances)."""
    import itertools
    from scipy.optimize import minimize
    import fake_engine
    import miosqp_b200
    from miosqp_b200 import engine
    pr = problems.random_miqp(10, 14, 6, 0.7, seed=11)[0]
    P = pr['P'].toarray(); A = pr['A'].toarray(); q, l, u, idx = pr['q'], pr['l'], pr['u'], pr['i_idx']
    free = np.setdiff1d(np.arange(10), idx)
    best = (np.inf, None)
    for bits in itertools.product((0.0, 1.0), repeat=6):
        xb = np.zeros(10); xb[idx] = bits
        # reduced QP in the 4 continuous variables
        Pf = P[np.ix_(free, free)]; qf = q[free] + P[np.ix_(free, idx)] @ np.array(bits); Af = A[:, free]; off = A[:, idx] @ np.array(bits)
        cons = [{"type": "ineq", "fun": lambda x: u - off - Af @ x, "jac": lambda x: -Af},
                {"type": "ineq", "fun": lambda x: Af @ x + off - l, "jac": lambda x: Af}]
        ref = minimize(lambda x: 0.5 * x @ Pf @ x + qf @ x, np.zeros(4), jac=lambda x: Pf @ x + qf, constraints=cons, method="SLSQP",
                       options={"ftol": 1e-13, "maxiter": 300})
        if not ref.success or np.any(Af @ ref.x + off > u + 1e-7) or np.any(Af @ ref.x + off < l - 1e-7):
            continue
        xb[free] = ref.x
        f = 0.5 * xb @ P @ xb + q @ xb
        if f < best[0]:
            best = (f, np.array(bits))
    assert best[1] is not None
    saved = engine.BatchedQP, engine.solve_multi
    engine.BatchedQP, engine.solve_multi = fake_engine.FakeBatchedQP, fake_engine.solve_multi
    try:
        s = miosqp_b200.MIOSQP()
        s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
                dict(problems.RANDOM_MIQP_SETTINGS), dict(problems.RANDOM_MIQP_QP_SETTINGS))
        r = s.solve()
    finally:
        engine.BatchedQP, engine.solve_multi = saved
    assert r.status == 'Solved'
    assert np.array_equal(r.x[idx], best[1])
    assert abs(r.upper_glob - best[0]) <= 5e-3 * (1 + abs(best[0]))


def test_per_node_rho_retyping_option():
    """eq_rho = 2 (test infrastructure for SURVEY section 8f2): what osqp >= 0.4 does in update_bounds -- rows whose bounds
    became equalities in a B&B child get rho x 1e3 and the KKT matrix is refactored.  At the root it is the contract
    (eq_rho = 1) bit for bit; in a child with fixed binaries it reaches the same optimum, normally in fewer iterations."""
    pr = problems.random_miqp(30, 60, 10, 0.7, seed=1)[0]
    P, q, A, l, u, i_idx = problems.extend(pr)
    s = dict(eps_abs=1e-6, eps_rel=1e-6, max_iter=20000)
    o1 = oracle.OSQP(); o1.setup(P, q, A, l, u, eq_rho=1, **s)
    o2 = oracle.OSQP(); o2.setup(P, q, A, l, u, eq_rho=2, **s)
    x0 = np.zeros(30); y0 = np.zeros(70)
    r1 = o1.solve_node(l, u, x0, y0); r2 = o2.solve_node(l, u, x0, y0)
    assert r1.info.iter == r2.info.iter and np.array_equal(r1.x, r2.x) and np.array_equal(r1.y, r2.y)
    lc, uc = l.copy(), u.copy()
    uc[-10:-5] = 0.0; lc[-5:-2] = 1.0                  # five binaries fixed to 0, three to 1: eight equality rows
    c1 = o1.solve_node(lc, uc, x0, y0); c2 = o2.solve_node(lc, uc, x0, y0)
    assert c1.info.status_val == c2.info.status_val == 1
    assert np.abs(c1.x - c2.x).max() <= 1e-4 * (1 + np.abs(c1.x).max())
    assert np.abs(c2.x[i_idx[:5]]).max() <= 1e-5 and np.abs(c2.x[i_idx[5:8]] - 1).max() <= 1e-5
    assert c2.info.iter != c1.info.iter                 # a different rho vector is a different ADMM trajectory
    # stateless: the retyped factor is private to the call
    r1b = o2.solve_node(l, u, x0, y0)
    assert r1b.info.iter == r1.info.iter and np.array_equal(r1b.x, r1.x)


# ---------------------------------------------------------------- adaptive rho with a fixed interval
def test_adaptive_rho_reaches_the_same_optimum_in_fewer_iterations():
    """osqp's adapt_rho restated (compute_rho_estimate on the scaled residuals, update outside [rho / 5, 5 rho], numeric
    refactorisation).  No reference-held vector exists for it (parity unpinned, as for the rest of the oracle): pinned by the
    property that matters -- same optimum to the solver tolerance, and far fewer iterations on a problem whose default rho is
    badly tuned (random_miqp: rho settles near 0.013 instead of 0.1)."""
    from miosqp_b200 import problems
    from oracle import oracle
    P, q, A, l, u, i_idx = problems.extend(problems.random_miqp(130, 200, 10, 0.7, seed=4)[0])
    n, m = A.shape[1], A.shape[0]
    qp = dict(eps_abs=1e-4, eps_rel=1e-4)
    fixed = oracle.OSQP(); fixed.setup(P, q, A, l, u, **qp)
    adapt = oracle.OSQP(); adapt.setup(P, q, A, l, u, adaptive_rho=True, adaptive_rho_interval=25, **qp)
    ls, us = problems.branched_nodes(l, u, len(i_idx), 5, np.random.default_rng(1))
    z = np.zeros((5, n)); zy = np.zeros((5, m))
    x1, y1, s1, i1, e1 = fixed.solve_batch(ls, us, z, zy)
    x2, y2, s2, i2, e2 = adapt.solve_batch(ls, us, z, zy)
    assert list(s1) == list(s2) == [1] * 5
    assert (e2["rho_updates"] >= 1).all() and (e1["rho_updates"] == 0).all()
    assert i2.sum() * 2 < i1.sum()
    assert np.abs(e1["obj"] - e2["obj"]).max() <= 5e-3 * (1 + np.abs(e1["obj"]).max())
    assert np.abs(x1 - x2).max() <= 2e-2
    # a node is a pure function of its inputs: the adapted rho does not leak into the next solve
    x3, y3, s3, i3, e3 = adapt.solve_batch(ls[::-1].copy(), us[::-1].copy(), z, zy)
    assert np.array_equal(x3[::-1], x2) and list(i3[::-1]) == list(i2)


def test_adaptive_rho_needs_a_fixed_interval():
    from oracle import oracle
    with pytest.raises(ValueError):
        oracle.normalize_settings({"adaptive_rho": True})
    assert oracle.normalize_settings({"adaptive_rho": True, "adaptive_rho_interval": 50})["adaptive_rho"] == 1
