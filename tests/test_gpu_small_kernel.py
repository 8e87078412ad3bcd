"""
GPU parity of the shared-memory-resident kernel for small sparse problems (miosqp_b200/csrc/bqp_small.cu: npad <= 64, at most
4 entries per row of A and 8 per column, m <= 192 -- BASELINE config 3, the power-converter MPC program) against the CPU oracle
and against the direct-load LDL' kernel it replaces for these shapes.  Same tolerances as tests/test_gpu_parity.py: identical status and
iteration count per node, x / y / objective / residuals / node bound to 1e-9 relative.
"""
import numpy as np
import pytest

from miosqp_b200 import engine, problems
from test_gpu_parity import _compare, _close, QP
from test_abi_and_layout import _mpc_problem

pytestmark = pytest.mark.gpu

SMALL = 5        # bqp_timing.kernel of the shared-memory-resident kernel (include/bqp.h)


def expect_small():
    t = engine.last_timing()
    assert t["kernel"] == SMALL and t["threads"] == 256, t


# (n, m, |i_idx|, density, seed): the last two have a singular sparse P -- every leaf ends dual infeasible (certificate path)
SHAPES = [(60, 130, 60, 0.02, 2), (30, 60, 8, 0.04, 2), (20, 40, 10, 0.08, 1), (60, 130, 60, 0.02, 3), (60, 90, 6, 0.02, 4), (40, 120, 6, 0.03, 3)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("warm", ["zero", "root"])
def test_small_kernel_against_oracle(oracle_mod, shape, warm):
    n, m, p, d, seed = shape
    pr = problems.random_miqp(n, m, p, d, seed=seed)[0]
    _compare(pr, 11, seed, QP, warm=warm, oracle_mod=oracle_mod)        # 11 leaves: one full tile and one of three
    expect_small()


@pytest.mark.parametrize("tt", [1, 2, 4, 8])
def test_small_kernel_tile_widths(oracle_mod, tt):
    """1, 2, 4 or 8 leaves per tile (threads of leaf pairs without a leaf skip the vector phases): same results."""
    pr = problems.random_miqp(60, 130, 60, 0.02, seed=2)[0]
    _compare(pr, 13, 2, QP, warm="root", tuning=(tt, 0), oracle_mod=oracle_mod)
    expect_small()
    assert engine.last_timing()["tile_nodes"] == tt


def test_small_kernel_tight_tolerance_and_max_iter(oracle_mod):
    pr = problems.random_miqp(20, 40, 10, 0.08, seed=1)[0]
    _compare(pr, 6, 1, dict(eps_abs=1e-7, eps_rel=1e-7, eps_prim_inf=1e-6, max_iter=8000), oracle_mod=oracle_mod)
    expect_small()
    r, _ = _compare(pr, 6, 1, dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=60), oracle_mod=oracle_mod)      # stops at the iteration limit
    expect_small()
    assert set(r.status) <= {2, -2}


def test_small_kernel_infeasible_nodes(oracle_mod):
    """Contradictory bounds on a general row give OSQP_PRIMAL_INFEASIBLE with NaN iterates."""
    pr = problems.random_miqp(30, 60, 8, 0.04, seed=2)[0]
    P, q, A, l, u, i_idx = problems.extend(pr)
    o = oracle_mod.OSQP(); o.setup(P, q, A, l, u, **QP)
    e = engine.BatchedQP().setup(P, q, A, l, u, i_idx=i_idx, **QP)
    ls = np.tile(l, (3, 1)); us = np.tile(u, (3, 1))
    ls[1, 0] = 50.0; us[1, 0] = 60.0
    ls[2, 1] = -60.0; us[2, 1] = -50.0
    x0 = np.zeros((3, 30)); y0 = np.zeros((3, A.shape[0]))
    xo, yo, so, io, _ = o.solve_batch(ls, us, x0, y0)
    r = e.solve_batch(ls, us, x0, y0)
    expect_small()
    assert list(so) == [1, -3, -3]
    assert list(r.status) == list(so) and list(r.iters) == list(io)
    assert np.isnan(r.x[1]).all() and np.isnan(r.y[2]).all() and np.isnan(r.lower[1])


def test_small_kernel_mpc_program(oracle_mod):
    """BASELINE config 3: the power-converter MPC program (n = 60, m = 150, 3 entries per row of A), branched leaves."""
    P, q, A, l, u, i_idx = _mpc_problem()
    n, m = 60, A.shape[0]
    rng = np.random.default_rng(3)
    ls, us = problems.branched_nodes(l, u, len(i_idx), 20, rng, depth=6)
    s = dict(eps_abs=1e-3, eps_rel=1e-3, eps_prim_inf=1e-4)
    o = oracle_mod.OSQP(); o.setup(P, q, A, l, u, **s)
    e = engine.BatchedQP().setup(P, q, A, l, u, i_idx=i_idx, **s)
    x0 = np.zeros((20, n)); y0 = np.zeros((20, m))
    xo, yo, so, io, extra = o.solve_batch(ls, us, x0, y0, threads=8)
    r = e.solve_batch(ls, us, x0, y0)
    expect_small()
    assert list(r.status) == list(so) and list(r.iters) == list(io)
    for b in range(20):
        if so[b] in (1, -2):
            xo[b, i_idx] = np.minimum(np.maximum(xo[b, i_idx], ls[b, -len(i_idx):]), us[b, -len(i_idx):])
    _close(r.x, xo); _close(r.y, yo)
    _close(r.pri_res, extra["pri_res"]); _close(r.dua_res, extra["dua_res"])


def test_small_and_direct_load_kernels_agree(monkeypatch):
    """The same leaves through the shared-memory-resident kernel (explicit reduced inverse on the FP64 tensor pipe) and
    through the direct-load kernel (blocked LDL' sweeps): same statuses and iteration counts, iterates to 1e-9."""
    pr = problems.random_miqp(60, 130, 60, 0.02, seed=2)[0]
    P, q, A, l, u, i_idx = problems.extend(pr)
    ls, us = problems.branched_nodes(l, u, len(i_idx), 13, np.random.default_rng(1))
    x0 = np.zeros((13, 60)); y0 = np.zeros((13, A.shape[0]))
    e = engine.BatchedQP().setup(P, q, A, l, u, i_idx=i_idx, **QP)
    r1 = e.solve_batch(ls, us, x0, y0)
    expect_small()
    monkeypatch.setenv("BQP_KERNEL", "direct")
    r0 = e.solve_batch(ls, us, x0, y0)
    assert engine.last_timing()["kernel"] == 0
    assert list(r1.status) == list(r0.status) and list(r1.iters) == list(r0.iters)
    _close(r1.x, r0.x); _close(r1.y, r0.y); _close(r1.lower, r0.lower)


def test_small_kernel_mixed_problems_one_launch(oracle_mod):
    """Problems of different widths (npad 32 and 64) and shapes in ONE launch: a CTA reads its own problem's blob."""
    shapes = [(20, 40, 10, 0.08, 1), (60, 130, 60, 0.02, 2), (30, 60, 8, 0.04, 2)]
    qps, os_, L, U, X0, Y0 = [], [], [], [], [], []
    rng = np.random.default_rng(5)
    for n, m, p, d, seed in shapes:
        P, q, A, l, u, i_idx = problems.extend(problems.random_miqp(n, m, p, d, seed=seed)[0])
        e = engine.BatchedQP().setup(P, q, A, l, u, i_idx=i_idx, **QP)
        o = oracle_mod.OSQP(); o.setup(P, q, A, l, u, **QP)
        ls, us = problems.branched_nodes(l, u, len(i_idx), 9, rng)
        for b in range(9):
            qps.append(e); os_.append(o); L.append(ls[b]); U.append(us[b]); X0.append(np.zeros(n)); Y0.append(np.zeros(A.shape[0]))
    xs, ys, sc = engine.solve_multi(qps, L, U, X0, Y0)
    expect_small()
    assert engine.last_timing()["launches"] == 1
    xo, yo, so, io, _ = oracle_mod.solve_multi(os_, L, U, X0, Y0, threads=8)
    assert list(sc.status) == list(so) and list(sc.iters) == list(io)
    for b in range(len(qps)):
        if so[b] in (1, -2):
            _close(ys[b], yo[b])


def test_small_kernel_results_do_not_depend_on_the_batch(oracle_mod):
    """A leaf's result is a pure function of its inputs: alone, or as any member of a full tile -- bit for bit."""
    pr = problems.random_miqp(30, 60, 8, 0.04, seed=2)[0]
    P, q, A, l, u, i_idx = problems.extend(pr)
    ls, us = problems.branched_nodes(l, u, len(i_idx), 8, np.random.default_rng(2))
    x0 = np.zeros((8, 30)); y0 = np.zeros((8, A.shape[0]))
    e = engine.BatchedQP().setup(P, q, A, l, u, i_idx=i_idx, **QP)
    full = e.solve_batch(ls, us, x0, y0)
    expect_small()
    for b in (0, 3, 7):
        one = e.solve_batch(ls[b:b + 1], us[b:b + 1], x0[b:b + 1], y0[b:b + 1])
        assert one.status[0] == full.status[b] and one.iters[0] == full.iters[b]
        assert np.array_equal(one.x[0], full.x[b]) and np.array_equal(one.y[0], full.y[b]) and one.lower[0] == full.lower[b]


@pytest.mark.parametrize("shape", [(60, 130, 60, 0.02, 2), (30, 60, 8, 0.5, 5), (600, 900, 20, 0.05, 12)])
def test_single_wait_and_stepwise_submission_agree_bit_for_bit(shape):
    """bqp_solve_multi waits once per single-launch batch (shared-memory-resident, direct-load and whole-GPU kernels); the staged
    bqp_batch_upload / run / download calls wait after every step.  Same kernels, same inputs: identical bits."""
    n, m, p, d, seed = shape
    P, q, A, l, u, i_idx = problems.extend(problems.random_miqp(n, m, p, d, seed=seed)[0])
    e = engine.BatchedQP().setup(P, q, A, l, u, i_idx=i_idx, **QP)
    ls, us = problems.branched_nodes(l, u, len(i_idx), 9, np.random.default_rng(4))
    qps = [e] * 9; X0 = [np.zeros(n)] * 9; Y0 = [np.zeros(A.shape[0])] * 9
    xs, ys, sc = engine.solve_multi(qps, list(ls), list(us), X0, Y0)
    kernel = engine.last_timing()["kernel"]
    assert kernel in (0, 4, 5) and engine.last_timing()["launches"] == 1
    rb = engine.ResidentBatch(qps, list(ls), list(us), X0, Y0)
    rb.run()
    xs2, ys2, sc2 = rb.download()
    assert engine.last_timing()["kernel"] == kernel
    assert list(sc.status) == list(sc2.status) and list(sc.iters) == list(sc2.iters)
    for b in range(9):
        assert np.array_equal(xs[b], xs2[b], equal_nan=True) and np.array_equal(ys[b], ys2[b], equal_nan=True)
    assert np.array_equal(sc.lower, sc2.lower, equal_nan=True) and np.array_equal(sc.pri_res, sc2.pri_res, equal_nan=True)


def test_wide_rows_stay_on_the_direct_load_kernel(oracle_mod):
    """More than 4 entries in a row of A (or 8 in a column): no shared-memory layout, the direct-load kernel serves it."""
    pr = problems.random_miqp(30, 60, 8, 0.5, seed=5)[0]
    _compare(pr, 5, 0, QP, oracle_mod=oracle_mod)
    assert engine.last_timing()["kernel"] == 0


def test_small_problems_through_the_bnb_drivers():
    """Branch and bound on small sparse MIQPs (shared-memory-resident kernel) through every driver: one tree at a time (native
    replay, single-wait submissions), lock-step, one stream per tree, rolling session -- and the Python replay: same decisions,
    node and ADMM iteration counts, incumbents to 1e-9."""
    import miosqp_b200
    cases = [(30, 60, 8, 0.04, 2), (20, 40, 10, 0.08, 1), (60, 130, 60, 0.02, 2)]

    def trees(replay):
        out = []
        for n, m, p, d, seed in cases:
            pr = problems.random_miqp(n, m, p, d, seed=seed)[0]
            s = miosqp_b200.MIOSQP()
            s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
                    dict(problems.RANDOM_MIQP_SETTINGS, replay=replay, speculation=4), dict(problems.RANDOM_MIQP_QP_SETTINGS))
            out.append(s)
        return out

    def sig(res, solvers):
        return [(r.status, s.work.iter_num, int(s.work.osqp_iter), [tuple(d) for d in s.work.decisions]) for r, s in zip(res, solvers)]

    runs = {}
    for driver in ("single", "python", "lockstep", "async", "rolling"):
        solvers = trees(None if driver == "python" else 'native')
        if driver in ("single", "python"):
            res = [s.solve() for s in solvers]
            expect_small()
        else:
            res = miosqp_b200.solve_many(solvers, async_threads=(0 if driver == "async" else None), rolling=(driver == "rolling"))
        runs[driver] = (sig(res, solvers), [float(r.upper_glob) for r in res], [np.array(r.x) for r in res])
        for s in solvers:
            s.work.solver.free()
    ref = runs["single"]
    assert all(it > 1 for _, it, _, _ in ref[0])                       # every tree branched at least once
    for driver, run in runs.items():
        assert run[0] == ref[0], driver
        for k in range(len(cases)):
            assert abs(run[1][k] - ref[1][k]) <= 1e-9 * (1 + abs(ref[1][k])), driver
            assert np.abs(run[2][k] - ref[2][k]).max() <= 1e-9 * (1 + np.abs(ref[2][k]).max()), driver
