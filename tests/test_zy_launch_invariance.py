"""
GPU regression test of the small-problem race fixed at the end of round 1.  It lives in its own
file, collected after the suites that were already green on a B200, because it was written after the round's last
GPU run (tools/batch_invariance.py, the long form of the same check, did run there: profiles/r01_launch_invariance_*).
"""
import numpy as np
import pytest

from miosqp_b200 import engine, problems
from test_gpu_parity import _close, panel_threads

pytestmark = pytest.mark.gpu


def test_small_problem_launch_invariance(oracle_mod):
    """Regression (round 1): with fewer than 14 row panels in A (here m_ext = 60 -> 8) the update warps of the panel
    kernel prefetched z, y for the first termination-check pass before the previous pass had finished writing them,
    so the dual residual -- and 2 of 111 termination decisions -- depended on timing.  Every node of a real B&B run
    must get the oracle's status and iteration count, alone or in one launch with all the others, bit for bit
    the same iterates either way (tools/batch_invariance.py is the long form of this test)."""
    import miosqp_b200
    pr = problems.random_miqp(40, 40, 20, 0.7, seed=3)[0]
    rec = []
    real = engine.solve_multi

    def spy(qps, l, u, x0, y0):
        rec.extend((np.array(l[k]), np.array(u[k]), np.array(x0[k]), np.array(y0[k])) for k in range(len(qps)))
        return real(qps, l, u, x0, y0)
    engine.solve_multi = spy
    try:
        s = miosqp_b200.MIOSQP()
        s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
                dict(problems.RANDOM_MIQP_SETTINGS), dict(problems.RANDOM_MIQP_QP_SETTINGS))
        s.solve()
    finally:
        engine.solve_multi = real
    assert len(rec) > 100
    L, U, X0, Y0 = (np.array([r[k] for r in rec]) for k in range(4))
    P, q, A, l, u, i_idx = problems.extend(pr)
    o = oracle_mod.OSQP(); o.setup(P, q, A, l, u, **problems.RANDOM_MIQP_QP_SETTINGS)
    xo, yo, so, io, extra = o.solve_batch(L, U, X0, Y0, threads=8)
    qp = s.work.solver
    together = qp.solve_batch(L, U, X0, Y0)
    assert engine.last_timing()["kernel"] == 3      # the row-split cluster kernel (n = 40: 2 column tiles, one CTA per tile)
    again = qp.solve_batch(L, U, X0, Y0)
    assert list(together.status) == list(so) and list(together.iters) == list(io)
    _close(together.pri_res, extra["pri_res"]); _close(together.dua_res, extra["dua_res"])
    for a, b in ((together.x, again.x), (together.y, again.y), (together.dua_res, again.dua_res), (together.pri_res, again.pri_res)):
        assert np.array_equal(a, b, equal_nan=True)
    for k in range(0, len(rec), 7):
        alone = qp.solve_batch(L[k:k + 1], U[k:k + 1], X0[k:k + 1], Y0[k:k + 1])
        assert alone.iters[0] == io[k] and alone.status[0] == so[k]
        assert np.array_equal(alone.x[0], together.x[k], equal_nan=True)
        assert abs(alone.dua_res[0] - together.dua_res[k]) <= 1e-12 * (1 + abs(together.dua_res[k]))
