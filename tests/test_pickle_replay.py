"""
BASELINE config 5: the reference's 49 hard QP relaxations (/root/reference/max_iter_examples/*.pickle, inputs
only -- the reference holds no expected outputs for them), replayed as ONE batch of 49 different problems and
compared with the CPU oracle: same status and iteration count, iterates and residuals to 1e-9 relative.
The fixture tests/golden/max_iter_examples.npz is produced by tests/golden/make_pickle_fixture.py.
"""
import os

import numpy as np
import pytest
import scipy.sparse as spa

from miosqp_b200 import engine, maxiter_problems

HERE = os.path.dirname(os.path.abspath(__file__))


def load_problems():
    return maxiter_problems.load_npz(os.path.join(HERE, "golden", "max_iter_examples.npz"))


def test_fixture_shape_and_settings():
    probs = load_problems()
    assert [p["name"] for p in probs] == list(range(28, 77))
    for p in probs:
        assert p["P"].shape == (20, 20) and p["A"].shape == (60, 20) and len(p["i_idx"]) == 10
        assert np.abs(p["P"].toarray() - p["P"].toarray().T).max() == 0.0          # stored as full symmetric P
        s = engine.normalize_settings(p["settings"])                               # legacy names are understood
        assert s["eps_prim_inf"] == p["settings"]["eps_inf"] and s["sigma"] == 0.01
    assert probs[-1]["settings"]["max_iter"] == 5000 and probs[0]["settings"]["max_iter"] == 2500
    groups = {(p["P"].toarray().tobytes(), p["A"].toarray().tobytes(), p["q"].tobytes()) for p in probs}
    assert len(groups) == 7


def test_oracle_runs_all(oracle_mod):
    """The oracle terminates on every fixture with a legal OSQP status; most hit max_iter or prove infeasibility."""
    st = []
    for p in load_problems():
        o = oracle_mod.OSQP(); o.setup(p["P"], p["q"], p["A"], p["l"], p["u"], **p["settings"])
        r = o.solve()
        st.append(r.info.status_val)
        assert r.info.iter <= p["settings"]["max_iter"]
    assert set(st) <= {1, 2, 3, 4, -2, -3, -4}


@pytest.mark.gpu
def test_replay_49_in_one_launch(oracle_mod):
    probs = load_problems()
    qps, os_, L, U, X0, Y0 = [], [], [], [], [], []
    for p in probs:
        e = engine.BatchedQP().setup(p["P"], p["q"], p["A"], p["l"], p["u"], **p["settings"])
        o = oracle_mod.OSQP(); o.setup(p["P"], p["q"], p["A"], p["l"], p["u"], **p["settings"])
        qps.append(e); os_.append(o); L.append(p["l"]); U.append(p["u"]); X0.append(np.zeros(20)); Y0.append(np.zeros(60))
    xs, ys, sc = engine.solve_multi(qps, L, U, X0, Y0)
    assert engine.last_timing()["launches"] == 1 and engine.last_timing()["tiles"] == 49
    xo, yo, so, io, _ = oracle_mod.solve_multi(os_, L, U, X0, Y0, threads=8)
    assert list(sc.status) == list(so)
    assert list(sc.iters) == list(io)
    for b in range(len(probs)):
        if so[b] in (1, 2, -2):
            assert np.abs(xs[b] - xo[b]).max() <= 1e-9 * (1 + np.abs(xo[b]).max())
            assert np.abs(ys[b] - yo[b]).max() <= 1e-9 * (1 + np.abs(yo[b]).max())
        else:
            assert np.isnan(xs[b]).all() and np.isnan(ys[b]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["grid", "stream"])
def test_cfg4_shape_fixed_iterations(oracle_mod, kernel, monkeypatch):
    """BASELINE config 4 shape (n=2000, m=4000, |i_idx|=200, 5 % dense) on the whole-GPU kernel (default) and, with BQP_GRID=0
    at setup, on the streamed kernel with sparse A groups and a 64-block dense tail; 2 leaves, iterates compared after a
    fixed 100 iterations (status max_iter)."""
    from miosqp_b200 import problems
    if kernel == "stream":
        monkeypatch.setenv("BQP_GRID", "0")
    pr = problems.random_miqp(2000, 4000, 200, 0.05, seed=1)[0]
    P, q, A, l, u, i_idx = problems.extend(pr)
    s = dict(eps_abs=1e-9, eps_rel=1e-9, eps_prim_inf=1e-9, eps_dual_inf=1e-9, max_iter=100)
    o = oracle_mod.OSQP(); o.setup(P, q, A, l, u, **s)
    e = engine.BatchedQP().setup(P, q, A, l, u, i_idx=i_idx, **s)
    ls, us = problems.branched_nodes(l, u, len(i_idx), 2, np.random.default_rng(0))
    x0 = np.zeros((2, 2000)); y0 = np.zeros((2, 4200))
    xo, yo, so, io, extra = o.solve_batch(ls, us, x0, y0, threads=2)
    r = e.solve_batch(ls, us, x0, y0)
    t = engine.last_timing()
    assert (t["kernel"], t["threads"]) == ((4, 512) if kernel == "grid" else (1, 13 * 32)), t
    assert list(r.status) == list(so) and list(r.iters) == list(io) == [100, 100]
    for b in range(2):
        xo[b, i_idx] = np.minimum(np.maximum(xo[b, i_idx], ls[b, -200:]), us[b, -200:])
    assert np.abs(r.x - xo).max() <= 1e-9 * (1 + np.abs(xo).max())
    assert np.abs(r.y - yo).max() <= 1e-9 * (1 + np.abs(yo).max())
    assert np.abs(r.pri_res - extra["pri_res"]).max() <= 1e-9 * (1 + extra["pri_res"].max())
