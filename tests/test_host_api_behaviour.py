"""
Edge behaviour of the host mirror of the reference API (miosqp_b200/miqp.py, tree.py, problem_data.py), stated
against the reference lines it follows and -- where /root/reference is mounted -- checked DIFFERENTIALLY against
the unmodified reference package running on the same CPU oracle (tests/osqp_shim): statuses, node counts, ADMM
iteration totals must be identical; bounds and solutions agree to 1e-12 (the oracle's one-node and batch entry
points unscale in a different order, a last-digit effect).
"""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as spa

from miosqp_b200 import problems

HERE = os.path.dirname(os.path.abspath(__file__))
HAVE_REF = os.path.isdir("/root/reference/miosqp")


@pytest.fixture
def cpu_engine(monkeypatch):
    import fake_engine
    from miosqp_b200 import engine
    monkeypatch.setattr(engine, "BatchedQP", fake_engine.FakeBatchedQP)
    monkeypatch.setattr(engine, "solve_multi", fake_engine.solve_multi)


def _reference():
    sys.path.insert(0, os.path.join(HERE, "osqp_shim"))
    if "/root/reference" not in sys.path:
        sys.path.insert(1, "/root/reference")
    import osqp
    osqp.set_backend("oracle")
    import miosqp
    return miosqp


def _both(pr, settings, qp_settings, x0=None, i_l="given", i_u="given"):
    """Run ours and (if mounted) the reference on one MIQP; returns [(results, workspace), ...]."""
    import miosqp_b200
    out = []
    mods = [miosqp_b200] + ([_reference()] if HAVE_REF else [])
    for mod in mods:
        s = mod.MIOSQP()
        s.setup(pr['P'], np.copy(pr['q']), pr['A'], np.copy(pr['l']), np.copy(pr['u']), pr['i_idx'],
                pr['i_l'] if i_l == "given" else i_l, pr['i_u'] if i_u == "given" else i_u, dict(settings), dict(qp_settings))
        if x0 is not None:
            s.set_x0(np.copy(x0))
        out.append((s.solve(), s.work))
    return out


def _near(a, b):
    return (np.isinf(a) and np.isinf(b) and a == b) or abs(a - b) <= 1e-12 * (1 + abs(b))


def _same(runs):
    (r0, w0) = runs[0]
    for r, w in runs[1:]:
        assert r.status == r0.status and w.iter_num == w0.iter_num and w.osqp_iter == w0.osqp_iter
        assert r.osqp_iter_avg == r0.osqp_iter_avg
        assert _near(r.upper_glob, r0.upper_glob) and _near(w.lower_glob, w0.lower_glob)
        if r0.status in ('Solved', 'Max-iter feasible'):
            assert np.abs(r.x - r0.x).max() <= 1e-12 * (1 + np.abs(r0.x).max())
            assert np.array_equal(r.x[w0.data.i_idx], r0.x[w0.data.i_idx])


def test_status_strings_byte_identical():
    import miosqp_b200 as m
    assert (m.MI_UNSOLVED, m.MI_SOLVED, m.MI_PRIMAL_INFEASIBLE, m.MI_DUAL_INFEASIBLE, m.MI_MAX_ITER_FEASIBLE, m.MI_MAX_ITER_UNSOLVED) == \
        ('Unolved', 'Solved', 'Primal Infeasible', 'Dual Infeasible', 'Max-iter feasible', 'Max-iter unsolved')   # constants.py:2-7
    if HAVE_REF:
        ref = _reference()
        for name in ("MI_UNSOLVED", "MI_SOLVED", "MI_PRIMAL_INFEASIBLE", "MI_DUAL_INFEASIBLE", "MI_MAX_ITER_FEASIBLE", "MI_MAX_ITER_UNSOLVED"):
            assert getattr(ref, name) == getattr(m, name)


@pytest.mark.parametrize("rule", [0, 1])
def test_solved_instance_identical_to_reference(cpu_engine, rule):
    pr = problems.random_miqp(30, 60, 10, 0.7, seed=1)[0]
    runs = _both(pr, dict(problems.RANDOM_MIQP_SETTINGS, tree_explor_rule=rule), problems.RANDOM_MIQP_QP_SETTINGS)
    assert runs[0][0].status == 'Solved'
    _same(runs)


def test_node_limit_statuses(cpu_engine):
    """max_iter_bb reached with / without an incumbent (workspace.py:352-373)."""
    pr = problems.random_miqp(40, 40, 20, 0.7, seed=3)[0]
    runs = _both(pr, dict(problems.RANDOM_MIQP_SETTINGS, max_iter_bb=30), problems.RANDOM_MIQP_QP_SETTINGS)
    assert runs[0][0].status == 'Max-iter feasible' and runs[0][1].iter_num == 30
    _same(runs)
    runs = _both(pr, dict(problems.RANDOM_MIQP_SETTINGS, max_iter_bb=2), problems.RANDOM_MIQP_QP_SETTINGS)
    assert runs[0][0].status in ('Max-iter unsolved', 'Max-iter feasible')
    _same(runs)


def test_infeasible_miqp(cpu_engine):
    """Contradictory rows: the root relaxation is primal infeasible, the tree empties, status 'Primal Infeasible'."""
    pr = problems.random_miqp(20, 30, 4, 0.7, seed=5)[0]
    pr['l'] = pr['l'].copy(); pr['u'] = pr['u'].copy()
    pr['A'] = spa.vstack([pr['A'], pr['A'][0]]).tocsc()
    pr['l'] = np.append(pr['l'], 100.0); pr['u'] = np.append(pr['u'], 200.0)     # row 0 again, but in [100, 200]
    runs = _both(pr, problems.RANDOM_MIQP_SETTINGS, problems.RANDOM_MIQP_QP_SETTINGS)
    assert runs[0][0].status == 'Primal Infeasible' and np.isinf(runs[0][0].upper_glob)
    _same(runs)


def test_default_integer_bounds_are_infinite(cpu_engine):
    """i_l / i_u = None -> -inf / +inf (solver.py:50-53): general integers, branching still terminates."""
    pr = problems.random_miqp(12, 20, 3, 0.7, seed=7)[0]
    runs = _both(pr, problems.RANDOM_MIQP_SETTINGS, problems.RANDOM_MIQP_QP_SETTINGS, i_l=None, i_u=None)
    assert runs[0][0].status == 'Solved'
    x = runs[0][0].x[pr['i_idx']]
    assert np.array_equal(x, np.round(x))
    _same(runs)


def test_set_x0_valid_and_invalid(cpu_engine, capsys):
    """A feasible integral x0 seeds the incumbent; anything else prints the reference's message and is ignored
    (workspace.py:94-111)."""
    pr = problems.random_miqp(30, 60, 10, 0.7, seed=1)[0]
    base = _both(pr, problems.RANDOM_MIQP_SETTINGS, problems.RANDOM_MIQP_QP_SETTINGS)
    good = np.copy(base[0][0].x)
    runs = _both(pr, problems.RANDOM_MIQP_SETTINGS, problems.RANDOM_MIQP_QP_SETTINGS, x0=good)
    assert runs[0][0].status == 'Solved' and runs[0][1].iter_num <= base[0][1].iter_num
    _same(runs)
    capsys.readouterr()
    bad = np.full(30, 0.5)
    runs = _both(pr, problems.RANDOM_MIQP_SETTINGS, problems.RANDOM_MIQP_QP_SETTINGS, x0=bad)
    assert capsys.readouterr().out.count('Invalid initial solution!') == len(runs)
    _same(runs)
    assert runs[0][1].iter_num == base[0][1].iter_num


def test_bad_rules_raise_value_error(cpu_engine):
    import miosqp_b200
    pr = problems.random_miqp(30, 60, 10, 0.7, seed=1)[0]       # its root relaxation is fractional, so it must branch
    for key, msg in (("tree_explor_rule", 'Tree exploring strategy not recognized'), ("branching_rule", 'No variable selection rule recognized!')):
        s = miosqp_b200.MIOSQP()
        s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
                dict(problems.RANDOM_MIQP_SETTINGS, **{key: 7}), dict(problems.RANDOM_MIQP_QP_SETTINGS))
        with pytest.raises(ValueError, match=msg):         # workspace.py:147, 224
            s.solve()


def test_update_vectors_errors_and_reuse(cpu_engine):
    """Dimension errors with the reference's messages (data.py:112-125); a re-solve after update_vectors equals a
    fresh setup on the new vectors (the factor is reused, solver.py:174-205) and reports run_time without setup."""
    import miosqp_b200
    pr = problems.random_miqp(30, 60, 10, 0.7, seed=1)[0]
    s = miosqp_b200.MIOSQP()
    s.setup(pr['P'], pr['q'], pr['A'], np.copy(pr['l']), np.copy(pr['u']), pr['i_idx'], pr['i_l'], pr['i_u'],
            dict(problems.RANDOM_MIQP_SETTINGS), dict(problems.RANDOM_MIQP_QP_SETTINGS))
    r1 = s.solve()
    for kw, msg in ((dict(q=np.zeros(3)), 'Wrong q dimension!'), (dict(l=np.zeros(3)), 'Wrong l dimension!'), (dict(u=np.zeros(3)), 'Wrong u dimension!')):
        with pytest.raises(ValueError, match=msg):
            s.update_vectors(**kw)
    q2 = pr['q'] + 0.1
    s.update_vectors(q=q2, l=pr['l'] - 0.05, u=pr['u'] + 0.05)
    r2 = s.solve()
    fresh = miosqp_b200.MIOSQP()
    fresh.setup(pr['P'], q2, pr['A'], pr['l'] - 0.05, pr['u'] + 0.05, pr['i_idx'], pr['i_l'], pr['i_u'],
                dict(problems.RANDOM_MIQP_SETTINGS), dict(problems.RANDOM_MIQP_QP_SETTINGS))
    r3 = fresh.solve()
    assert r2.status == r3.status == 'Solved' and s.work.decisions == fresh.work.decisions
    # the cost scaling c stays the one computed from the FIRST q (OSQP update_lin_cost rescales, never re-equilibrates), so the
    # ADMM trajectory differs from a fresh setup within the solver tolerance eps_abs = eps_rel = 1e-3
    assert abs(r2.upper_glob - r3.upper_glob) <= 2e-3 * (1 + abs(r3.upper_glob))
    assert r1.run_time >= s.work.setup_time and r2.run_time == s.work.solve_time        # solver.py:155-164


def test_qp_settings_contract():
    from miosqp_b200 import engine
    with pytest.raises(ValueError):
        engine.normalize_settings({"adaptive_rho": True})        # outside the parity contract, never silently ignored
    with pytest.raises(TypeError):
        engine.normalize_settings({"no_such_setting": 1})
    assert engine.normalize_settings({"eq_rho": 2})["eq_rho"] == 2     # per-node re-typing: Woodbury path of the rows kernel
    with pytest.raises(ValueError):
        engine.normalize_settings({"eq_rho": 3})
    s = engine.normalize_settings({"eps_inf": 1e-5, "eps_unb": 1e-6, "polishing": False, "verbose": True})
    assert s["eps_prim_inf"] == 1e-5 and s["eps_dual_inf"] == 1e-6
