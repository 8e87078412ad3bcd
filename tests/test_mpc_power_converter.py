"""
BASELINE config 3: the power-converter MPC closed loop (horizon N=10, n=60, m_ext=150), warm-started re-solves
through MIOSQP.setup / update_vectors / set_x0 / solve exactly as Model.compute_mpc_input drives them
(/root/reference/examples/power_converter/power_converter.py:421-476).  Golden answers come from the unmodified
reference package on the CPU oracle (tests/golden/make_mpc_golden.py).  Required per time step: identical
branching sequence, node count and total ADMM iterations; objective and solution to 1e-9.
"""
import os

import numpy as np
import pytest
import scipy.sparse as spa

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "mpc_power_converter.npz"))
SETTINGS = {'eps_int_feas': 1e-02, 'max_iter_bb': 2000, 'tree_explor_rule': 1, 'branching_rule': 0,
            'verbose': False, 'print_interval': 1}                       # power_converter.py:451-459
QP_SETTINGS = {'eps_abs': 1e-03, 'eps_rel': 1e-03, 'eps_prim_inf': 1e-04, 'verbose': False}   # :461-466


def _closed_loop(steps, tol, **extra_settings):
    import hashlib
    import miosqp_b200
    P = spa.csc_matrix(G["P"]); A = spa.csc_matrix(G["A"])
    solver = None
    for k in range(steps):
        q, l, u, x0 = G["q_%d" % k], G["l_%d" % k].copy(), G["u_%d" % k].copy(), G["x0_%d" % k]
        if solver is None:
            solver = miosqp_b200.MIOSQP()
            solver.setup(P, q, A, l, u, G["i_idx"], G["i_l"], G["i_u"], dict(SETTINGS, **extra_settings), dict(QP_SETTINGS))
        else:
            solver.update_vectors(q, l, u)
        solver.set_x0(x0)
        res = solver.solve()
        w = solver.work
        assert res.status == str(G["status_%d" % k]), "step %d" % k      # (17 of the first 1000 instants stop at the node limit)
        dec = np.array(w.decisions, dtype=np.int64).reshape(-1, 2)
        if "dec_%d" % k in G:
            assert np.array_equal(dec, G["dec_%d" % k]), "step %d" % k
        assert np.array_equal(np.frombuffer(hashlib.sha256(dec.tobytes()).digest(), dtype=np.uint8), G["dech_%d" % k]), "step %d" % k
        assert [w.iter_num, w.osqp_iter] == list(G["stats_%d" % k])
        assert abs(res.upper_glob - float(G["obj_%d" % k])) <= tol * (1 + abs(float(G["obj_%d" % k])))
        assert np.abs(res.x - G["sol_%d" % k]).max() <= tol * (1 + np.abs(G["sol_%d" % k]).max())
    return solver


def test_fixture_matches_config3():
    assert G["P"].shape == (60, 60) and G["A"].shape == (90, 60) and len(G["i_idx"]) == 60
    assert int(G["steps"]) == 120


def test_closed_loop_replay_cpu(monkeypatch):
    import fake_engine
    from miosqp_b200 import engine
    monkeypatch.setattr(engine, "BatchedQP", fake_engine.FakeBatchedQP)
    monkeypatch.setattr(engine, "solve_multi", fake_engine.solve_multi)
    _closed_loop(3, 1e-13)


def test_closed_loop_native_replay_cpu_node_limit_steps(monkeypatch):
    """first 24 instants (five of them stop at the reference's node limit max_iter_bb = 2000) through the native replay with
    look-ahead on the CPU stand-in engine: statuses, decisions, counts as the reference's"""
    import fake_engine
    from miosqp_b200 import engine
    monkeypatch.setattr(engine, "BatchedQP", fake_engine.FakeBatchedQP)
    monkeypatch.setattr(engine, "solve_multi", fake_engine.solve_multi)
    _closed_loop(24, 1e-9, replay='native', speculation=16)


@pytest.mark.gpu
def test_closed_loop_engine():
    s = _closed_loop(4, 1e-9)
    assert s.work.batched_nodes >= s.work.iter_num - 1


@pytest.mark.gpu
def test_closed_loop_engine_120_steps_native_lookahead():
    """120 sampling instants of BASELINE config 3 (start-up transient and the beginning of the steady state) against the
    UNMODIFIED reference package on the CPU oracle: per instant the same status, branching sequence (sha256), node and ADMM
    iteration counts, objective and plan -- with the native replay and a look-ahead of 32 nodes per launch."""
    s = _closed_loop(120, 1e-9, replay='native', speculation=32)
    assert s.work.spec_hits > 0
