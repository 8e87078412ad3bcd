"""
The native config-3 workload builder (miosqp_b200/power_converter.py) against the reference's own model builder
(fixture tests/golden/power_converter_model.npz, generated from /root/reference/examples/power_converter by
tests/golden/make_power_converter_model_golden.py) and against the closed-loop golden of the unmodified reference
package on the CPU oracle (tests/golden/mpc_power_converter.npz).

Tolerances: matrices 1e-12 relative to the largest entry (the prediction matrices are accumulated in a different
order than the reference's matrix_power loops); closed loop: identical branching decisions, node and ADMM iteration
counts, solutions to 1e-9.
"""
import os

import numpy as np
import pytest

from miosqp_b200 import power_converter as pc

HERE = os.path.dirname(os.path.abspath(__file__))
M = np.load(os.path.join(HERE, "golden", "power_converter_model.npz"))
G = np.load(os.path.join(HERE, "golden", "mpc_power_converter.npz"))


def _close(a, b, tol=1e-12):
    a = np.asarray(a, float); b = np.asarray(b, float)
    assert a.shape == b.shape, (a.shape, b.shape)
    fin = np.isfinite(b)
    assert np.array_equal(fin, np.isfinite(a)) and np.array_equal(a[~fin], b[~fin])
    assert np.abs(a[fin] - b[fin]).max() <= tol * max(1., np.abs(b[fin]).max())


def test_system_matrices_and_initial_state():
    d = pc.Drive()
    s = pc.System(d, 300, 5.5)
    _close(s.A, M["sys_A"]); _close(s.B, M["sys_B"]); _close(s.C, M["sys_C"])
    _close(d.initial_state(), M["x0"])
    assert d.steps_per_period == int(M["Nstpp"])


@pytest.mark.parametrize("N", [1, 3, 10])
def test_mpc_program_matches_reference_builder(N):
    s = pc.System(pc.Drive(), 300, 5.5)
    p = pc.MpcProgram(s, N, pc.TailCost(s, 0.95, "delta_550"))
    _close(p.P.toarray(), M["P_%d" % N]); _close(p.A.toarray(), M["A_%d" % N])
    _close(p.q_x, M["q_x_%d" % N]); _close(p.q_u, M["q_u_%d" % N]); _close(p.SA, M["SA_%d" % N])
    _close(p.l, M["l_%d" % N]); _close(p.u, M["u_%d" % N])
    assert p.P.shape == (6 * N, 6 * N) and p.A.shape == (9 * N, 6 * N)
    assert np.array_equal(p.i_idx, np.arange(6 * N))


def test_default_tail_is_stage_cost():
    s = pc.System(pc.Drive())
    t = pc.TailCost(s, 0.95, None)
    _close(t.P0, s.C.T.dot(s.C)); assert t.r0 == 0. and not t.q0.any()


def test_on_transitions():
    assert pc.on_transitions([1, 0, -1], [0, 0, -1])[0] == 1
    assert pc.on_transitions([0, 0, 0], [1, -1, 0]).tolist() == [0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 0]
    assert pc.on_transitions([0, -1, 0], [0, 0, 0])[7] == 1
    assert pc.on_transitions([1, 1, 1], [1, 1, 1]).sum() == 0
    assert pc.on_transitions([1, 0, 0], [-1, 0, 0]).sum() == 0       # a two-level jump is not counted (utils.py:81-98)


def _check_against_golden(res, steps, tol, decisions):
    for k in range(steps):
        assert np.array_equal(np.array(decisions[k], dtype=np.int64).reshape(-1, 2), G["dec_%d" % k]), "step %d" % k
        assert [res.nodes[k] + 1, res.admm_iters[k]] == list(G["stats_%d" % k])
        assert abs(res.obj[k] - float(G["obj_%d" % k])) <= tol * (1 + abs(float(G["obj_%d" % k])))
    q, l, u = res.program.vectors(res.X[:, steps - 1])
    _close(q, G["q_%d" % (steps - 1)], 1e-10); _close(u, G["u_%d" % (steps - 1)], 1e-10)


def _run(steps):
    decisions = []
    res = pc.closed_loop(steps, N=10, on_step=lambda k, s, r: decisions.append(list(s.work.decisions)))
    return res, decisions


def test_closed_loop_from_native_model_cpu(monkeypatch):
    import fake_engine
    from miosqp_b200 import engine
    monkeypatch.setattr(engine, "BatchedQP", fake_engine.FakeBatchedQP)
    monkeypatch.setattr(engine, "solve_multi", fake_engine.solve_multi)
    res, dec = _run(4)
    _check_against_golden(res, 4, 1e-9, dec)
    assert res.U.shape == (6, 4) and set(np.unique(res.U[:3])) <= {-1., 0., 1.}
    assert res.phase_currents.shape == (3, 4) and np.isfinite(res.torque).all()


@pytest.mark.gpu
def test_closed_loop_from_native_model_engine():
    res, dec = _run(8)
    _check_against_golden(res, 8, 1e-9, dec)
