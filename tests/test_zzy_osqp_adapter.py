"""
The sequential `osqp`-shaped adapter (miosqp_b200/osqp_compat.py): the six-call surface the reference uses on its
solver object (workspace.py:63-68, node.py:102-125, solver.py:185, osqp.constant), SURVEY section 8b(i).
  * CPU: the UNMODIFIED reference package (where /root/reference is mounted) runs on the adapter -- with the oracle-backed
    stand-in below it -- and must reproduce the golden B&B bit for bit in its decisions.
  * GPU: the same call sequence against the oracle's osqp object, node by node (collected after the suites that were
    green on a B200 before this file was written).
"""
import json
import os
import sys

import numpy as np
import pytest

from miosqp_b200 import problems

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "bnb_random_miqp.json")) as f:
    GOLDEN = json.load(f)


def test_constants_match_osqp_values():
    from miosqp_b200 import osqp_compat
    want = {"OSQP_SOLVED": 1, "OSQP_MAX_ITER_REACHED": -2, "OSQP_PRIMAL_INFEASIBLE": -3, "OSQP_DUAL_INFEASIBLE": -4, "OSQP_UNSOLVED": -10}
    for k, v in want.items():
        assert osqp_compat.constant(k) == v


@pytest.mark.parametrize("name", ["cfg1_seed1", "small_seed5"])
def test_unmodified_reference_on_the_adapter_cpu(monkeypatch, name):
    if not os.path.isdir("/root/reference/miosqp"):
        pytest.skip("reference tree not present on this box")
    import fake_engine
    from miosqp_b200 import engine
    monkeypatch.setattr(engine, "BatchedQP", fake_engine.FakeBatchedQP)
    sys.path.insert(0, os.path.join(HERE, "osqp_shim"))
    if "/root/reference" not in sys.path:
        sys.path.insert(1, "/root/reference")
    import osqp
    osqp.set_backend("b200")
    try:
        import miosqp
        from miosqp import workspace as ref_ws
        decisions = []
        orig = ref_ws.Workspace.branch

        def branch(self, leaf):
            self.pick_nextvar(leaf)
            decisions.append([int(leaf.constr_idx), int(leaf.nextvar_idx)])
            self.add_left(leaf); self.add_right(leaf)
        monkeypatch.setattr(ref_ws.Workspace, "branch", branch)
        c = GOLDEN[name]["case"]; g = GOLDEN[name]["result"]
        pr = problems.random_miqp(c["n"], c["m"], c["p"], c["density"], seed=c["seed"])[0]
        s = miosqp.MIOSQP()
        s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
                dict(problems.RANDOM_MIQP_SETTINGS), dict(problems.RANDOM_MIQP_QP_SETTINGS))
        r = s.solve()
        assert decisions == [list(d) for d in g["decisions"]]
        assert r.status == g["status"] and s.work.iter_num == g["iter_num"] and s.work.osqp_iter == g["osqp_iter"]
        assert abs(r.upper_glob - g["upper_glob"]) <= 1e-12 * (1 + abs(g["upper_glob"]))
        del orig
    finally:
        osqp.set_backend("oracle")


def _adapter_sequence(make):
    """setup -> solve -> update(l,u) + warm_start -> solve -> update(q) -> solve, as Node.solve / update_vectors drive it."""
    pr = problems.random_miqp(50, 100, 5, 0.7, seed=1)[0]
    P, q, A, l, u, i_idx = problems.extend(pr)
    o = make()
    o.setup(P, q, A, l, u, **problems.RANDOM_MIQP_QP_SETTINGS)
    out = []
    r = o.solve(); out.append(r)
    l2, u2 = l.copy(), u.copy(); u2[-2] = 0.0; l2[-1] = 1.0          # two branched binaries
    o.update(l=l2, u=u2); o.warm_start(x=r.x, y=r.y)
    r2 = o.solve(); out.append(r2)
    o.update(q=q * 0.5 + 0.1); o.warm_start(x=r2.x, y=r2.y)
    out.append(o.solve())
    with pytest.raises(ValueError):
        bad = l.copy(); bad[0] = u[0] + 1.0
        o.update(l=bad, u=u)
    return out


def test_adapter_call_sequence_cpu(monkeypatch):
    import fake_engine
    from miosqp_b200 import engine, osqp_compat
    from oracle import oracle
    monkeypatch.setattr(engine, "BatchedQP", fake_engine.FakeBatchedQP)
    a = _adapter_sequence(osqp_compat.OSQP); b = _adapter_sequence(oracle.OSQP)
    for ra, rb in zip(a, b):
        assert ra.info.status_val == rb.info.status_val and ra.info.iter == rb.info.iter
        assert np.abs(ra.x - rb.x).max() <= 1e-12 * (1 + np.abs(rb.x).max())


@pytest.mark.gpu
def test_adapter_call_sequence_engine():
    from miosqp_b200 import osqp_compat
    from oracle import oracle
    a = _adapter_sequence(osqp_compat.OSQP); b = _adapter_sequence(oracle.OSQP)
    for ra, rb in zip(a, b):
        assert ra.info.status_val == rb.info.status_val and ra.info.iter == rb.info.iter
        assert np.abs(ra.x - rb.x).max() <= 1e-9 * (1 + np.abs(rb.x).max())
        assert np.abs(ra.y - rb.y).max() <= 1e-9 * (1 + np.abs(rb.y).max())
