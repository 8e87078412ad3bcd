"""
BASELINE config 4 (random_miqp n=2000 m=4000 |i_idx|=200, 5 % dense; the stream kernel's size) through the B&B, against
the UNMODIFIED reference package on the CPU oracle stopped at its own node limit (tests/golden/bnb_cfg4_nodelimit.json,
made by `python tests/golden/make_bnb_golden.py --cfg4`): identical branching decisions, node count, ADMM iteration
total and status; bounds to 1e-9.  The GPU test is collected last (written after the round's last GPU run).
"""
import json
import os

import numpy as np
import pytest

from miosqp_b200 import problems

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "bnb_cfg4_nodelimit.json")) as f:
    G = json.load(f)


def _run(**extra):
    import miosqp_b200
    c, g = G["case"], G["result"]
    pr = problems.random_miqp(c["n"], c["m"], c["p"], c["density"], seed=c["seed"])[0]
    s = miosqp_b200.MIOSQP()
    s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
            dict(problems.RANDOM_MIQP_SETTINGS, max_iter_bb=c["max_iter_bb"], **extra), dict(problems.RANDOM_MIQP_QP_SETTINGS))
    r = s.solve()
    w = s.work
    assert [list(d) for d in w.decisions] == g["decisions"]
    assert r.status == g["status"] and w.iter_num == g["iter_num"] and w.osqp_iter == g["osqp_iter"]
    assert (np.isinf(r.upper_glob) and np.isinf(g["upper_glob"])) or abs(r.upper_glob - g["upper_glob"]) <= 1e-9 * (1 + abs(g["upper_glob"]))
    assert abs(w.lower_glob - g["lower_glob"]) <= 1e-9 * (1 + abs(g["lower_glob"]))
    return w


def test_cfg4_node_limited_bnb_cpu(monkeypatch):
    import fake_engine
    from miosqp_b200 import engine
    monkeypatch.setattr(engine, "BatchedQP", fake_engine.FakeBatchedQP)
    monkeypatch.setattr(engine, "solve_multi", fake_engine.solve_multi)
    w = _run(replay='native', speculation=4)
    assert w.spec_nodes > 0          # look-ahead nodes rode along (a 9-node dive never returns to them: no adoption expected)


@pytest.mark.gpu
def test_cfg4_node_limited_bnb_stream_kernel(monkeypatch):
    from miosqp_b200 import engine
    monkeypatch.setenv("BQP_GRID", "0")
    _run(speculation=4)
    assert engine.last_timing()["kernel"] == 1          # the TMA-streamed two-sweep kernel (one CTA per tile)


@pytest.mark.gpu
def test_cfg4_node_limited_bnb_engine():
    from miosqp_b200 import engine
    w = _run(speculation=4)
    assert engine.last_timing()["kernel"] == 4          # the whole-GPU kernel (n > 512: one tile on every SM)
    assert w.spec_nodes > 0          # look-ahead nodes rode along (a 9-node dive never returns to them: no adoption expected)
