"""
GPU: explicit solve contexts (bqp_ctx_*), rolling sessions (bqp_session_*) and the three multi-tree B&B drivers
(lock-step, asynchronous on one stream per tree, rolling session) against each other, against the goldens of the
UNMODIFIED reference package, and -- BASELINE config 2's size -- against tests/golden/bnb_cfg2.json.
"""
import ctypes as C
import json
import os
import threading

import numpy as np
import pytest

import miosqp_b200
from miosqp_b200 import engine, problems, miqp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
QP = dict(eps_abs=1e-3, eps_rel=1e-3, eps_prim_inf=1e-4)


def _nodes(n, m, p, seed, count):
    P, q, A, l, u, i_idx = problems.extend(problems.random_miqp(n, m, p, 0.7, seed=seed)[0])
    ls, us = problems.branched_nodes(l, u, len(i_idx), count, np.random.default_rng(seed))
    e = engine.BatchedQP().setup(P, q, A, l, u, i_idx=i_idx, **QP)
    return e, ls, us, n, A.shape[0]


def test_context_solve_matches_default_context_bitwise():
    e, ls, us, n, m = _nodes(130, 200, 10, 4, 7)
    x0 = [np.zeros(n)] * 7; y0 = [np.zeros(m)] * 7
    xs, ys, sc = engine.solve_multi([e] * 7, list(ls), list(us), x0, y0)
    ctx = engine.Context(0, run_to_completion=True)
    xc, yc, scc = ctx.solve_multi([e] * 7, list(ls), list(us), x0, y0)
    assert ctx.last_timing()["launches"] == 1                      # run to completion: no rounds
    assert list(sc.status) == list(scc.status) and list(sc.iters) == list(scc.iters)
    for a, b in zip(xs + ys, xc + yc):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.array_equal(sc.lower, scc.lower, equal_nan=True)
    ctx.free()


def test_contexts_from_threads_do_not_interfere():
    """4 host threads, one context each, different problems: every thread gets what the default context computes alone"""
    cases = [_nodes(60 + 20 * k, 100 + 10 * k, 5, 10 + k, 5) for k in range(4)]
    want = []
    for e, ls, us, n, m in cases:
        want.append(engine.solve_multi([e] * 5, list(ls), list(us), [np.zeros(n)] * 5, [np.zeros(m)] * 5))
    got = [None] * 4
    ctxs = [engine.Context(0) for _ in range(4)]

    def work(k):
        e, ls, us, n, m = cases[k]
        for _ in range(3):
            got[k] = ctxs[k].solve_multi([e] * 5, list(ls), list(us), [np.zeros(n)] * 5, [np.zeros(m)] * 5)
    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in th]; [t.join() for t in th]
    for k in range(4):
        assert list(want[k][2].iters) == list(got[k][2].iters)
        for a, b in zip(want[k][0] + want[k][1], got[k][0] + got[k][1]):
            assert np.array_equal(a, b, equal_nan=True)


def test_session_appends_while_running(monkeypatch):
    """nodes appended to an open session in two instalments, solved in rounds of 25 iterations: same results, bit for bit,
    as one closed batch (clusters fixed: the automatic cluster size of sessions changes the summation order)"""
    monkeypatch.setenv("BQP_ROWS_AUTO_CLUSTER", "0")
    monkeypatch.setenv("BQP_ROUND_ITERS", "25")
    e, ls, us, n, m = _nodes(130, 200, 10, 4, 9)
    x0 = [np.zeros(n)] * 9; y0 = [np.zeros(m)] * 9
    xs, ys, sc = engine.solve_multi([e] * 9, list(ls), list(us), x0, y0)
    L = engine.lib()
    pa = engine._ptr_array
    assert L.bqp_session_begin(None) == 0
    first = C.c_int(-1)
    hs = (C.c_void_p * 9)(*[e._h.value] * 9)
    keep = [[np.ascontiguousarray(v) for v in seq] for seq in (ls, us, x0, y0)]
    assert L.bqp_session_append(None, 5, hs, pa(keep[0][:5]), pa(keep[1][:5]), pa(keep[2][:5]), pa(keep[3][:5]), C.byref(first)) == 0
    assert first.value == 0
    fin = (C.c_int * 16)(); nfin = C.c_int(0); running = C.c_int(0)
    done = []
    for _ in range(3):                                              # three rounds with the first five nodes only
        assert L.bqp_session_round(None, fin, 16, C.byref(nfin), C.byref(running)) == 0
        done += [fin[k] for k in range(nfin.value)]
    assert L.bqp_session_append(None, 4, hs, pa(keep[0][5:]), pa(keep[1][5:]), pa(keep[2][5:]), pa(keep[3][5:]), C.byref(first)) == 0
    assert first.value == 5
    for _ in range(400):
        assert L.bqp_session_round(None, fin, 16, C.byref(nfin), C.byref(running)) == 0
        done += [fin[k] for k in range(nfin.value)]
        if running.value == 0:
            break
    assert sorted(done) == list(range(9))
    for b in range(9):
        x = np.empty(n); y = np.empty(m); st = np.zeros(1, np.int32); it = np.zeros(1, np.int32); lo = np.zeros(1)
        out = engine._NodeOut(engine._i(st), engine._i(it), None, None, None, engine._d(lo))
        assert L.bqp_session_fetch(None, b, engine._d(x), engine._d(y), C.byref(out)) == 0
        assert st[0] == sc.status[b] and it[0] == sc.iters[b]
        assert np.array_equal(x, xs[b], equal_nan=True) and np.array_equal(y, ys[b], equal_nan=True)
        assert np.array_equal(lo[0], sc.lower[b], equal_nan=True)


def _trees(cases, speculation=0):
    prs = [problems.random_miqp(c["n"], c["m"], c["p"], c["density"], seed=c["seed"])[0] for c in cases]
    out = []
    for pr in prs:
        s = miosqp_b200.MIOSQP()
        s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
                dict(problems.RANDOM_MIQP_SETTINGS, replay='native', speculation=speculation), dict(problems.RANDOM_MIQP_QP_SETTINGS))
        out.append(s)
    return out


def _sig(res, solvers):
    return [(r.status, s.work.iter_num, int(s.work.osqp_iter), [tuple(d) for d in s.work.decisions]) for r, s in zip(res, solvers)]


@pytest.mark.parametrize("speculation", [0, 6])
def test_three_drivers_same_trees_and_goldens(speculation):
    with open(os.path.join(HERE, "golden", "bnb_random_miqp.json")) as f:
        gold = json.load(f)
    names = sorted(gold)
    runs = {}
    for driver in ("lockstep", "async", "rolling"):
        solvers = _trees([gold[k]["case"] for k in names], speculation)
        res = miosqp_b200.solve_many(solvers, async_threads=(0 if driver == "async" else None), rolling=(driver == "rolling"))
        runs[driver] = (_sig(res, solvers), [float(r.upper_glob) for r in res], [np.array(r.x) for r in res])
        for s in solvers:
            s.work.solver.free()
    for k, name in enumerate(names):
        g = gold[name]["result"]
        for driver in runs:
            st, it, oi, dec = runs[driver][0][k]
            assert dec == [tuple(d) for d in g["decisions"]], (driver, name)
            assert st == g["status"] and it == g["iter_num"] and oi == g["osqp_iter"], (driver, name)
            assert abs(runs[driver][1][k] - g["upper_glob"]) <= 1e-9 * (1 + abs(g["upper_glob"]))
            assert np.abs(runs[driver][2][k] - np.array(g["x"])).max() <= 1e-9 * (1 + np.abs(np.array(g["x"])).max())


def test_cfg2_size_bnb_matches_reference_golden():
    """BASELINE config 2's size (n=500, m=1000, |i_idx|=50): the first instance of the 100-instance workload, B&B to completion
    on the engine, against the UNMODIFIED reference package run on the CPU oracle (tests/golden/make_bnb_golden.py --cfg2):
    identical branching sequence, node and ADMM-iteration counts, incumbent to 1e-9."""
    with open(os.path.join(HERE, "golden", "bnb_cfg2.json")) as f:
        gold = json.load(f)
    prs = problems.random_miqp(500, 1000, 50, 0.7, seed=1, count=2)
    solvers = miqp.setup_many(prs, dict(problems.RANDOM_MIQP_SETTINGS, replay='native'), dict(problems.RANDOM_MIQP_QP_SETTINGS))
    res = miosqp_b200.solve_many(solvers, rolling=True)
    for k in range(2):
        g = gold["cfg2_inst%d" % k]["result"]
        w = solvers[k].work
        assert [tuple(d) for d in w.decisions] == [tuple(d) for d in g["decisions"]]
        assert res[k].status == g["status"] and w.iter_num == g["iter_num"] and int(w.osqp_iter) == g["osqp_iter"]
        assert abs(res[k].upper_glob - g["upper_glob"]) <= 1e-9 * (1 + abs(g["upper_glob"]))
        assert np.abs(res[k].x - np.array(g["x"])).max() <= 1e-9 * (1 + np.abs(np.array(g["x"])).max())
    # the Python replay of one of them (tree.py) takes the same decisions
    s = miosqp_b200.MIOSQP()
    pr = prs[0]
    s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
            dict(problems.RANDOM_MIQP_SETTINGS), dict(problems.RANDOM_MIQP_QP_SETTINGS))
    r = s.solve()
    g = gold["cfg2_inst0"]["result"]
    assert [tuple(d) for d in s.work.decisions] == [tuple(d) for d in g["decisions"]] and s.work.osqp_iter == g["osqp_iter"]
    assert r.status == g["status"]


def test_cfg2_size_bnb_adaptive_rho_matches_reference_golden():
    """The same under osqp's adaptive_rho (fixed interval 50): the three first config-2 instances, B&B to completion on the
    rolling session (rows kernel, spectral inverse, automatic cluster sizes), against the UNMODIFIED reference package on the
    adaptive CPU oracle (make_bnb_golden.py --cfg2 3 --adaptive 50): identical decisions, node and iteration counts -- with
    a fifth of the fixed-rho run's ADMM iterations."""
    with open(os.path.join(HERE, "golden", "bnb_cfg2_adaptive50.json")) as f:
        gold = json.load(f)
    with open(os.path.join(HERE, "golden", "bnb_cfg2.json")) as f:
        fixed = json.load(f)
    prs = problems.random_miqp(500, 1000, 50, 0.7, seed=1, count=3)
    qp = dict(problems.RANDOM_MIQP_QP_SETTINGS, adaptive_rho=True, adaptive_rho_interval=50)
    solvers = miqp.setup_many(prs, dict(problems.RANDOM_MIQP_SETTINGS, replay='native'), qp)
    res = miosqp_b200.solve_many(solvers, rolling=True)
    for k in range(3):
        g = gold["cfg2_inst%d" % k]["result"]
        w = solvers[k].work
        assert [tuple(d) for d in w.decisions] == [tuple(d) for d in g["decisions"]]
        assert res[k].status == g["status"] and w.iter_num == g["iter_num"] and int(w.osqp_iter) == g["osqp_iter"]
        assert abs(res[k].upper_glob - g["upper_glob"]) <= 1e-9 * (1 + abs(g["upper_glob"]))
        assert np.abs(res[k].x - np.array(g["x"])).max() <= 1e-9 * (1 + np.abs(np.array(g["x"])).max())
        assert g["osqp_iter"] * 4 < fixed["cfg2_inst%d" % k]["result"]["osqp_iter"]
