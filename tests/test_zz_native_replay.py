"""
The native B&B replay (include/bqp.h `bqp_bnb_solve`, miosqp_b200/csrc/bqp_bnb.cpp; `settings['replay'] = 'native'`)
against (a) the goldens of the UNMODIFIED reference package on the CPU oracle, (b) the Python replay of tree.py
including its look-ahead bookkeeping, (c) the closed-loop MPC golden.  On a GPU-less box the C++ loop gets its
node results from the oracle through the `bqp_solve_fn` hook; the `gpu` tests run it on the CUDA engine.
Required: identical branching decisions, node counts, ADMM iteration totals and statuses; objective and solution
to 1e-9 (the C++ objective sums in a different order than numpy's dot: last-digit differences).
"""
import json
import os

import numpy as np
import pytest

from miosqp_b200 import problems

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "bnb_random_miqp.json")) as f:
    GOLDEN = json.load(f)
G_MPC = np.load(os.path.join(HERE, "golden", "mpc_power_converter.npz"))


@pytest.fixture
def cpu_engine(monkeypatch):
    import fake_engine
    from miosqp_b200 import engine
    monkeypatch.setattr(engine, "BatchedQP", fake_engine.FakeBatchedQP)
    monkeypatch.setattr(engine, "solve_multi", fake_engine.solve_multi)


def _solve(pr, x0=None, **settings):
    import miosqp_b200
    s = miosqp_b200.MIOSQP()
    s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
            dict(problems.RANDOM_MIQP_SETTINGS, **settings), dict(problems.RANDOM_MIQP_QP_SETTINGS))
    if x0 is not None:
        s.set_x0(x0)
    return s.solve(), s.work


def _golden_case(name, tol, **settings):
    c = GOLDEN[name]["case"]; g = GOLDEN[name]["result"]
    pr = problems.random_miqp(c["n"], c["m"], c["p"], c["density"], seed=c["seed"])[0]
    r, w = _solve(pr, replay='native', **settings)
    assert [list(d) for d in w.decisions] == [list(d) for d in g["decisions"]]
    assert r.status == g["status"] and w.iter_num == g["iter_num"] and w.osqp_iter == g["osqp_iter"]
    assert abs(r.upper_glob - g["upper_glob"]) <= tol * (1 + abs(g["upper_glob"]))
    assert np.abs(r.x - np.array(g["x"])).max() <= tol * (1 + np.abs(np.array(g["x"])).max())
    assert abs(r.osqp_iter_avg - g["osqp_iter_avg"]) < 1e-12
    return w


@pytest.mark.parametrize("name", sorted(GOLDEN))
@pytest.mark.parametrize("speculation", [0, 16])
def test_native_replay_matches_reference_cpu(cpu_engine, name, speculation):
    w = _golden_case(name, 1e-9, speculation=speculation)
    assert w.leaves == [] and w.batches >= 1


def test_second_solve_without_update_returns_the_stored_result_cpu(cpu_engine):
    """solve() twice without update_vectors: the reference's loop finds no leaf left and returns what it has; so must the
    native replay (it used to rebuild the tree from its root with the old incumbent: other counters, other averages)"""
    import miosqp_b200
    pr = problems.random_miqp(40, 40, 20, 0.7, seed=3)[0]
    outs = {}
    for replay in (None, 'native'):
        m = miosqp_b200.MIOSQP()
        m.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
                dict(problems.RANDOM_MIQP_SETTINGS, replay=replay), dict(problems.RANDOM_MIQP_QP_SETTINGS))
        r1 = m.solve(); w = m.work
        first = (r1.status, float(r1.upper_glob), w.iter_num, int(w.osqp_iter), w.batches, float(r1.osqp_iter_avg))
        r2 = m.solve()
        assert (r2.status, float(r2.upper_glob), w.iter_num, int(w.osqp_iter), w.batches, float(r2.osqp_iter_avg)) == first
        assert np.array_equal(r1.x, r2.x)
        outs[replay] = first
    a, b = outs[None], outs['native']
    assert a[0] == b[0] and a[2:] == b[2:] and abs(a[1] - b[1]) <= 1e-12 * (1 + abs(a[1]))   # (objective summed in another order)


@pytest.mark.parametrize("rule", [0, 1])
def test_native_equals_python_replay_with_lookahead_cpu(cpu_engine, rule):
    """Same launches, same look-ahead choices, same adoptions as tree.py -- not just the same answer."""
    pr = problems.random_miqp(40, 40, 20, 0.7, seed=3)[0]
    for spec in (0, 32):
        rp, wp = _solve(pr, speculation=spec, tree_explor_rule=rule)
        rn, wn = _solve(pr, speculation=spec, tree_explor_rule=rule, replay='native')
        assert [tuple(d) for d in wp.decisions] == [tuple(d) for d in wn.decisions]
        assert (rp.status, wp.iter_num, wp.osqp_iter) == (rn.status, wn.iter_num, wn.osqp_iter)
        assert (wp.batches, wp.batched_nodes, wp.spec_nodes, wp.spec_hits) == (wn.batches, wn.batched_nodes, wn.spec_nodes, wn.spec_hits)
        assert abs(rp.upper_glob - rn.upper_glob) <= 1e-12 * (1 + abs(rp.upper_glob)) and np.abs(rp.x - rn.x).max() <= 1e-12
        assert abs(wp.lower_glob - wn.lower_glob) <= 1e-12 * (1 + abs(wp.lower_glob))


def test_native_statuses_and_errors_cpu(cpu_engine, capsys):
    pr = problems.random_miqp(40, 40, 20, 0.7, seed=3)[0]
    for limit in (30, 2):                                                   # node limit with / without incumbent
        rp, wp = _solve(pr, max_iter_bb=limit)
        rn, wn = _solve(pr, max_iter_bb=limit, replay='native')
        assert rp.status == rn.status and wp.iter_num == wn.iter_num == limit and wp.osqp_iter == wn.osqp_iter
    easy = problems.random_miqp(30, 60, 10, 0.7, seed=1)[0]
    for key, msg in (("tree_explor_rule", 'Tree exploring strategy not recognized'), ("branching_rule", 'No variable selection rule recognized!')):
        with pytest.raises(ValueError, match=msg):                          # workspace.py:147, 224
            _solve(easy, replay='native', **{key: 7})
    # infeasible MIQP
    import scipy.sparse as spa
    bad = problems.random_miqp(20, 30, 4, 0.7, seed=5)[0]
    bad['A'] = spa.vstack([bad['A'], bad['A'][0]]).tocsc()
    bad['l'] = np.append(bad['l'], 100.0); bad['u'] = np.append(bad['u'], 200.0)
    rn, wn = _solve(bad, replay='native')
    assert rn.status == 'Primal Infeasible' and np.isinf(rn.upper_glob)
    # incumbent seeding through set_x0 (workspace.py:94-111): valid seed prunes, invalid one is reported and ignored
    r0, w0 = _solve(easy, replay='native')
    r1, w1 = _solve(easy, x0=np.copy(r0.x), replay='native')
    assert r1.status == 'Solved' and w1.iter_num <= w0.iter_num and abs(r1.upper_glob - r0.upper_glob) <= 1e-9
    capsys.readouterr()
    r2, w2 = _solve(easy, x0=np.full(30, 0.5), replay='native')
    assert 'Invalid initial solution!' in capsys.readouterr().out and w2.iter_num == w0.iter_num


def _mpc_closed_loop(steps, tol, speculation):
    from miosqp_b200 import power_converter as pc
    dec = []
    res = pc.closed_loop(steps, N=10, speculation=speculation, replay='native',
                         on_step=lambda k, s, r: dec.append([list(d) for d in s.work.decisions]))
    for k in range(steps):
        assert np.array_equal(np.array(dec[k], dtype=np.int64).reshape(-1, 2), G_MPC["dec_%d" % k]), "step %d" % k
        assert [res.nodes[k] + 1, res.admm_iters[k]] == list(G_MPC["stats_%d" % k])
        assert abs(res.obj[k] - float(G_MPC["obj_%d" % k])) <= tol * (1 + abs(float(G_MPC["obj_%d" % k])))
    return res


@pytest.mark.parametrize("speculation", [0, 32])
def test_native_mpc_closed_loop_cpu(cpu_engine, speculation):
    """update_vectors + set_x0 + native solve, step after step (power_converter.py:421-476), against the reference's golden"""
    res = _mpc_closed_loop(4, 1e-9, speculation)
    assert set(np.unique(res.U[:3])) <= {-1., 0., 1.}


def test_abi_argument_checks():
    import ctypes as C
    from miosqp_b200 import engine
    res = engine._BnbResult(); st = engine._BnbSettings(1e-3, 10, 1, 0, 0, 1e-3)
    x = np.zeros(4)
    assert engine.lib().bqp_bnb_solve(None, None, C.byref(st), None, float("inf"), None, None, engine._d(x), C.byref(res), None, 0) == -1
    assert b"Tree exploring" in engine.lib().bqp_strerror(-7) and b"variable selection" in engine.lib().bqp_strerror(-8)


# ------------------------------------------------------------------ the real engine
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_native_replay_on_engine(name):
    _golden_case(name, 1e-9, speculation=16)


@pytest.mark.gpu
def test_native_mpc_closed_loop_on_engine():
    _mpc_closed_loop(8, 1e-9, 32)


def test_native_replay_under_sanitizers(tmp_path):
    """csrc/bqp_bnb.cpp compiled with -fsanitize=address,undefined against a fake solve function (separable QPs whose
    relaxation is a clip, so every optimum is known): 200 random MIQPs x look-ahead budgets x rules x node limits."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(HERE)
    exe = str(tmp_path / "harness")
    cmd = ["g++", "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
           "-I", os.path.join(root, "include"), os.path.join(HERE, "native", "bnb_asan_harness.cpp"),
           os.path.join(root, "miosqp_b200", "csrc", "bqp_bnb.cpp"), "-o", exe]
    build = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if build.returncode != 0 and "sanitize" in build.stdout:
        pytest.skip("sanitizer runtime not installed")
    assert build.returncode == 0, build.stdout
    run = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert run.returncode == 0 and "asan harness ok" in run.stdout, run.stdout[-2000:]


@pytest.mark.gpu
def test_setup_many_on_device_equals_one_by_one():
    """bqp_setup_many (parallel host halves, sequential uploads) then one launch over all problems: same results as
    problems set up one by one."""
    from miosqp_b200 import engine
    shapes = [(130, 200, 10, 0.7, 4), (50, 100, 5, 0.7, 1), (130, 200, 10, 0.7, 5)]
    items = [problems.extend(problems.random_miqp(n, m, p, d, seed=seed)[0]) for (n, m, p, d, seed) in shapes]
    qp = dict(problems.RANDOM_MIQP_QP_SETTINGS)
    many = engine.setup_many(items, **qp)
    single = [engine.BatchedQP().setup(P, q, A, l, u, i_idx=i, **qp) for (P, q, A, l, u, i) in items]
    L = [it[3] for it in items]; U = [it[4] for it in items]
    X0 = [np.zeros(it[2].shape[1]) for it in items]; Y0 = [np.zeros(it[2].shape[0]) for it in items]
    xa, ya, sa = engine.solve_multi(many, L, U, X0, Y0)
    xb, yb, sb = engine.solve_multi(single, L, U, X0, Y0)
    assert list(sa.status) == list(sb.status) and list(sa.iters) == list(sb.iters)
    for a, b in zip(xa, xb):
        assert np.array_equal(a, b, equal_nan=True)


def _many(replay, speculation, seeds=(3, 4, 5)):
    import miosqp_b200
    solvers = []
    for seed in seeds:
        pr = problems.random_miqp(40, 40, 20, 0.7, seed=seed)[0]
        s = miosqp_b200.MIOSQP()
        s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
                dict(problems.RANDOM_MIQP_SETTINGS, speculation=speculation, replay=replay), dict(problems.RANDOM_MIQP_QP_SETTINGS))
        solvers.append(s)
    return solvers, miosqp_b200.solve_many(solvers)


def _same_many(a, b, tol, counts=True):
    (sa, ra), (sb, rb) = a, b
    for s1, s2, r1, r2 in zip(sa, sb, ra, rb):
        assert [tuple(d) for d in s1.work.decisions] == [tuple(d) for d in s2.work.decisions]
        assert (r1.status, s1.work.iter_num, s1.work.osqp_iter) == (r2.status, s2.work.iter_num, s2.work.osqp_iter)
        if counts:
            assert (s1.work.batches, s1.work.batched_nodes, s1.work.spec_hits) == (s2.work.batches, s2.work.batched_nodes, s2.work.spec_hits)
        assert abs(r1.upper_glob - r2.upper_glob) <= tol * (1 + abs(r1.upper_glob)) and np.abs(r1.x - r2.x).max() <= tol


@pytest.mark.parametrize("speculation", [0, 16])
def test_native_lock_step_equals_python_solve_many_cpu(cpu_engine, monkeypatch, speculation):
    """bqp_bnb_solve_many against miqp.solve_many: same per-instance B&B, same launches, same look-ahead adoptions; and
    each instance equals its own single solve."""
    import fake_engine
    from miosqp_b200 import engine
    monkeypatch.setattr(engine, "native_solve_many_fn", fake_engine.native_solve_many_fn)
    py = _many(None, speculation); nat = _many('native', speculation)
    _same_many(py, nat, 1e-12)
    pr = problems.random_miqp(40, 40, 20, 0.7, seed=4)[0]
    r1, w1 = _solve(pr, replay='native', speculation=speculation)
    assert [tuple(d) for d in w1.decisions] == [tuple(d) for d in nat[0][1].work.decisions] and r1.upper_glob == nat[1][1].upper_glob


@pytest.mark.gpu
def test_native_lock_step_on_engine():
    _same_many(_many(None, 8), _many('native', 8), 1e-9, counts=False)
