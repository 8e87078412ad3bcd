"""
The example harnesses (examples/*.py: the reference's run_example / run_maxiter_problem scripts without Gurobi and
plots) and the host-side look-ahead (`settings['speculation']`, miosqp_b200/tree.py).  On a GPU-less box the engine
is replaced by the oracle-backed test stand-in; the `gpu` tests run the same harnesses on the CUDA engine.
"""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "examples"))


@pytest.fixture
def cpu_engine(monkeypatch):
    import fake_engine
    from miosqp_b200 import engine
    monkeypatch.setattr(engine, "BatchedQP", fake_engine.FakeBatchedQP)
    monkeypatch.setattr(engine, "solve_multi", fake_engine.solve_multi)
    monkeypatch.setattr(engine, "native_solve_many_fn", fake_engine.native_solve_many_fn)


def _mpc(capsys, argv):
    import power_converter_mpc
    power_converter_mpc.main(argv)
    return json.loads(capsys.readouterr().out.strip().splitlines()[-1])["rows"]


def _check_mpc_rows(rows):
    assert [r["T"] for r in rows] == [1, 3]
    for r in rows:
        assert r["node_limit_steps"] == 0 and r["nodes_per_step"] >= 1 and r["launches_per_step"] >= 1
        assert r["miosqp_min"] <= r["miosqp_avg"] <= r["miosqp_max"]
    assert rows[1]["spec_hit_rate"] > 0.3 and rows[1]["launches_per_step"] < rows[1]["nodes_per_step"]


def test_power_converter_example_cpu(cpu_engine, capsys):
    _check_mpc_rows(_mpc(capsys, ["--horizons", "1,3", "--steps", "40", "--speculation", "64"]))


def test_power_converter_example_native_replay_cpu(cpu_engine, capsys):
    a = _mpc(capsys, ["--horizons", "1,3", "--steps", "25", "--speculation", "16"])
    b = _mpc(capsys, ["--horizons", "1,3", "--steps", "25", "--speculation", "16", "--replay", "native"])
    for ra, rb in zip(a, b):
        for key in ("nodes_per_step", "admm_iters_per_step", "launches_per_step", "solved_nodes_per_step", "fsw_hz", "node_limit_steps"):
            assert ra[key] == rb[key], key


def test_random_miqp_example_cpu(cpu_engine, capsys):
    import random_miqp
    random_miqp.main(["--sizes", "10,5,2;50,25,5", "--repeat", "3", "--replay", "native"])
    assert "t_osqp_avg" in capsys.readouterr().out
    random_miqp.main(["--sizes", "10,5,2;50,25,5", "--repeat", "3"])
    out = capsys.readouterr().out
    assert "t_osqp_avg" in out and "one instance at a time" in out
    random_miqp.main(["--sizes", "50,25,5", "--repeat", "3", "--together", "--speculation", "8"])
    assert "lock-step" in capsys.readouterr().out
    random_miqp.main(["--sizes", "50,25,5", "--repeat", "3", "--together", "--speculation", "8", "--replay", "native"])
    assert "lock-step" in capsys.readouterr().out


def test_replay_maxiter_example_cpu(cpu_engine, capsys):
    import replay_maxiter
    replay_maxiter.main(["--only", "28,76"])
    out = capsys.readouterr().out
    assert "2 problems in one launch" in out and "summary:" in out


def test_frontier_split_example_single_process_cpu(cpu_engine, capsys):
    import frontier_split
    frontier_split.main(["--vars", "40", "--rows", "40", "--ints", "20", "--density", "0.7", "--seed", "3", "--speculation", "32"])
    rec = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert rec["status"] == "Solved" and rec["nodes"] >= 100 and rec["launches"] * 3 < rec["nodes"]


def _bnb(speculation, rule=1):
    import miosqp_b200
    from miosqp_b200 import problems
    pr = problems.random_miqp(40, 40, 20, 0.7, seed=3)[0]          # ~110 nodes: a tree worth looking ahead in
    s = miosqp_b200.MIOSQP()
    s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
            dict(problems.RANDOM_MIQP_SETTINGS, speculation=speculation, tree_explor_rule=rule), dict(problems.RANDOM_MIQP_QP_SETTINGS))
    r = s.solve()
    return r, s.work


@pytest.mark.parametrize("rule", [0, 1])
def test_speculation_changes_nothing_but_the_launch_count_cpu(cpu_engine, rule):
    """Look-ahead results are adopted only when computed from exactly the inputs the replay produces, so the B&B
    (decisions, node and iteration counts, incumbent) is identical; only the number of launches drops."""
    r0, w0 = _bnb(0, rule)
    r1, w1 = _bnb(32, rule)
    assert w0.decisions == w1.decisions and w0.iter_num == w1.iter_num and w0.osqp_iter == w1.osqp_iter
    assert r0.status == r1.status and r0.upper_glob == r1.upper_glob and np.array_equal(r0.x, r1.x)
    assert w0.spec_nodes == 0 and w1.spec_nodes > 0 and w1.spec_hits > 0
    assert w1.batches * 2 < w0.batches and w1.batched_nodes >= w0.batched_nodes


def test_speculation_closed_loop_identical_cpu(cpu_engine):
    from miosqp_b200 import power_converter as pc
    runs = []
    for spec in (0, 16, 256):
        dec = []
        r = pc.closed_loop(6, N=10, speculation=spec, on_step=lambda k, s, rr: dec.append((list(s.work.decisions), s.work.iter_num, s.work.osqp_iter)))
        runs.append((dec, r.U, r.obj, r.X, r.solver.work.batches))
    for other in runs[1:]:
        assert other[0] == runs[0][0]
        for a, b in zip(other[1:4], runs[0][1:4]):
            assert np.array_equal(a, b)
        assert other[4] < runs[0][4]
    assert runs[2][4] * 4 < runs[0][4]          # 256 nodes of look-ahead: > 4x fewer launches on this workload


def test_speculation_is_rank_consistent_cpu(cpu_engine):
    """Shadow nodes are a deterministic function of cached results, so replicated replays (multi-GPU frontier
    splitting) build the same launch lists."""
    from miosqp_b200 import power_converter as pc
    a = pc.closed_loop(3, N=3, speculation=8); b = pc.closed_loop(3, N=3, speculation=8)
    assert a.solver.work.batched_nodes == b.solver.work.batched_nodes and np.array_equal(a.U, b.U)


@pytest.mark.gpu
def test_power_converter_example_engine(capsys):
    _check_mpc_rows(_mpc(capsys, ["--horizons", "1,3", "--steps", "40", "--speculation", "64"]))


@pytest.mark.gpu
def test_speculation_changes_nothing_engine():
    r0, w0 = _bnb(0)
    r1, w1 = _bnb(64)
    # a launch with look-ahead tiles its nodes differently (other tile widths and tile mates): same decisions and
    # iteration counts, iterates to the parity tolerance of tests/test_gpu_parity.py
    assert w0.decisions == w1.decisions and w0.iter_num == w1.iter_num and w0.osqp_iter == w1.osqp_iter
    assert abs(r0.upper_glob - r1.upper_glob) <= 1e-9 * (1 + abs(r0.upper_glob))
    assert np.abs(r0.x - r1.x).max() <= 1e-9 * (1 + np.abs(r0.x).max())
    assert w1.spec_hits > 0 and w1.batches * 2 < w0.batches


@pytest.mark.gpu
def test_replay_maxiter_example_engine(capsys):
    import replay_maxiter
    replay_maxiter.main([])
    assert "49 problems in one launch" in capsys.readouterr().out


def test_solve_many_with_lookahead_cpu(cpu_engine):
    """Lock-step over several MIQPs with per-instance look-ahead: same B&B per instance, fewer launches."""
    import miosqp_b200
    from miosqp_b200 import problems

    def build(spec):
        out = []
        for seed in (3, 4, 5):
            pr = problems.random_miqp(40, 40, 20, 0.7, seed=seed)[0]
            s = miosqp_b200.MIOSQP()
            s.setup(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], pr['i_l'], pr['i_u'],
                    dict(problems.RANDOM_MIQP_SETTINGS, speculation=spec), dict(problems.RANDOM_MIQP_QP_SETTINGS))
            out.append(s)
        return out
    a, b = build(0), build(16)
    ra, rb = miosqp_b200.solve_many(a), miosqp_b200.solve_many(b)
    for sa, sb, xa, xb in zip(a, b, ra, rb):
        assert sa.work.decisions == sb.work.decisions and sa.work.osqp_iter == sb.work.osqp_iter
        assert xa.status == xb.status and xa.upper_glob == xb.upper_glob and np.array_equal(xa.x, xb.x)
        assert sb.work.spec_hits > 0
    assert max(s.work.batches for s in b) * 2 < max(s.work.batches for s in a)


def test_frontier_split_example_two_ranks_gloo(tmp_path):
    """examples/frontier_split.py under torch.distributed.run with 2 ranks (gloo, stand-in engine): the replicated replays
    agree, rank 0 reports the same B&B as the single-process run, and each rank solved about half of every batch."""
    import subprocess
    args = ["--vars", "40", "--rows", "40", "--ints", "20", "--density", "0.7", "--seed", "3", "--speculation", "16"]
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "_frontier_split_worker.py")] + args + ["--dist-backend", "gloo"]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1                                     # rank 0 only
    two = json.loads(lines[0])
    one = subprocess.run([sys.executable, os.path.join(HERE, "_frontier_split_worker.py")] + args, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=240, env=env)
    assert one.returncode == 0, one.stderr[-2000:]
    single = json.loads([ln for ln in one.stdout.splitlines() if ln.startswith("{")][0])
    for key in ("status", "upper_glob", "nodes", "admm_iters", "launches", "solved_nodes"):
        assert two[key] == single[key], key
    assert "2 GPU(s)" in two["workload"]
