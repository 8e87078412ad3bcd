"""
MIOSQP -- the reference's public API (/root/reference/miosqp/solver.py:32-212) on the batched CUDA engine.

    m = MIOSQP(); m.setup(P, q, A, l, u, i_idx, i_l, i_u, settings, qp_settings)
    res = m.solve();  m.update_vectors(q=..., l=..., u=...);  m.set_x0(x0)

`solve()` alternates (1) one engine launch over every open leaf that has no result yet and (2) a host replay of
the reference's sequential loop over the cached results (tree.py).  `solve_many` drives several MIQPs in
lock-step so that ONE launch per B&B step covers the frontiers of all of them (BASELINE config 2: 100 instances).
"""
from __future__ import print_function

from time import perf_counter, time

import numpy as np

from . import engine
from .problem_data import Data
from .results import Results
from .tree import Workspace


class MIOSQP(object):
    def __init__(self):
        self.data = None
        self.work = None

    def setup(self, P, q, A, l, u, i_idx, i_l, i_u, settings, qp_settings):
        start = time()
        if i_l is None:
            i_l = -np.inf * np.ones(len(i_idx))
        if i_u is None:
            i_u = np.inf * np.ones(len(i_idx))
        self.data = Data(P, q, A, l, u, i_idx, i_l, i_u)
        self.work = Workspace(self.data, settings, qp_settings)
        self.work.setup_time = time() - start

    # -- one B&B replay step over cached results; returns False when the chosen leaf still needs the engine
    def _replay(self):
        work = self.work
        while work.can_continue():
            if work.pending():
                return True            # unsolved leaves in the frontier: batch them before choosing
            leaf = work.choose_leaf(work.settings['tree_explor_rule'])
            leaf.solve()
            work.bound_and_branch(leaf)
            if work.settings['verbose'] and work.iter_num % work.settings['print_interval'] == 0:
                work.print_progress(leaf)
            work.iter_num += 1
        return False

    def _begin(self):
        self._t0 = time()
        if self.work.settings['verbose']:
            self.work.print_headline()

    def _finish(self):
        work = self.work
        work.osqp_iter_avg = work.osqp_iter / work.iter_num
        work.get_return_status()
        work.get_return_solution()
        if work.settings['verbose']:
            work.print_footer()
        work.solve_time = time() - self._t0
        if work.first_run:
            work.first_run = 0
            work.run_time = work.setup_time + work.solve_time
        else:
            work.run_time = work.solve_time
        if work.settings['verbose']:
            print("Elapsed time: %.4es" % work.run_time)
        return Results(work.x, work.upper_glob, work.run_time, work.status, work.osqp_solve_time, work.osqp_iter_avg)

    def _solve_native(self):
        """settings['replay'] = 'native': the same loop in C++ (csrc/bqp_bnb.cpp, bqp_bnb_solve) -- no interpreter time
        per node; one call in, the reference's Results out.  Leaves `work` as the Python replay would."""
        work = self.work
        if not work.leaves and work.iter_num > 1:
            # the tree is finished (a second solve() without update_vectors): the reference's loop -- and the Python replay --
            # find no leaf and return what they have; do not rebuild the tree from its root
            return self._finish()
        x, r, decisions = engine.bnb_solve(work.solver, work.data, work.settings, work.qp_settings['eps_abs'],
                                           work.x if np.isfinite(work.upper_glob) else None, work.upper_glob)
        return self._absorb_native(x, r, decisions)

    def _absorb_native(self, x, r, decisions):
        from .constants import (MI_UNSOLVED, MI_SOLVED, MI_PRIMAL_INFEASIBLE, MI_DUAL_INFEASIBLE,
                                MI_MAX_ITER_FEASIBLE, MI_MAX_ITER_UNSOLVED)
        work = self.work
        work.x, work.upper_glob, work.lower_glob = x, r["upper_glob"], r["lower_glob"]
        work.iter_num, work.osqp_iter, work.osqp_solve_time = r["iter_num"], r["osqp_iter"], r["osqp_solve_time"]
        work.batches += r["batches"]; work.batched_nodes += r["batched_nodes"]
        work.spec_nodes += r["spec_nodes"]; work.spec_hits += r["spec_hits"]
        work.decisions = decisions
        work.leaves = []                       # the native tree is not materialised on the Python side
        work.osqp_iter_avg = work.osqp_iter / work.iter_num
        work.status = (MI_UNSOLVED, MI_SOLVED, MI_PRIMAL_INFEASIBLE, MI_DUAL_INFEASIBLE, MI_MAX_ITER_FEASIBLE,
                       MI_MAX_ITER_UNSOLVED)[r["status"]]
        if work.settings['verbose']:
            work.print_footer()
        work.solve_time = time() - self._t0
        if work.first_run:
            work.first_run = 0
            work.run_time = work.setup_time + work.solve_time
        else:
            work.run_time = work.solve_time
        return Results(work.x, work.upper_glob, work.run_time, work.status, work.osqp_solve_time, work.osqp_iter_avg)

    def solve(self, dist_ctx=None):
        """dist_ctx = sharding.DistCtx or a (rank, world, group[, device]) tuple: split every frontier batch across the ranks'
        GPUs (config 4); all ranks replay the same tree and agree on the incumbent with one all-reduce(MIN) per B&B step."""
        self._begin()
        if self.work.settings.get('replay') == 'native' and dist_ctx is None:
            return self._solve_native()
        if dist_ctx is not None:
            from . import sharding
            dist_ctx = sharding.DistCtx.of(dist_ctx)
        while self._replay():
            self.work.solve_pending(dist_ctx)
            if dist_ctx is not None:
                best, same = sharding.agree_incumbent(self.work.upper_glob, ctx=dist_ctx)
                if not same:
                    raise RuntimeError("replicated B&B replays diverged: incumbent %r vs global %r" % (self.work.upper_glob, best))
        return self._finish()

    def update_vectors(self, q=None, l=None, u=None):
        work = self.work
        work.data.update_vectors(q, l, u)
        if q is not None:
            work.solver.update_q(q)          # factor kept; only the scaled linear cost is re-uploaded
        work.reset()
        work.solve_time = 0.
        work.run_time = 0.

    def set_x0(self, x0):
        self.work.set_x0(x0)


def setup_many(problems, settings, qp_settings, threads=0):
    """`MIOSQP().setup(...)` for many MIQPs sharing their settings: the factorisations run on all host threads
    (engine.setup_many / bqp_setup_many).  problems: dicts with P, q, A, l, u, i_idx, i_l, i_u (problems.random_miqp)."""
    start = time()
    datas = []
    for pr in problems:
        i_l = pr.get('i_l'); i_u = pr.get('i_u')
        i_l = -np.inf * np.ones(len(pr['i_idx'])) if i_l is None else i_l
        i_u = np.inf * np.ones(len(pr['i_idx'])) if i_u is None else i_u
        datas.append(Data(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], i_l, i_u))
    qp = dict(qp_settings or {})
    qps = engine.setup_many([(d.P, d.q, d.A, d.l, d.u, d.i_idx) for d in datas], threads=threads, **qp)
    out = []
    per = (time() - start) / max(1, len(datas))
    for d, s in zip(datas, qps):
        m = MIOSQP()
        m.data = d
        m.work = Workspace(d, dict(settings), dict(qp), solver=s)
        m.work.setup_time = per
        out.append(m)
    return out


def wrap_many(problems, qps, settings, qp_settings):
    """MIOSQP objects around problems whose factors are already resident (`qps`: engine.BatchedQP, same order): what
    setup_many does after the factorisation."""
    out = []
    for pr, s in zip(problems, qps):
        i_l = pr.get('i_l'); i_u = pr.get('i_u')
        i_l = -np.inf * np.ones(len(pr['i_idx'])) if i_l is None else i_l
        i_u = np.inf * np.ones(len(pr['i_idx'])) if i_u is None else i_u
        m = MIOSQP()
        m.data = Data(pr['P'], pr['q'], pr['A'], pr['l'], pr['u'], pr['i_idx'], i_l, i_u)
        m.work = Workspace(m.data, dict(settings), dict(qp_settings or {}), solver=s)
        out.append(m)
    return out


def solve_many(solvers, on_batch=None, async_threads=None, rolling=False):
    """Solve several set-up MIOSQP objects together.  Lock-step (default): every step flattens the unsolved leaves of ALL
    frontiers into one node batch (one kernel launch).  async_threads is not None (native replay only; 0 = automatic):
    every MIQP runs its own replay/launch loop on its own CUDA stream (bqp_bnb_solve_async), so none waits for another
    one's slowest leaf.  rolling=True (native replay only): all trees share one engine session; after every round of 100 ADMM
    iterations the trees whose leaves have terminated are replayed and their children join the next launch
    (bqp_bnb_solve_rolling).  Either way each instance's result is identical to its own `solve()`.
    `on_batch(n_nodes, seconds)` is called after every launch of the Python lock-step loop (benchmarks)."""
    for s in solvers:
        s._begin()
    if solvers and all(s.work.settings.get('replay') == 'native' for s in solvers):
        # the loop itself in C++ (bqp_bnb_solve_many / bqp_bnb_solve_async): no interpreter time per node
        works = [s.work for s in solvers]
        many_fn = engine.native_solve_many_fn([w.solver for w in works])
        outs = engine.bnb_solve_many([w.solver for w in works], [w.data for w in works], [w.settings for w in works],
                                     [w.qp_settings['eps_abs'] for w in works],
                                     [(w.x if np.isfinite(w.upper_glob) else None) for w in works],
                                     [w.upper_glob for w in works], many_fn=many_fn, async_threads=async_threads, rolling=rolling)
        return [s._absorb_native(x, r, d) for s, (x, r, d) in zip(solvers, outs)]
    active = list(solvers)
    while active:
        active = [s for s in active if s._replay()]
        nodes, owners = [], []
        for s in active:
            w = s.work
            mine = w.pending()
            budget = int(w.settings.get('speculation', 0) or 0)
            if mine and budget > 0:              # look-ahead nodes of this instance ride along (tree.py speculate)
                ahead = w.speculate(budget)
                w.spec_nodes += len(ahead)
                mine = mine + ahead
            for nd in mine:
                nodes.append(nd); owners.append(w)
        if not nodes:
            continue
        # longest-first submission: children of slow-converging parents go first (tile order = submission order)
        order = sorted(range(len(nodes)), key=lambda k: -getattr(nodes[k], "parent_iters", 0))
        nodes = [nodes[k] for k in order]; owners = [owners[k] for k in order]
        t0 = perf_counter()
        xs, ys, sc = engine.solve_multi([w.solver for w in owners], [nd.l for nd in nodes], [nd.u for nd in nodes],
                                        [nd.x for nd in nodes], [nd.y for nd in nodes])
        dt = perf_counter() - t0
        Workspace.absorb(nodes, xs, ys, sc, dt)
        for w in set(owners):
            w.batches += 1
        for w in owners:
            w.batched_nodes += 1
        if on_batch is not None:
            on_batch(len(nodes), dt)
    return [s._finish() for s in solvers]
