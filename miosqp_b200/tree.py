"""
Branch-and-bound tree state with a BATCHED frontier.

The reference explores one node per iteration and solves it on the spot
(/root/reference/miosqp/solver.py:85-123 -> node.py:96-143).  Here every open leaf whose relaxation has not
been solved yet is flattened into one contiguous node batch and solved by ONE launch of the CUDA engine
(`Workspace.solve_pending`); the reference's sequential logic is then REPLAYED on the host over the cached
results, so the sequence of (chosen leaf, branching variable, incumbent updates, pruning) is the one the
reference would produce -- including its quirks, each marked "reference quirk" below.

Cache isolation: until the replay consumes a leaf, its visible fields stay as at creation (lower = parent's
bound, x/y = warm start, status = OSQP_UNSOLVED, num_iter = 0), because the reference's prune(), lower_glob and
leaf-selection rules read those fields of UNSOLVED leaves (workspace.py:145,278-280,334).
"""
from __future__ import print_function

from time import perf_counter

import numpy as np

from . import engine
from .constants import (MI_UNSOLVED, MI_SOLVED, MI_PRIMAL_INFEASIBLE, MI_DUAL_INFEASIBLE,
                        MI_MAX_ITER_FEASIBLE, MI_MAX_ITER_UNSOLVED)

OSQP_SOLVED = engine.CONSTANTS['OSQP_SOLVED']
OSQP_MAX_ITER_REACHED = engine.CONSTANTS['OSQP_MAX_ITER_REACHED']
OSQP_PRIMAL_INFEASIBLE = engine.CONSTANTS['OSQP_PRIMAL_INFEASIBLE']
OSQP_DUAL_INFEASIBLE = engine.CONSTANTS['OSQP_DUAL_INFEASIBLE']
OSQP_UNSOLVED = engine.CONSTANTS['OSQP_UNSOLVED']


class Node(object):
    """One B&B node = one QP relaxation (same attributes as the reference's Node, node.py:5-94)."""

    def __init__(self, data, l, u, solver, depth=0, lower=None, x0=None, y0=None):
        self.data, self.solver = data, solver
        self.l, self.u = l, u
        self.depth = depth
        self.lower = -np.inf if lower is None else lower
        self.frac_idx = None
        self.intinf = None
        self.num_iter = 0
        self.osqp_solve_time = 0
        self.x = np.zeros(data.n) if x0 is None else x0
        self.y = np.zeros(data.m + data.n_int) if y0 is None else y0
        self.status = OSQP_UNSOLVED
        self.nextvar_idx = None
        self.constr_idx = None
        self.cached = None          # (status, iters, seconds, x, y) of a batch solve not yet consumed
        self.shadow = None          # speculation: (left, right) children solved ahead of the replay, () if none can exist

    def solve(self):
        """Consume the cached batch result (solving this node alone first if nobody batched it)."""
        if self.cached is None:
            t0 = perf_counter()
            r = self.solver.solve_batch(self.l[None, :], self.u[None, :], self.x[None, :], self.y[None, :])
            self.cached = (int(r.status[0]), int(r.iters[0]), perf_counter() - t0, r.x[0], r.y[0])
        self.status, self.num_iter, self.osqp_solve_time, self.x, self.y = self.cached
        self.cached = None
        if self.status == OSQP_SOLVED or self.status == OSQP_MAX_ITER_REACHED:
            # integer entries exactly inside this node's bounds, then the bound from the clipped point (node.py:128-143)
            k, idx = self.data.n_int, self.data.i_idx
            self.x[idx] = np.minimum(np.maximum(self.x[idx], self.l[-k:]), self.u[-k:])
            self.lower = self.data.compute_obj_val(self.x)


class Workspace(object):
    """Frontier (`leaves`), incumbent and counters of one MIQP; one device-resident factor per workspace."""

    def __init__(self, data, settings, qp_settings=None, solver=None):
        self.data = data
        self.settings = settings
        self.qp_settings = {} if qp_settings is None else qp_settings
        # factor once: P and A never change afterwards (workspace.py:63-68); `solver`: already set up (miqp.setup_many)
        self.solver = solver if solver is not None else engine.BatchedQP().setup(
            data.P, data.q, data.A, data.l, data.u, i_idx=data.i_idx, **self.qp_settings)
        self.first_run = 1
        self.setup_time = self.solve_time = self.run_time = 0.
        self.batches = 0            # kernel launches issued for this workspace
        self.batched_nodes = 0      # nodes solved in them (>= consumed nodes: speculation)
        self.spec_nodes = 0         # of those, nodes solved ahead of the replay (settings['speculation'])
        self.spec_hits = 0          # speculative results the replay went on to consume
        self.reset()

    def reset(self):
        """Fresh tree on the current (q, l, u): what setup and update_vectors leave behind (solver.py:187-205)."""
        self.leaves = [Node(self.data, self.data.l, self.data.u, self.solver)]
        self.iter_num = 1           # reference quirk: counts from 1, so osqp_iter_avg divides by nodes + 1
        self.osqp_solve_time = 0.
        self.osqp_iter = 0
        self.osqp_iter_avg = 0
        self.lower_glob = -np.inf
        self.upper_glob = np.inf
        self.status = MI_UNSOLVED
        self.x = np.empty(self.data.n)
        self.decisions = []         # (constr_idx, nextvar_idx) per branching, for the parity tests

    # ------------------------------------------------------------------ batching
    def pending(self):
        """Open leaves whose relaxation has not been submitted to the engine yet."""
        return [leaf for leaf in self.leaves if leaf.cached is None and leaf.status == OSQP_UNSOLVED]

    @staticmethod
    def absorb(nodes, xs, ys, scalars, seconds):
        """Attach the results of one launch to their nodes; the launch time is shared by iteration count."""
        total = float(max(1, int(np.sum(scalars.iters))))
        for k, node in enumerate(nodes):
            share = seconds * float(scalars.iters[k]) / total
            node.cached = (int(scalars.status[k]), int(scalars.iters[k]), share, xs[k], ys[k])

    def solve_pending(self, dist_ctx=None):
        """One launch over every unsolved open leaf.  With dist_ctx (sharding.DistCtx, or a (rank, world, group[, device])
        tuple) the batch is dealt round-robin to the ranks, each solves its share on its own GPU, and one
        all_gather_into_tensor returns all results."""
        nodes = self.pending()
        if not nodes:
            return 0
        budget = int(self.settings.get('speculation', 0) or 0)
        if budget > 0:
            ahead = self.speculate(budget)
            self.spec_nodes += len(ahead)
            nodes = nodes + ahead
        t0 = perf_counter()
        if dist_ctx is None:
            mine = list(range(len(nodes)))
        else:
            from . import sharding
            dist_ctx = sharding.DistCtx.of(dist_ctx)
            mine = sharding.split_nodes(len(nodes), dist_ctx.rank, dist_ctx.world)
        local = {}
        if mine:
            sub = [nodes[k] for k in mine]
            xs, ys, sc = engine.solve_multi([self.solver] * len(sub), [nd.l for nd in sub], [nd.u for nd in sub],
                                            [nd.x for nd in sub], [nd.y for nd in sub])
            dt = perf_counter() - t0
            total = float(max(1, int(np.sum(sc.iters))))
            for j, k in enumerate(mine):
                local[k] = (int(sc.status[j]), int(sc.iters[j]), dt * float(sc.iters[j]) / total, xs[j], ys[j])
        if dist_ctx is not None:
            merged = sharding.exchange_node_results(local, len(nodes), self.data.n, self.data.m + self.data.n_int, dist_ctx)
        else:
            merged = [local[k] for k in range(len(nodes))]
        for node, res in zip(nodes, merged):
            node.cached = res
        self.batches += 1
        self.batched_nodes += len(nodes)
        return len(nodes)

    # ------------------------------------------------------------------ speculation (host-side look-ahead)
    def prospect(self, node):
        """The two children branch() WOULD create from `node` once the replay consumes its cached result, as
        unsolved shadow nodes (not in `leaves`); () when it cannot branch.  This is node.py:128-143 followed by
        workspace.py:245-264,205-230,157-203 evaluated on a copy: it depends on the node's own result only, never on
        the incumbent, which is what makes solving the children early legal.  Nothing visible changes."""
        st, _, _, x, y = node.cached
        if st != OSQP_SOLVED and st != OSQP_MAX_ITER_REACHED:
            return ()
        k, idx = self.data.n_int, self.data.i_idx
        xc = np.copy(x)
        xc[idx] = np.minimum(np.maximum(xc[idx], node.l[-k:]), node.u[-k:])
        x_int = xc[idx]
        dist = abs(x_int - np.round(x_int))
        frac = np.where(dist > self.settings['eps_int_feas'])[0]
        if len(frac) == 0:
            return ()
        nextvar = frac[int(np.argmax(dist[frac]))]
        row, var = self.data.m + nextvar, idx[nextvar]
        lower = self.data.compute_obj_val(xc)
        if lower > self.upper_glob:
            return ()                                   # the replay will drop it (workspace.py:299-300)
        kids = []
        for side in (0, 1):
            l, u = np.copy(node.l), np.copy(node.u)
            if side == 0:
                u[row] = np.floor(xc[var])
            else:
                l[row] = np.ceil(xc[var])
            if l[row] > u[row]:
                return ()                               # the engine would reject the whole launch; leave it to the replay
            kids.append(Node(self.data, l, u, self.solver, depth=node.depth + 1, lower=lower, x0=xc, y0=y))
        return tuple(kids)

    def speculate(self, budget):
        """Up to `budget` shadow nodes for this launch: children of solved-but-unconsumed nodes (open leaves and
        earlier shadows), the ones the exploration rule would reach first going first."""
        cands, stack = [], list(self.leaves)
        while stack:
            nd = stack.pop()
            if nd.cached is None:
                continue
            if nd.shadow is None:
                cands.append(nd)
            else:
                stack.extend(nd.shadow)
        rule = self.settings['tree_explor_rule']
        if rule == 0 or np.isinf(self.upper_glob):
            cands.sort(key=lambda nd: -nd.depth)
        else:
            cands.sort(key=lambda nd: -nd.lower)        # visible lower = the bound the "best bound" rule compares
        ahead = []
        for nd in cands:
            if len(ahead) + 2 > budget:
                break
            nd.shadow = self.prospect(nd)
            ahead.extend(nd.shadow)
        return ahead

    # ------------------------------------------------------------------ reference logic, replayed
    def set_x0(self, x0):
        root = self.leaves[0]
        if self.satisfies_lin_constraints(x0, root.l, root.u) and self.is_int_feas(x0, root):
            self.x = x0
            self.upper_glob = self.data.compute_obj_val(x0)
        else:
            print('Invalid initial solution!\n')
            self.upper_glob = np.inf
            self.x = np.empty(self.data.n)

    def can_continue(self):
        return len(self.leaves) > 0 and self.iter_num < self.settings['max_iter_bb']

    def choose_leaf(self, tree_explor_rule):
        if tree_explor_rule == 0 or (tree_explor_rule == 1 and np.isinf(self.upper_glob)):
            pick = int(np.argmax([leaf.depth for leaf in self.leaves]))      # depth first, first maximum wins
        elif tree_explor_rule == 1:
            # reference quirk: "best bound" phase takes the LARGEST lower bound (workspace.py:145)
            pick = int(np.argmax([leaf.lower for leaf in self.leaves]))
        else:
            raise ValueError('Tree exploring strategy not recognized')
        return self.leaves.pop(pick)

    def _child(self, leaf, l, u, side):
        # children warm-start from (and share) the parent's solution arrays (workspace.py:174-176)
        child = Node(self.data, l, u, self.solver, depth=leaf.depth + 1, lower=leaf.lower, x0=leaf.x, y0=leaf.y)
        child.parent_iters = leaf.num_iter      # scheduling hint only (longest-first submission)
        if leaf.shadow:
            # a result solved ahead of the replay is adopted only if it was computed from exactly these inputs
            sh = leaf.shadow[side]
            if sh.cached is not None and np.array_equal(sh.l, l) and np.array_equal(sh.u, u) \
                    and np.array_equal(sh.x, leaf.x) and np.array_equal(sh.y, leaf.y):
                child.cached, child.shadow = sh.cached, sh.shadow
                self.spec_hits += 1
        self.leaves.append(child)

    def add_left(self, leaf):
        l, u = np.copy(leaf.l), np.copy(leaf.u)
        u[leaf.constr_idx] = np.floor(leaf.x[leaf.nextvar_idx])
        self._child(leaf, l, u, 0)

    def add_right(self, leaf):
        l, u = np.copy(leaf.l), np.copy(leaf.u)
        l[leaf.constr_idx] = np.ceil(leaf.x[leaf.nextvar_idx])
        self._child(leaf, l, u, 1)

    def pick_nextvar(self, leaf):
        if self.settings['branching_rule'] != 0:
            raise ValueError('No variable selection rule recognized!')
        x_frac = leaf.x[self.data.i_idx[leaf.frac_idx]]
        nextvar = leaf.frac_idx[int(np.argmax(abs(x_frac - np.round(x_frac))))]   # most fractional, first wins
        leaf.constr_idx = self.data.m + nextvar
        leaf.nextvar_idx = self.data.i_idx[nextvar]

    def satisfies_lin_constraints(self, x, l, u):
        z = self.data.A_op.dot(x)
        eps = self.qp_settings['eps_abs']       # reference quirk: KeyError when qp_settings lacks eps_abs
        return not (np.any(z < l - eps) or np.any(z > u + eps))

    def is_int_feas(self, x, leaf):
        x_int = x[self.data.i_idx]
        frac = abs(x_int - np.round(x_int)) > self.settings['eps_int_feas']
        leaf.frac_idx = np.where(frac)[0].tolist()
        leaf.intinf = np.sum(frac)
        return not leaf.intinf > 0

    def get_integer_solution(self, x):
        x_int = np.copy(x)
        x_int[self.data.i_idx] = np.round(x[self.data.i_idx])      # half-to-even, as np.round
        return x_int

    def prune(self):
        # reference quirk (workspace.py:274-280): the list is mutated while being iterated, so the element that
        # slides into a removed slot is NOT examined in this pass.  Reproduced with an explicit cursor.
        k = 0
        while k < len(self.leaves):
            if self.leaves[k].lower > self.upper_glob:
                del self.leaves[k]
            k += 1

    def bound_and_branch(self, leaf):
        self.osqp_iter += leaf.num_iter
        self.osqp_solve_time += leaf.osqp_solve_time
        if leaf.status == OSQP_PRIMAL_INFEASIBLE or leaf.status == OSQP_DUAL_INFEASIBLE:
            return
        if leaf.lower > self.upper_glob:
            return
        if self.is_int_feas(leaf.x, leaf):
            self.x = leaf.x
            self.upper_glob = leaf.lower
            self.prune()
            return
        # rounding heuristic against the ROOT bounds (workspace.py:321-328)
        x_int = self.get_integer_solution(leaf.x)
        if self.satisfies_lin_constraints(x_int, self.data.l, self.data.u):
            obj_int = self.data.compute_obj_val(x_int)
            if obj_int < self.upper_glob:
                self.upper_glob = obj_int
                self.x = x_int
                self.prune()
        self.branch(leaf)
        self.lower_glob = min([lf.lower for lf in self.leaves])

    def branch(self, leaf):
        self.pick_nextvar(leaf)
        self.decisions.append((int(leaf.constr_idx), int(leaf.nextvar_idx)))
        self.add_left(leaf)
        self.add_right(leaf)

    def get_return_status(self):
        finished = self.iter_num < self.settings['max_iter_bb']
        if self.upper_glob != np.inf:
            self.status = MI_SOLVED if finished else MI_MAX_ITER_FEASIBLE
        elif self.upper_glob >= 0:
            self.status = MI_PRIMAL_INFEASIBLE if finished else MI_MAX_ITER_UNSOLVED
        else:
            self.status = MI_DUAL_INFEASIBLE

    def get_return_solution(self):
        if self.status == MI_SOLVED or self.status == MI_MAX_ITER_FEASIBLE:
            self.x[self.data.i_idx] = np.round(self.x[self.data.i_idx])

    # ------------------------------------------------------------------ progress table (workspace.py:386-433)
    def print_headline(self):
        print("     Nodes      |           Current Node        |             Objective Bounds             |   Cur Node")
        print("Explr\tUnexplr\t|      Obj\tDepth\tIntInf  |    Lower\t   Upper\t    Gap    |     Iter")

    def print_progress(self, leaf):
        gap = "    --- " if self.upper_glob == np.inf else \
            "%8.2f%%" % ((self.upper_glob - self.lower_glob) / abs(self.lower_glob) * 100)
        infeasible = leaf.status == OSQP_PRIMAL_INFEASIBLE or leaf.status == OSQP_DUAL_INFEASIBLE
        obj = np.inf if infeasible else leaf.lower
        intinf = "  ---" if leaf.intinf is None else "%5d" % leaf.intinf
        print("%4d\t%4d\t  %10.2e\t%4d\t%s\t  %10.2e\t%10.2e\t%s\t%5d" %
              (self.iter_num, len(self.leaves), obj, leaf.depth, intinf, self.lower_glob, self.upper_glob, gap,
               leaf.num_iter), end='')
        print("!" if leaf.status == OSQP_MAX_ITER_REACHED else "")

    def print_footer(self):
        print("\n")
        print("Status: %s" % self.status)
        if self.status == MI_SOLVED:
            print("Objective bound: %6.3e" % self.upper_glob)
        print("Total number of OSQP iterations: %d" % self.osqp_iter)
