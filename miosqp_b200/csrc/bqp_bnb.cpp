// bqp_bnb.cpp -- native branch-and-bound replay with look-ahead (bqp_bnb_solve, include/bqp.h).
//
// The reference runs its B&B loop in Python, one node per iteration (/root/reference/miosqp/solver.py:85-123,
// workspace.py:128-384, node.py:96-143).  miosqp_b200/tree.py replays that logic over batched results; at BASELINE
// config 3 (hundreds of tiny nodes per MPC step) its ~90 us of interpreter time per node is what remains once the
// look-ahead has removed the host round trips.  This file is the same replay, statement for
// statement, in C++: frontier, leaf selection, clip + objective, integer-feasibility test, rounding heuristic against
// the ROOT bounds, most-fractional branching, the reference's prune()/iter_num/"largest lower bound" quirks, and the
// look-ahead of tree.py (shadow children adopted only when computed from exactly the replay's inputs).
// It talks to the engine only through the public entry point bqp_solve_multi -- or through a caller-supplied solve
// function (CPU tests drive it with the oracle; nothing here links to oracle/).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/bqp.h"

namespace {

using Vec = std::vector<double>;
using VecP = std::shared_ptr<Vec>;
constexpr double kInf = std::numeric_limits<double>::infinity();

// BQP_BNB_TIMERS=1: where the wall time of the single-tree driver goes, summed over the process and printed at exit (stderr)
struct BnbTimers {
  bool on = std::getenv("BQP_BNB_TIMERS") != nullptr;
  double t_collect = 0, t_engine = 0, t_absorb = 0, t_advance = 0; long long launches = 0, nodes = 0;
  long long real_nodes = 0, max_real = 0, max_all = 0;      // per launch: iterations of its slowest open leaf / slowest node incl. look-ahead, summed
  ~BnbTimers() {
    if (on && launches)
      std::fprintf(stderr, "BNB %lld launches, %lld nodes (%lld open leaves): collect %.3f s, engine %.3f s, absorb %.3f s, replay %.3f s; "
                   "iterations of the slowest node per launch, summed: %lld over the open leaves, %lld incl. look-ahead\n", launches, nodes, real_nodes,
                   t_collect, t_engine, t_absorb, t_advance, max_real, max_all);
  }
};
static BnbTimers g_bt;
static inline double bt_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Node;
using NodeP = std::shared_ptr<Node>;

struct Node {
  Vec l, u;                 // extended bounds of this node
  VecP x, y;                // warm start before solve(), solution after (children share the parent's arrays)
  int depth = 0;
  double lower = -kInf;
  int status = BQP_UNSOLVED, num_iter = 0;
  double seconds = 0.0;
  // result of a launch not yet consumed by the replay (cache isolation: nothing above changes until solve())
  bool has_cached = false;
  int c_status = BQP_UNSOLVED, c_iters = 0; double c_seconds = 0.0; VecP cx, cy;
  // look-ahead: 0 = not examined, 1 = cannot branch, 2 = sh[0], sh[1] exist
  int shadow_state = 0;
  NodeP sh[2];
  std::vector<int> frac_idx;
  int constr_idx = -1, nextvar_idx = -1;
  int parent_iters = 0;     // scheduling hint only: longest-first submission in the lock-step driver
  // the look-ahead computed 1/2 x'Px + q'x of the cached result at the clipped point: solve() evaluates the same function of the
  // same vector (same code, same order of operations), so the replay takes the value instead of a second pass over P
  bool has_spec_lower = false; double spec_lower = 0.0;
};

struct Csc { int rows = 0, cols = 0; const int *p = nullptr, *i = nullptr; const double *x = nullptr; };

struct Tree {
  int n = 0, m = 0, n_int = 0, m_ext = 0;       // m = ORIGINAL rows, m_ext = m + n_int
  Csc P, A;
  const double *q = nullptr; const int *i_idx = nullptr;
  Vec l_root, u_root;
  bqp_bnb_settings s;
  std::vector<NodeP> leaves;
  int iter_num = 1;                              // reference quirk: counts from 1 (solver.py:130 divides by nodes + 1)
  long long osqp_iter = 0; double osqp_solve_time = 0.0;
  double upper_glob = kInf, lower_glob = -kInf;
  Vec x_best;
  std::vector<int> decisions;
  int batches = 0; long long batched_nodes = 0, spec_nodes = 0, spec_hits = 0;
  bqp_handle h = nullptr; bqp_solve_fn fn = nullptr; void *ctx = nullptr;
  bqp_ctx ectx = nullptr;                        // engine context (own stream) of the thread driving this tree, if any
  Vec tmp;

  // y = M x in CSC column order: the loop scipy's csc_matvec runs.  Large matrices (config 4: P is a dense 2000 x 2000, 8 ms
  // per product on one core, more than the GPU needs for the whole relaxation) are split by ROW RANGE over host threads:
  // every y[r] still receives its terms in ascending column order, so the result is bit-identical to the serial loop.
  // Needs ascending row indices inside every column (checked once per matrix).
  static bool sorted_rows(const Csc &M) {
    for (int j = 0; j < M.cols; j++)
      for (int k = M.p[j] + 1; k < M.p[j + 1]; k++) if (M.i[k] <= M.i[k - 1]) return false;
    return true;
  }
  static void matvec_rows(const Csc &M, const double *x, double *y, int r0, int r1) {
    for (int r = r0; r < r1; r++) y[r] = 0.0;
    for (int j = 0; j < M.cols; j++) {
      const int *b = M.i + M.p[j], *e = M.i + M.p[j + 1];
      const int *lo = (r0 == 0) ? b : std::lower_bound(b, e, r0);
      const double xj = x[j];
      for (const int *k = lo; k < e && *k < r1; k++) y[*k] += M.x[k - M.i] * xj;
    }
  }
  int par_P = -1, par_A = -1;                    // -1 not examined, 0 serial, > 0 host threads
  bool dense_P = false;                          // P stores every entry (config 3): mat-vec without the index stream
  static int par_threads(const Csc &M) {
    const long long nnz = M.p[M.cols];
    if (nnz < 400000 || !sorted_rows(M)) return 0;
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(2u, std::min(16u, hw ? hw : 2u));
  }
  // the same for a matrix that stores every entry (column j = M.x + j * rows): no index stream, the rows of the range vectorise
  static void matvec_rows_dense(const Csc &M, const double *x, double *y, int r0, int r1) {
    for (int r = r0; r < r1; r++) y[r] = 0.0;
    const size_t rows = (size_t)M.rows;
    for (int j = 0; j < M.cols; j++) {
      const double xj = x[j];
      const double *__restrict__ c = M.x + (size_t)j * rows;
      for (int r = r0; r < r1; r++) y[r] += c[r] * xj;
    }
  }
  static void matvec(const Csc &M, const double *x, double *y, int threads = 0, bool dense = false) {
    if (threads > 1) {
      std::vector<std::thread> th;
      auto fn = dense ? matvec_rows_dense : matvec_rows;
      for (int t = 1; t < threads; t++)
        th.emplace_back(fn, std::cref(M), x, y, (int)((long long)M.rows * t / threads), (int)((long long)M.rows * (t + 1) / threads));
      fn(M, x, y, 0, (int)((long long)M.rows / threads));
      for (auto &t : th) t.join();
      return;
    }
    for (int r = 0; r < M.rows; r++) y[r] = 0.0;
    if (dense) {
      // every column holds all rows in ascending order (checked by the caller: config 3's P): the same additions in the same
      // order as the indexed loop below, without the index stream -- the compiler vectorises over the rows
      const int rows = M.rows;
      for (int j = 0; j < M.cols; j++) {
        const double xj = x[j];
        const double *__restrict__ c = M.x + (size_t)j * rows;
        for (int r = 0; r < rows; r++) y[r] += c[r] * xj;
      }
      return;
    }
    for (int j = 0; j < M.cols; j++) { const double xj = x[j]; for (int k = M.p[j]; k < M.p[j + 1]; k++) y[M.i[k]] += M.x[k] * xj; }
  }
  static bool is_dense(const Csc &M) {
    if ((long long)M.p[M.cols] != (long long)M.rows * M.cols) return false;
    for (int j = 0; j < M.cols; j++) {
      if (M.p[j] != j * M.rows) return false;
      for (int r = 0; r < M.rows; r++) if (M.i[M.p[j] + r] != r) return false;
    }
    return true;
  }
  double obj(const Vec &x) {                     // data.py:99-103
    tmp.resize(std::max(n, m_ext));
    if (par_P < 0) { par_P = par_threads(P); dense_P = is_dense(P); }
    matvec(P, x.data(), tmp.data(), par_P, dense_P);
    double a = 0.0, b = 0.0;
    for (int j = 0; j < n; j++) { a += x[j] * tmp[j]; b += q[j] * x[j]; }
    return .5 * a + b;
  }
  bool satisfies_lin(const Vec &x, const Vec &l, const Vec &u) {      // workspace.py:232-243
    tmp.resize(std::max(n, m_ext));
    if (par_A < 0) par_A = par_threads(A);
    matvec(A, x.data(), tmp.data(), par_A);
    for (int r = 0; r < m_ext; r++) if (tmp[r] < l[r] - s.eps_abs || tmp[r] > u[r] + s.eps_abs) return false;
    return true;
  }
  // fractional integer entries of x (workspace.py:245-264); returns true when there are none
  bool int_feas(const Vec &x, std::vector<int> &frac) const {
    frac.clear();
    for (int k = 0; k < n_int; k++) { const double v = x[i_idx[k]]; if (std::fabs(v - std::nearbyint(v)) > s.eps_int_feas) frac.push_back(k); }
    return frac.empty();
  }
  int most_fractional(const Vec &x, const std::vector<int> &frac) const {   // workspace.py:205-230, first maximum wins
    int best = frac[0]; double bd = -1.0;
    for (int k : frac) { const double v = x[i_idx[k]], d = std::fabs(v - std::nearbyint(v)); if (d > bd) { bd = d; best = k; } }
    return best;
  }
  void clip_int(Vec &x, const Vec &l, const Vec &u) const {                  // node.py:131-136
    for (int k = 0; k < n_int; k++) { double &v = x[i_idx[k]]; v = std::fmin(std::fmax(v, l[m + k]), u[m + k]); }
  }

  static bool unsolved(const Node &nd) { return !nd.has_cached && nd.status == BQP_UNSOLVED; }

  // ---- look-ahead (tree.py prospect / speculate)
  void prospect(Node &nd) {
    nd.shadow_state = 1;
    if (nd.c_status != BQP_SOLVED && nd.c_status != BQP_MAX_ITER_REACHED) return;
    auto xc = std::make_shared<Vec>(*nd.cx);
    clip_int(*xc, nd.l, nd.u);
    std::vector<int> frac;
    if (int_feas(*xc, frac)) return;
    const int nextvar = most_fractional(*xc, frac), row = m + nextvar, var = i_idx[nextvar];
    const double lower = obj(*xc);
    nd.has_spec_lower = true; nd.spec_lower = lower;
    if (lower > upper_glob) return;                           // the replay will drop it (workspace.py:299-300)
    NodeP kids[2];
    for (int side = 0; side < 2; side++) {
      auto c = std::make_shared<Node>();
      c->l = nd.l; c->u = nd.u;
      if (side == 0) c->u[row] = std::floor((*xc)[var]); else c->l[row] = std::ceil((*xc)[var]);
      if (c->l[row] > c->u[row]) return;                      // the engine would reject the whole launch
      c->x = xc; c->y = nd.cy; c->depth = nd.depth + 1; c->lower = lower; c->parent_iters = nd.c_iters;
      kids[side] = c;
    }
    nd.sh[0] = kids[0]; nd.sh[1] = kids[1]; nd.shadow_state = 2;
  }
  void speculate(int budget, std::vector<Node *> &batch) {
    std::vector<Node *> cands, stack;
    for (auto &lf : leaves) stack.push_back(lf.get());
    while (!stack.empty()) {
      Node *nd = stack.back(); stack.pop_back();
      if (!nd->has_cached) continue;
      if (nd->shadow_state == 0) cands.push_back(nd);
      else if (nd->shadow_state == 2) { stack.push_back(nd->sh[0].get()); stack.push_back(nd->sh[1].get()); }
    }
    const bool depth_first = s.tree_explor_rule == 0 || std::isinf(upper_glob);
    std::stable_sort(cands.begin(), cands.end(), [&](const Node *a, const Node *b) {
      return depth_first ? a->depth > b->depth : a->lower > b->lower; });
    int added = 0;
    for (Node *nd : cands) {
      if (added + 2 > budget) break;
      prospect(*nd);
      if (nd->shadow_state == 2) { batch.push_back(nd->sh[0].get()); batch.push_back(nd->sh[1].get()); added += 2; }
    }
    spec_nodes += added;
  }

  // ---- one launch over every unsolved open leaf (+ look-ahead)
  void collect(std::vector<Node *> &batch) {
    const size_t first = batch.size();
    for (auto &lf : leaves) if (unsolved(*lf)) batch.push_back(lf.get());
    if (batch.size() == first) return;
    if (s.speculation > 0) speculate(s.speculation, batch);
    for (size_t b = first; b < batch.size(); b++) { batch[b]->cx = std::make_shared<Vec>(n); batch[b]->cy = std::make_shared<Vec>(m_ext); }
  }
  static void absorb(Node &nd, int status, int iters, double seconds) {
    nd.has_spec_lower = false;
    nd.has_cached = true; nd.c_status = status; nd.c_iters = iters; nd.c_seconds = seconds;
  }
  int launch() {
    std::vector<Node *> batch;
    const double tc0 = g_bt.on ? bt_now() : 0.0;
    collect(batch);
    if (g_bt.on) g_bt.t_collect += bt_now() - tc0;
    if (batch.empty()) return BQP_OK;
    const int B = (int)batch.size();
    std::vector<const double *> pl(B), pu(B), px0(B), py0(B);
    std::vector<double *> px(B), py(B);
    std::vector<int> status(B), iters(B);
    for (int b = 0; b < B; b++) {
      Node &nd = *batch[b];
      pl[b] = nd.l.data(); pu[b] = nd.u.data(); px0[b] = nd.x->data(); py0[b] = nd.y->data();
      px[b] = nd.cx->data(); py[b] = nd.cy->data();
    }
    const auto t0 = std::chrono::steady_clock::now();
    int rc;
    if (fn) rc = fn(ctx, B, pl.data(), pu.data(), px0.data(), py0.data(), px.data(), py.data(), status.data(), iters.data());
    else {
      std::vector<bqp_handle> hs(B, h);
      bqp_node_out out; std::memset(&out, 0, sizeof(out));
      out.status = status.data(); out.iters = iters.data();
      rc = ectx ? bqp_ctx_solve_multi(ectx, B, hs.data(), pl.data(), pu.data(), px0.data(), py0.data(), px.data(), py.data(), &out)
                : bqp_solve_multi(B, hs.data(), pl.data(), pu.data(), px0.data(), py0.data(), px.data(), py.data(), &out);
    }
    if (rc) return rc;
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long long total = 0; for (int b = 0; b < B; b++) total += iters[b];
    if (total < 1) total = 1;
    const double ta0 = g_bt.on ? bt_now() : 0.0;
    for (int b = 0; b < B; b++) absorb(*batch[b], status[b], iters[b], dt * (double)iters[b] / (double)total);
    if (g_bt.on) {
      g_bt.t_engine += dt; g_bt.t_absorb += bt_now() - ta0; g_bt.launches++; g_bt.nodes += B;
      int nreal = 0, mr = 0, ma = 0;
      for (auto &lf : leaves) for (int b = 0; b < B; b++) if (batch[b] == lf.get()) { nreal++; mr = std::max(mr, iters[b]); }
      for (int b = 0; b < B; b++) ma = std::max(ma, iters[b]);
      g_bt.real_nodes += nreal; g_bt.max_real += mr; g_bt.max_all += ma;
    }
    batches++; batched_nodes += B;
    return BQP_OK;
  }

  // ---- reference logic, replayed
  int choose_leaf(NodeP &out) {                                           // workspace.py:128-155
    size_t pick = 0;
    if (s.tree_explor_rule == 0 || (s.tree_explor_rule == 1 && std::isinf(upper_glob))) {
      for (size_t k = 1; k < leaves.size(); k++) if (leaves[k]->depth > leaves[pick]->depth) pick = k;
    } else if (s.tree_explor_rule == 1) {
      // reference quirk: the "best bound" phase takes the LARGEST lower bound (workspace.py:145)
      for (size_t k = 1; k < leaves.size(); k++) if (leaves[k]->lower > leaves[pick]->lower) pick = k;
    } else return BQP_BNB_E_EXPLOR_RULE;
    out = leaves[pick];
    leaves.erase(leaves.begin() + (long)pick);
    return BQP_OK;
  }
  void solve_node(Node &nd) {                                             // node.py:96-143 over the cached result
    nd.status = nd.c_status; nd.num_iter = nd.c_iters; nd.seconds = nd.c_seconds; nd.x = nd.cx; nd.y = nd.cy;
    nd.has_cached = false;
    if (nd.status == BQP_SOLVED || nd.status == BQP_MAX_ITER_REACHED) {
      clip_int(*nd.x, nd.l, nd.u);
      nd.lower = nd.has_spec_lower ? nd.spec_lower : obj(*nd.x);
    }
  }
  void prune() {
    // reference quirk (workspace.py:274-280): the list is mutated while iterated, so the element that slides into a
    // removed slot is not examined in this pass
    for (size_t k = 0; k < leaves.size(); k++) if (leaves[k]->lower > upper_glob) leaves.erase(leaves.begin() + (long)k);
  }
  void add_child(Node &leaf, int side) {                                  // workspace.py:157-203
    auto c = std::make_shared<Node>();
    c->l = leaf.l; c->u = leaf.u;
    if (side == 0) c->u[leaf.constr_idx] = std::floor((*leaf.x)[leaf.nextvar_idx]);
    else c->l[leaf.constr_idx] = std::ceil((*leaf.x)[leaf.nextvar_idx]);
    c->x = leaf.x; c->y = leaf.y; c->depth = leaf.depth + 1; c->lower = leaf.lower; c->parent_iters = leaf.num_iter;
    if (leaf.shadow_state == 2) {
      const Node &sh = *leaf.sh[side];
      // adopt a result solved ahead of the replay only if it was computed from exactly these inputs
      if (sh.has_cached && sh.l == c->l && sh.u == c->u && *sh.x == *c->x && *sh.y == *c->y) {
        c->has_cached = true; c->c_status = sh.c_status; c->c_iters = sh.c_iters; c->c_seconds = sh.c_seconds; c->cx = sh.cx; c->cy = sh.cy;
        c->has_spec_lower = sh.has_spec_lower; c->spec_lower = sh.spec_lower;
        c->shadow_state = sh.shadow_state; c->sh[0] = sh.sh[0]; c->sh[1] = sh.sh[1];
        spec_hits++;
      }
    }
    leaves.push_back(c);
  }
  int bound_and_branch(Node &leaf) {                                      // workspace.py:282-334
    osqp_iter += leaf.num_iter; osqp_solve_time += leaf.seconds;
    if (leaf.status == BQP_PRIMAL_INFEASIBLE || leaf.status == BQP_DUAL_INFEASIBLE) return BQP_OK;
    if (leaf.lower > upper_glob) return BQP_OK;
    if (int_feas(*leaf.x, leaf.frac_idx)) { x_best = *leaf.x; upper_glob = leaf.lower; prune(); return BQP_OK; }
    Vec x_int = *leaf.x;                                                  // rounding heuristic against the ROOT bounds
    for (int k = 0; k < n_int; k++) x_int[i_idx[k]] = std::nearbyint(x_int[i_idx[k]]);
    if (satisfies_lin(x_int, l_root, u_root)) {
      const double o = obj(x_int);
      if (o < upper_glob) { upper_glob = o; x_best = x_int; prune(); }
    }
    if (s.branching_rule != 0) return BQP_BNB_E_BRANCH_RULE;
    const int nextvar = most_fractional(*leaf.x, leaf.frac_idx);
    leaf.constr_idx = m + nextvar; leaf.nextvar_idx = i_idx[nextvar];
    decisions.push_back(leaf.constr_idx); decisions.push_back(leaf.nextvar_idx);
    add_child(leaf, 0); add_child(leaf, 1);
    lower_glob = kInf;
    for (auto &lf : leaves) lower_glob = std::fmin(lower_glob, lf->lower);
    return BQP_OK;
  }
  // replay until the frontier holds an unsolved leaf (returns 1), or the tree is finished (returns 0); < 0: error
  int advance() {
    while (!leaves.empty() && iter_num < s.max_iter_bb) {
      for (auto &lf : leaves) if (unsolved(*lf)) return 1;
      NodeP leaf;
      int rc = choose_leaf(leaf); if (rc) return rc;
      solve_node(*leaf);
      rc = bound_and_branch(*leaf); if (rc) return rc;
      iter_num++;
    }
    return 0;
  }
  int run() {
    for (;;) {
      const double t0 = g_bt.on ? bt_now() : 0.0;
      const int a = advance();
      if (g_bt.on) g_bt.t_advance += bt_now() - t0;
      if (a <= 0) return a;
      const int rc = launch(); if (rc) return rc;
    }
  }
  void init(bqp_handle h_, const bqp_problem *p, const bqp_bnb_settings *s_, const double *x_incumbent, double upper_incumbent) {
    n = p->n; m_ext = p->m; n_int = p->n_int; m = p->m - p->n_int;
    P.rows = P.cols = p->n; P.p = p->Pp; P.i = p->Pi; P.x = p->Px;
    A.rows = p->m; A.cols = p->n; A.p = p->Ap; A.i = p->Ai; A.x = p->Ax;
    q = p->q; i_idx = p->i_idx; s = *s_; h = h_;
    l_root.assign(p->l, p->l + p->m); u_root.assign(p->u, p->u + p->m);
    auto root = std::make_shared<Node>();
    root->l = l_root; root->u = u_root;
    root->x = std::make_shared<Vec>(p->n, 0.0); root->y = std::make_shared<Vec>(p->m, 0.0);
    leaves.push_back(root);
    x_best.assign(p->n, 0.0);
    if (x_incumbent && std::isfinite(upper_incumbent)) { x_best.assign(x_incumbent, x_incumbent + p->n); upper_glob = upper_incumbent; }
  }
  void finish(const bqp_problem *p, double *x, bqp_bnb_result *res, int *dec, int decisions_cap) {
    std::memset(res, 0, sizeof(*res));
    res->iter_num = iter_num; res->osqp_iter = osqp_iter; res->osqp_solve_time = osqp_solve_time;
    res->upper_glob = upper_glob; res->lower_glob = lower_glob;
    res->batches = batches; res->batched_nodes = batched_nodes; res->spec_nodes = spec_nodes; res->spec_hits = spec_hits;
    res->n_decisions = (int)(decisions.size() / 2);
    res->open_leaves = (int)leaves.size();
    if (dec) std::memcpy(dec, decisions.data(), sizeof(int) * std::min<size_t>(decisions.size(), 2 * (size_t)std::max(0, decisions_cap)));
    // workspace.py:352-384
    const bool finished = iter_num < s.max_iter_bb;
    if (upper_glob != kInf) res->status = finished ? BQP_MI_SOLVED : BQP_MI_MAX_ITER_FEASIBLE;
    else if (upper_glob >= 0) res->status = finished ? BQP_MI_PRIMAL_INFEASIBLE : BQP_MI_MAX_ITER_UNSOLVED;
    else res->status = BQP_MI_DUAL_INFEASIBLE;
    if (res->status == BQP_MI_SOLVED || res->status == BQP_MI_MAX_ITER_FEASIBLE)
      for (int k = 0; k < n_int; k++) x_best[p->i_idx[k]] = std::nearbyint(x_best[p->i_idx[k]]);
    std::memcpy(x, x_best.data(), sizeof(double) * (size_t)p->n);
  }
};

}  // namespace

static bool problem_ok(const bqp_problem *p) {
  return p && p->n > 0 && p->m >= p->n_int && p->n_int >= 0 && p->Pp && p->Pi && p->Px && p->Ap && p->Ai && p->Ax && p->q && p->l &&
         p->u && (!p->n_int || p->i_idx);
}

extern "C" int bqp_bnb_solve(bqp_handle h, const bqp_problem *p, const bqp_bnb_settings *s, const double *x_incumbent,
                             double upper_incumbent, bqp_solve_fn fn, void *ctx, double *x, bqp_bnb_result *res,
                             int *decisions, int decisions_cap) {
  if (!s || !x || !res || (!h && !fn) || !problem_ok(p)) return BQP_E_ARG;
  Tree t;
  t.init(h, p, s, x_incumbent, upper_incumbent);
  t.fn = fn; t.ctx = ctx;
  const int rc = t.run();
  t.finish(p, x, res, decisions, decisions_cap);
  return rc;
}

// Lock-step over several MIQPs (miqp.py solve_many; BASELINE config 2): every step replays each tree up to its next
// unsolved leaves, flattens all of them (+ per-tree look-ahead) into ONE launch, longest-first by the parent's iteration
// count, and shares the launch time by iteration count.  Each tree's result equals its own bqp_bnb_solve.
extern "C" int bqp_bnb_solve_many(int count, const bqp_handle *h, const bqp_problem *const *p, const bqp_bnb_settings *s,
                                  const double *const *x_incumbent, const double *upper_incumbent, bqp_solve_many_fn fn,
                                  void *ctx, double *const *x, bqp_bnb_result *res, int *const *decisions, int decisions_cap) {
  if (count <= 0 || !p || !s || !x || !res || (!h && !fn)) return BQP_E_ARG;
  for (int k = 0; k < count; k++) if (!problem_ok(p[k]) || !x[k] || (!fn && !h[k])) return BQP_E_ARG;
  std::vector<std::unique_ptr<Tree>> trees;
  for (int k = 0; k < count; k++) {
    trees.emplace_back(new Tree());
    trees.back()->init(h ? h[k] : nullptr, p[k], &s[k], x_incumbent ? x_incumbent[k] : nullptr, upper_incumbent ? upper_incumbent[k] : kInf);
  }
  std::vector<char> active((size_t)count, 1);
  int rc = BQP_OK;
  for (;;) {
    std::vector<Node *> batch; std::vector<int> owner;
    for (int k = 0; k < count && !rc; k++) {
      if (!active[(size_t)k]) continue;
      const int a = trees[(size_t)k]->advance();
      if (a < 0) { rc = a; break; }
      if (a == 0) { active[(size_t)k] = 0; continue; }
      const size_t first = batch.size();
      trees[(size_t)k]->collect(batch);
      owner.insert(owner.end(), batch.size() - first, k);
    }
    if (rc || batch.empty()) break;
    const int B = (int)batch.size();
    std::vector<int> order((size_t)B);
    for (int b = 0; b < B; b++) order[(size_t)b] = b;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return batch[(size_t)a]->parent_iters > batch[(size_t)b]->parent_iters; });
    std::vector<const double *> pl(B), pu(B), px0(B), py0(B);
    std::vector<double *> px(B), py(B);
    std::vector<int> status(B), iters(B), own(B);
    std::vector<bqp_handle> hs(B);
    for (int j = 0; j < B; j++) {
      Node &nd = *batch[(size_t)order[(size_t)j]];
      own[j] = owner[(size_t)order[(size_t)j]];
      hs[j] = h ? h[own[j]] : nullptr;
      pl[j] = nd.l.data(); pu[j] = nd.u.data(); px0[j] = nd.x->data(); py0[j] = nd.y->data(); px[j] = nd.cx->data(); py[j] = nd.cy->data();
    }
    const auto t0 = std::chrono::steady_clock::now();
    if (fn) rc = fn(ctx, B, own.data(), pl.data(), pu.data(), px0.data(), py0.data(), px.data(), py.data(), status.data(), iters.data());
    else {
      bqp_node_out out; std::memset(&out, 0, sizeof(out));
      out.status = status.data(); out.iters = iters.data();
      rc = bqp_solve_multi(B, hs.data(), pl.data(), pu.data(), px0.data(), py0.data(), px.data(), py.data(), &out);
    }
    if (rc) break;
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long long total = 0; for (int j = 0; j < B; j++) total += iters[j];
    if (total < 1) total = 1;
    std::vector<char> seen((size_t)count, 0);
    for (int j = 0; j < B; j++) {
      Tree::absorb(*batch[(size_t)order[(size_t)j]], status[j], iters[j], dt * (double)iters[j] / (double)total);
      Tree &t = *trees[(size_t)own[j]];
      t.batched_nodes++;
      if (!seen[(size_t)own[j]]) { seen[(size_t)own[j]] = 1; t.batches++; }
    }
  }
  for (int k = 0; k < count; k++) trees[(size_t)k]->finish(p[k], x[k], &res[k], decisions ? decisions[k] : nullptr, decisions_cap);
  return rc;
}

// Asynchronous variant (include/bqp.h): the lock-step driver above makes every tree wait, at every B&B step, for the slowest
// leaf of all trees; here each tree runs its own loop on its own engine context (CUDA stream), and the device's block
// scheduler interleaves the tiles of all of them.  Trees do not interact, so each result equals bqp_bnb_solve's.
extern "C" int bqp_bnb_solve_async(int count, const bqp_handle *h, const bqp_problem *const *p, const bqp_bnb_settings *s,
                                   const double *const *x_incumbent, const double *upper_incumbent, double *const *x,
                                   bqp_bnb_result *res, int *const *decisions, int decisions_cap, int threads) {
  if (count <= 0 || !h || !p || !s || !x || !res) return BQP_E_ARG;
  for (int k = 0; k < count; k++) if (!problem_ok(p[k]) || !x[k] || !h[k]) return BQP_E_ARG;
  int nt = threads > 0 ? threads : 128;
  nt = std::max(1, std::min(nt, count));
  std::atomic<int> next(0), first_err(BQP_OK);
  // engine contexts are kept for later calls (a closed-loop user calls this once per sampling instant): creating one costs
  // stream + event + first-use allocations
  static std::mutex pool_mu;
  static std::vector<std::pair<int, bqp_ctx>> pool_ctx;     // (device, context)
  auto take_ctx = [&](int device, bqp_ctx *out) {
    {
      std::lock_guard<std::mutex> lk(pool_mu);
      for (size_t i = 0; i < pool_ctx.size(); i++)
        if (pool_ctx[i].first == device) { *out = pool_ctx[i].second; pool_ctx.erase(pool_ctx.begin() + (long)i); return (int)BQP_OK; }
    }
    return bqp_ctx_create(device, 1, out);
  };
  auto worker = [&]() {
    bqp_ctx ectx = nullptr; int edev = -1;
    for (;;) {
      const int k = next.fetch_add(1);
      if (k >= count) break;
      Tree t;
      t.init(h[k], p[k], &s[k], x_incumbent ? x_incumbent[k] : nullptr, upper_incumbent ? upper_incumbent[k] : kInf);
      int rc = BQP_OK;
      const int dev = bqp_handle_device(h[k]);
      if (ectx && edev != dev) { std::lock_guard<std::mutex> lk(pool_mu); pool_ctx.emplace_back(edev, ectx); ectx = nullptr; }
      if (!ectx) { rc = take_ctx(dev, &ectx); edev = dev; }
      if (!rc) { t.ectx = ectx; rc = t.run(); }
      t.finish(p[k], x[k], &res[k], decisions ? decisions[k] : nullptr, decisions_cap);
      if (rc) { int ok = BQP_OK; first_err.compare_exchange_strong(ok, rc); }
    }
    if (ectx) { std::lock_guard<std::mutex> lk(pool_mu); pool_ctx.emplace_back(edev, ectx); }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; t++) pool.emplace_back(worker);
  worker();
  for (auto &th : pool) th.join();
  return first_err.load();
}

// Rolling variant (include/bqp.h): one engine session shared by all trees.  A round is one launch; after it, every tree whose
// outstanding leaves have all terminated absorbs their results, replays (advance) and appends its next unsolved leaves
// (+ look-ahead), which join the next round next to the leaves of the other trees that are still iterating.
static int rolling_on_ctx(bqp_ctx sctx, int count, const bqp_handle *h, const bqp_problem *const *p, const bqp_bnb_settings *s,
                          const double *const *x_incumbent, const double *upper_incumbent, double *const *x,
                          bqp_bnb_result *res, int *const *decisions, int decisions_cap, int *rounds, int host_threads) {
  std::vector<std::unique_ptr<Tree>> trees;
  for (int k = 0; k < count; k++) {
    trees.emplace_back(new Tree());
    trees.back()->init(h[k], p[k], &s[k], x_incumbent ? x_incumbent[k] : nullptr, upper_incumbent ? upper_incumbent[k] : kInf);
  }
  struct Pending { int tree; Node *node; std::chrono::steady_clock::time_point t0; };
  std::vector<Pending> pend;                      // by session node id
  std::vector<int> outstanding((size_t)count, 0);
  std::vector<char> active((size_t)count, 1);
  int rc = bqp_session_begin(sctx), nrounds = 0;
  std::vector<int> fin;
  // BQP_ROLLING_TIMERS=1: where the wall time of the run goes (host replay / append / round = launch + wait / fetch), to stderr
  const bool timers = std::getenv("BQP_ROLLING_TIMERS") != nullptr;
  double t_replay = 0, t_append = 0, t_round = 0, t_fetch = 0; long long n_fetch = 0, n_append = 0;
  auto tick = []() { return std::chrono::steady_clock::now(); };
  auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
  while (!rc) {
    auto tp = tick();
    // replay every tree that waits for nothing and collect its next leaves.  The replay of one node costs two sparse
    // mat-vecs on the host (objective at the clipped point, feasibility of the rounded point: ~1 ms at n = 500), the trees
    // are independent: spread them over host threads
    std::vector<int> ready;
    for (int k = 0; k < count; k++) if (active[(size_t)k] && outstanding[(size_t)k] == 0) ready.push_back(k);
    std::vector<std::vector<Node *>> got(ready.size());
    std::vector<int> adv(ready.size(), 0);
    {
      std::atomic<size_t> nexti(0);
      auto work = [&]() {
        for (;;) {
          const size_t i = nexti.fetch_add(1);
          if (i >= ready.size()) break;
          Tree &t = *trees[(size_t)ready[i]];
          adv[i] = t.advance();
          if (adv[i] == 1) t.collect(got[i]);
        }
      };
      const size_t nth = std::min<size_t>(ready.size(), (size_t)std::max(1, host_threads));
      std::vector<std::thread> pool;
      for (size_t t = 1; t < nth; t++) pool.emplace_back(work);
      work();
      for (auto &th : pool) th.join();
    }
    t_replay += since(tp); tp = tick();
    std::vector<Node *> batch; std::vector<int> owner;
    for (size_t i = 0; i < ready.size() && !rc; i++) {
      const int k = ready[i];
      if (adv[i] < 0) { rc = adv[i]; break; }
      if (adv[i] == 0) { active[(size_t)k] = 0; continue; }
      batch.insert(batch.end(), got[i].begin(), got[i].end());
      owner.insert(owner.end(), got[i].size(), k);
      outstanding[(size_t)k] = (int)got[i].size();
      trees[(size_t)k]->batches++;
    }
    if (rc) break;
    if (!batch.empty()) {
      const int B = (int)batch.size();
      std::vector<const double *> pl(B), pu(B), px0(B), py0(B);
      std::vector<bqp_handle> hs(B);
      for (int j = 0; j < B; j++) {
        Node &nd = *batch[(size_t)j];
        hs[j] = h[owner[(size_t)j]];
        pl[j] = nd.l.data(); pu[j] = nd.u.data(); px0[j] = nd.x->data(); py0[j] = nd.y->data();
      }
      int first_id = 0;
      rc = bqp_session_append(sctx, B, hs.data(), pl.data(), pu.data(), px0.data(), py0.data(), &first_id);
      if (rc) break;
      if ((size_t)first_id != pend.size()) { rc = BQP_E_ARG; break; }
      const auto now = std::chrono::steady_clock::now();
      for (int j = 0; j < B; j++) { pend.push_back({owner[(size_t)j], batch[(size_t)j], now}); trees[(size_t)owner[(size_t)j]]->batched_nodes++; }
    }
    n_append += (long long)batch.size();
    t_append += since(tp); tp = tick();
    bool any = false;
    for (int k = 0; k < count; k++) any = any || outstanding[(size_t)k] > 0;
    if (!any) break;                              // every tree is finished
    fin.assign(pend.size(), 0);
    int nfin = 0, running = 0;
    rc = bqp_session_round(sctx, fin.data(), (int)fin.size(), &nfin, &running);
    if (rc) break;
    nrounds++;
    t_round += since(tp); tp = tick();
    const auto now = std::chrono::steady_clock::now();
    for (int j = 0; j < nfin && !rc; j++) {
      Pending &pd = pend[(size_t)fin[(size_t)j]];
      int status = BQP_UNSOLVED, iters = 0;
      bqp_node_out out; std::memset(&out, 0, sizeof(out));
      out.status = &status; out.iters = &iters;
      rc = bqp_session_fetch(sctx, fin[(size_t)j], pd.node->cx->data(), pd.node->cy->data(), &out);
      // a node's time = from the append to the end of the round it terminated in, shared by the tree's leaves in flight
      Tree::absorb(*pd.node, status, iters, std::chrono::duration<double>(now - pd.t0).count() / 2.0);
      outstanding[(size_t)pd.tree]--;
    }
    n_fetch += nfin;
    t_fetch += since(tp);
  }
  if (timers)
    std::fprintf(stderr, "ROLLING %d rounds: replay %.3f s, append %.3f s (%lld nodes), round %.3f s, fetch %.3f s (%lld nodes)\n", nrounds,
                 t_replay, t_append, n_append, t_round, t_fetch, n_fetch);
  for (int k = 0; k < count; k++) trees[(size_t)k]->finish(p[k], x[k], &res[k], decisions ? decisions[k] : nullptr, decisions_cap);
  if (rounds) *rounds = nrounds;
  return rc;
}

// `sessions` > 1: the trees are dealt round-robin to that many sessions, each on its own engine context (CUDA stream) and
// host thread: while one session replays its finished trees on the host, the other sessions' rounds keep the GPU busy, and
// the tiles of their launches share the SMs.  sessions <= 0: 1 -- measured on the 100 config-2 trees, 2 / 3 / 4 sessions take
// 9.9 / 8.7 / 9.4 s against 8.2 s for one: each session then plans for its share of the SMs only, and the longest trees,
// which bound the run, get a smaller machine.
extern "C" int bqp_bnb_solve_rolling(int count, const bqp_handle *h, const bqp_problem *const *p, const bqp_bnb_settings *s,
                                     const double *const *x_incumbent, const double *upper_incumbent, double *const *x,
                                     bqp_bnb_result *res, int *const *decisions, int decisions_cap, int *rounds, int sessions) {
  if (count <= 0 || !h || !p || !s || !x || !res) return BQP_E_ARG;
  for (int k = 0; k < count; k++) if (!problem_ok(p[k]) || !x[k] || !h[k]) return BQP_E_ARG;
  int ns = sessions > 0 ? sessions : 1;
  ns = std::max(1, std::min(ns, std::min(count, 8)));
  const int hw = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  if (ns == 1) return rolling_on_ctx(nullptr, count, h, p, s, x_incumbent, upper_incumbent, x, res, decisions, decisions_cap, rounds, hw);
  std::vector<int> rcs((size_t)ns, BQP_OK), nr((size_t)ns, 0);
  auto run = [&](int g) {
    std::vector<bqp_handle> hh; std::vector<const bqp_problem *> pp; std::vector<bqp_bnb_settings> ss;
    std::vector<const double *> xi; std::vector<double> ui; std::vector<double *> xx; std::vector<int *> dd; std::vector<int> idx;
    for (int k = g; k < count; k += ns) {
      idx.push_back(k); hh.push_back(h[k]); pp.push_back(p[k]); ss.push_back(s[k]);
      xi.push_back(x_incumbent ? x_incumbent[k] : nullptr); ui.push_back(upper_incumbent ? upper_incumbent[k] : kInf);
      xx.push_back(x[k]); dd.push_back(decisions ? decisions[k] : nullptr);
    }
    std::vector<bqp_bnb_result> rr(idx.size());
    bqp_ctx c = nullptr;
    int rc = bqp_ctx_create(bqp_handle_device(hh[0]), 0, &c);
    if (!rc) {
      bqp_ctx_set_sm_share(c, ns);
      rc = rolling_on_ctx(c, (int)idx.size(), hh.data(), pp.data(), ss.data(), xi.data(), ui.data(), xx.data(), rr.data(),
                          decisions ? dd.data() : nullptr, decisions_cap, &nr[(size_t)g], std::max(1, hw / ns));
    }
    for (size_t i = 0; i < idx.size(); i++) res[idx[i]] = rr[i];
    bqp_ctx_free(c);
    rcs[(size_t)g] = rc;
  };
  std::vector<std::thread> pool;
  for (int g = 1; g < ns; g++) pool.emplace_back(run, g);
  run(0);
  for (auto &th : pool) th.join();
  int rc = BQP_OK, total = 0;
  for (int g = 0; g < ns; g++) { if (!rc) rc = rcs[(size_t)g]; total = std::max(total, nr[(size_t)g]); }
  if (rounds) *rounds = total;
  return rc;
}
