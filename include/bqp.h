/*
 * bqp.h -- C ABI of the B200 batched QP-relaxation engine (libbqp.so).
 *
 * This is the drop-in boundary for miOSQP's hot path: every entry point below
 * replaces one call the reference makes on its `osqp` object
 * (file:line under /root/reference/miosqp/):
 *
 *   bqp_setup             <- osqp.OSQP().setup(P,q,A,l,u,**qp_settings)     workspace.py:63-68
 *   bqp_update_q          <- solver.update(q=q)                             solver.py:185
 *   bqp_solve_batch       <- Node.solve(): update(l,u)+warm_start(x,y)+solve()+clip+objective
 *                            for B nodes of one problem at once              node.py:96-143
 *   bqp_solve_multi       <- the same for nodes of several set-up problems in ONE launch
 *                            (frontier of many MIQP instances, BASELINE cfg 2)
 *   bqp_bnb_solve         <- MIOSQP.solve(): the B&B while-loop itself, natively  solver.py:85-172
 *   bqp_bnb_solve_many    <- the same for several MIQPs in lock-step (one launch per B&B step over all frontiers)
 *   bqp_bnb_solve_async   <- the same, every MIQP advancing at its own pace on its own stream (no lock-step)
 *   bqp_ctx_*             <- one solve context (stream + staging) per host thread: the reference's one osqp object per Workspace
 *   bqp_setup_many        <- setup of many problems, host halves on all host threads
 *   BQP_* status codes    <- osqp.constant('OSQP_*')                        node.py:88,128-129
 *   bqp_free              <- garbage collection of the osqp object
 *
 * Conventions: plain pointers and sizes, no torch/numpy types.  All vectors
 * are FP64 HOST buffers owned by the caller (the engine copies in/out and
 * never keeps a caller pointer); batch arrays are node-major ([B][m], [B][n]).
 * Return value 0 = OK, negative = bqp_error.  Per-node solver outcomes use
 * OSQP's integer status codes in status[].  A handle is not thread-safe.
 * There is NO CPU fallback: setup and every solve call fail with BQP_E_CUDA
 * when no sm_100 device is usable.
 */
#ifndef BQP_H
#define BQP_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bqp_instance *bqp_handle;

typedef enum {
  BQP_OK = 0,
  BQP_E_ARG = -1,        /* null pointer / bad dimension / bad setting */
  BQP_E_BOUNDS = -2,     /* l > u in some row (osqp raises ValueError) */
  BQP_E_NONCONVEX = -3,  /* reduced KKT matrix not positive definite */
  BQP_E_CUDA = -4,       /* CUDA runtime error or no usable device */
  BQP_E_ALLOC = -5,      /* out of host or device memory */
  BQP_E_UNSUPPORTED = -6, /* setting outside the parity contract (e.g. adaptive_rho) or problem too large for one CTA */
  BQP_BNB_E_EXPLOR_RULE = -7, /* bqp_bnb_solve: 'Tree exploring strategy not recognized'   workspace.py:147 */
  BQP_BNB_E_BRANCH_RULE = -8  /* bqp_bnb_solve: 'No variable selection rule recognized!'   workspace.py:224 */
} bqp_error;

/* OSQP status codes written to status[] (osqp.constant values) */
enum {
  BQP_SOLVED = 1, BQP_SOLVED_INACCURATE = 2, BQP_PRIMAL_INFEASIBLE_INACCURATE = 3,
  BQP_DUAL_INFEASIBLE_INACCURATE = 4, BQP_MAX_ITER_REACHED = -2, BQP_PRIMAL_INFEASIBLE = -3,
  BQP_DUAL_INFEASIBLE = -4, BQP_NON_CVX = -7, BQP_UNSOLVED = -10
};

typedef struct {
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  int max_iter;
  int scaling;            /* Ruiz passes (osqp default 10) */
  int check_termination;  /* residual test period (osqp default 25), >= 1 */
  int eq_rho;             /* 1: type rho per row ONCE at setup (equality x1e3, free rows RHO_MIN) */
  int device;             /* CUDA device ordinal */
  /* osqp adaptive_rho with a FIXED interval (a multiple of check_termination; osqp's automatic interval is derived from
   * wall-clock setup time and is not reproducible: refused).  Every node starts from `rho` and adapts on its own
   * (adapt_rho / compute_rho_estimate of osqp 0.6); runs on the whole-GPU kernel through the spectral form of the reduced
   * KKT inverse, K(rho)^-1 = V diag(1 / (1 + (rho - rho0) mu)) V': any rho per node without refactorisation. */
  int adaptive_rho, adaptive_rho_interval;
  double adaptive_rho_tolerance;   /* osqp default 5 */
} bqp_settings;

/* QP  min 1/2 x'Px + q'x  s.t. l <= Ax <= u ;  P upper-triangular CSC, A CSC (m x n).
 * For miOSQP, A/l/u are the EXTENDED data of data.py:5-33: the last n_int rows of A are
 * I[i_idx,:]; i_idx (variable index per integer row, in row order) drives the in-kernel
 * clip of node.py:131-136.  n_int = 0 gives a plain batched QP solver. */
typedef struct {
  int n, m;
  const int *Pp, *Pi; const double *Px;
  const int *Ap, *Ai; const double *Ax;
  const double *q, *l, *u;
  int n_int; const int *i_idx;
} bqp_problem;

/* per-node results beyond x,y (arrays of length B; any pointer may be NULL) */
typedef struct {
  int *status;        /* OSQP status code                         node.py:111 */
  int *iters;         /* ADMM iterations                          node.py:118 */
  double *obj;        /* osqp info.obj_val                                    */
  double *pri_res;    /* osqp info.pri_res                                    */
  double *dua_res;    /* osqp info.dua_res                                    */
  double *lower;      /* 1/2 x'Px+q'x at the clipped x (node.lower); NaN unless status in {1,-2}  node.py:128-143 */
} bqp_node_out;

/* device-side timing of the last call, CUDA events on the engine's stream (milliseconds) */
typedef struct {
  double h2d_ms, kernel_ms, d2h_ms;
  long long h2d_bytes, d2h_bytes;
  int launches, tiles, tile_nodes, threads;
  long long smem_bytes;
  long long node_iters;     /* sum over nodes of ADMM iterations executed until that node terminated */
  long long tile_iters;     /* sum over tiles of iterations the tile ran (= max over its nodes) */
  long long stream_bytes;   /* matrix/factor bytes the kernel streamed: sum over tiles of iterations x per-iteration panel bytes (+ checks) */
  int kernel;               /* which kernel ran: 0 direct-load tile kernel, 1 TMA-streamed two-pass kernel, 2 fused single-pass panel kernel (column-split pair), 3 row-split cluster kernel, 4 whole-GPU kernel (one tile on every SM), 5 shared-memory-resident kernel (npad <= 64) */
  int ring_slots;           /* shared-memory ring depth (TMA stages / panels in flight) of the first launch */
} bqp_timing;

void bqp_default_settings(bqp_settings *s);
int bqp_setup(const bqp_problem *p, const bqp_settings *s, bqp_handle *out);
/* bqp_setup for `count` problems sharing one settings struct: host halves on `threads` host threads (0 = all), device
 * uploads one by one.  out[count].  All or nothing: on any error every handle is freed and out[] is all NULL.
 * host_only != 0 is the layout-test variant of bqp_debug_host_setup (no device touched). */
int bqp_setup_many(int count, const bqp_problem *const *p, const bqp_settings *s, bqp_handle *out, int threads, int host_only);
int bqp_update_q(bqp_handle h, const double *q);
int bqp_solve_batch(bqp_handle h, int B, const double *l, const double *u, const double *x0, const double *y0,
                    double *x, double *y, const bqp_node_out *out);
/* node b belongs to handles[b]; l[b],u[b],y0[b],y[b] have handles[b]->m entries, x0[b],x[b] have n.
 * All handles must live on the same device. */
int bqp_solve_multi(int B, const bqp_handle *handles, const double *const *l, const double *const *u,
                    const double *const *x0, const double *const *y0, double *const *x, double *const *y,
                    const bqp_node_out *out);
/* staged variant (inputs stay resident in HBM between runs; used to time the kernel alone):
 * upload = pack + H2D, run = all kernels from the resident inputs (blocking), download = D2H + unpack */
int bqp_batch_upload(int B, const bqp_handle *handles, const double *const *l, const double *const *u,
                     const double *const *x0, const double *const *y0);
int bqp_batch_run(void);
int bqp_batch_download(double *const *x, double *const *y, const bqp_node_out *out);
int bqp_last_timing(bqp_timing *t);
int bqp_free(bqp_handle h);

/* Explicit solve contexts.  The handle-less calls above share ONE process-wide context (one stream, one set of staging
 * buffers, one resident batch); a bqp_ctx is the same thing as an object: its own CUDA stream, pinned staging and device
 * buffers, so several host threads can each drive their own frontier on the same device at the same time (their kernels
 * overlap on the GPU) -- what miOSQP's one-`osqp`-object-per-Workspace gives the reference (workspace.py:63).
 * run_to_completion != 0: every tile runs until its last node terminates (one launch per call) instead of rounds of 100
 * ADMM iterations re-tiled by the host -- the right shape when the call holds one small tile.  A context is not
 * thread-safe; different contexts are independent.  bqp_free / bqp_update_q synchronise every context of the device. */
typedef struct bqp_context *bqp_ctx;
int bqp_ctx_create(int device, int run_to_completion, bqp_ctx *out);
int bqp_ctx_solve_multi(bqp_ctx ctx, int B, const bqp_handle *handles, const double *const *l, const double *const *u,
                        const double *const *x0, const double *const *y0, double *const *x, double *const *y,
                        const bqp_node_out *out);
int bqp_ctx_last_timing(bqp_ctx ctx, bqp_timing *t);
int bqp_ctx_free(bqp_ctx ctx);
/* Dense-A problems normally run on a fixed cluster size chosen from the problem alone (2 CTAs per tile at n = 500), so that
 * a node's result does not depend on what else is in the launch -- bit for bit.  on != 0 lets the context (NULL: the
 * process-wide one) spread a tile over 4 or 8 CTAs whenever a launch holds few tiles (<= 33 / <= 16): the per-iteration
 * latency drops from 80 us to 45 / ~25 us at n = 500, which is what a single B&B tree (two leaves per step) is bound by.
 * The summation order of the column sums then depends on the schedule: results agree to rounding, not to the last bit.
 * Rolling sessions (bqp_session_*) have it on by default; BQP_ROWS_AUTO_CLUSTER=0/1 overrides everything. */
int bqp_ctx_set_auto_cluster(bqp_ctx ctx, int on);
/* the context shares the device with `parts` - 1 other contexts that are busy at the same time: it plans at most 1 / parts
 * of the SMs per launch (rounds, automatic cluster sizes) */
int bqp_ctx_set_sm_share(bqp_ctx ctx, int parts);

/* Rolling session on a context (ctx == NULL: the process-wide one): the resident batch is OPEN -- nodes are appended
 * while earlier ones are still iterating, every bqp_session_round is ONE launch (a round of 100 ADMM iterations over the
 * running nodes, critical path first, at most one tile per SM pair) and reports the nodes that terminated in it, whose
 * results bqp_session_fetch then copies out.  This is how many B&B trees share one GPU without waiting for each other's
 * slowest leaf: a tree whose leaves have terminated is replayed and its children join the next round.
 * ids count from 0 in append order; finished_ids[cap]; out arrays of bqp_session_fetch have length 1. */
int bqp_session_begin(bqp_ctx ctx);
int bqp_session_append(bqp_ctx ctx, int B, const bqp_handle *handles, const double *const *l, const double *const *u,
                       const double *const *x0, const double *const *y0, int *first_id);
int bqp_session_round(bqp_ctx ctx, int *finished_ids, int cap, int *n_finished, int *running);
int bqp_session_fetch(bqp_ctx ctx, int id, double *x, double *y, const bqp_node_out *out);

/* ---- native branch-and-bound replay (miosqp_b200/csrc/bqp_bnb.cpp) ------------------------------------------------
 * bqp_bnb_solve <- MIOSQP.solve(): the while-loop of solver.py:85-123 with workspace.py:128-384 and node.py:96-143,
 * every open leaf solved by the batched engine (one bqp_solve_multi per B&B step, plus `speculation` look-ahead nodes).
 * Same decisions as the reference's sequential loop; statuses are the strings of constants.py:2-7 by index. */
enum { BQP_MI_UNSOLVED = 0, BQP_MI_SOLVED = 1, BQP_MI_PRIMAL_INFEASIBLE = 2, BQP_MI_DUAL_INFEASIBLE = 3,
       BQP_MI_MAX_ITER_FEASIBLE = 4, BQP_MI_MAX_ITER_UNSOLVED = 5 };
typedef struct {
  double eps_int_feas;     /* settings['eps_int_feas']                                   workspace.py:257 */
  int max_iter_bb;         /* settings['max_iter_bb']                                    workspace.py:113-126 */
  int tree_explor_rule;    /* 0 depth first, 1 two-phase                                 workspace.py:128-155 */
  int branching_rule;      /* 0 most fractional                                          workspace.py:205-230 */
  int speculation;         /* look-ahead nodes per launch (0 = the reference's two nodes per step) */
  double eps_abs;          /* qp_settings['eps_abs']: feasibility slack of the rounding heuristic   workspace.py:238 */
} bqp_bnb_settings;
typedef struct {
  int status;              /* BQP_MI_* */
  int iter_num;            /* starts at 1, as the reference's counter */
  long long osqp_iter;     /* ADMM iterations of the consumed nodes */
  double osqp_solve_time, upper_glob, lower_glob;
  int batches;             /* launches */
  long long batched_nodes, spec_nodes, spec_hits;
  int n_decisions, open_leaves;
} bqp_bnb_result;
/* batch solver used instead of the engine (tests): same contract as bqp_solve_multi with one problem, returns 0 or a bqp_error */
typedef int (*bqp_solve_fn)(void *ctx, int B, const double *const *l, const double *const *u, const double *const *x0,
                            const double *const *y0, double *const *x, double *const *y, int *status, int *iters);
/* p: the UNSCALED problem as Data holds it (P as the caller passed it -- full symmetric for the reference's examples --,
 * q, extended A, root l/u, i_idx).  x_incumbent/upper_incumbent: what set_x0 accepted (NULL / +inf: none).
 * fn == NULL: nodes are solved on the device through h.  x[n]: the returned solution (integer entries rounded).
 * decisions: up to decisions_cap (constr_idx, nextvar_idx) pairs. */
int bqp_bnb_solve(bqp_handle h, const bqp_problem *p, const bqp_bnb_settings *s, const double *x_incumbent,
                  double upper_incumbent, bqp_solve_fn fn, void *ctx, double *x, bqp_bnb_result *res,
                  int *decisions, int decisions_cap);

/* lock-step variant over `count` MIQPs (miqp.py solve_many): one launch per B&B step covers every tree's unsolved leaves
 * (+ per-tree look-ahead); s, res: [count]; x, decisions, x_incumbent: [count] pointers; h: [count] handles (NULL with fn).
 * fn additionally receives, per node, the index of the problem it belongs to. */
typedef int (*bqp_solve_many_fn)(void *ctx, int B, const int *owner, const double *const *l, const double *const *u,
                                 const double *const *x0, const double *const *y0, double *const *x, double *const *y,
                                 int *status, int *iters);
int bqp_bnb_solve_many(int count, const bqp_handle *h, const bqp_problem *const *p, const bqp_bnb_settings *s,
                       const double *const *x_incumbent, const double *upper_incumbent, bqp_solve_many_fn fn, void *ctx,
                       double *const *x, bqp_bnb_result *res, int *const *decisions, int decisions_cap);

/* asynchronous variant: `threads` host threads (0 = min(count, 128)), each with its own solve context (CUDA stream), take the MIQPs
 * from a shared counter and run each one's loop (replay -> one launch of its unsolved leaves + look-ahead, run to
 * completion -> replay ...) independently: no MIQP waits for another one's slowest leaf, and the GPU's block scheduler
 * packs the tiles of all of them.  Every tree's result equals its own bqp_bnb_solve. */
int bqp_bnb_solve_async(int count, const bqp_handle *h, const bqp_problem *const *p, const bqp_bnb_settings *s,
                        const double *const *x_incumbent, const double *upper_incumbent, double *const *x, bqp_bnb_result *res,
                        int *const *decisions, int decisions_cap, int threads);

/* rolling variant: the trees share a bqp_session; after every round the trees whose outstanding leaves have all terminated
 * are replayed and their new leaves appended.  sessions > 1 deals the trees to that many sessions (own context, stream and
 * host thread each), so that one session's host replay overlaps the others' rounds (<= 0: one session, the measured best).  Each tree's result
 * equals its own bqp_bnb_solve up to the rounding of the cluster size (bqp_ctx_set_auto_cluster). */
int bqp_bnb_solve_rolling(int count, const bqp_handle *h, const bqp_problem *const *p, const bqp_bnb_settings *s,
                          const double *const *x_incumbent, const double *upper_incumbent, double *const *x, bqp_bnb_result *res,
                          int *const *decisions, int decisions_cap, int *rounds, int sessions);

/* tuning knobs (0 = automatic): nodes per tile (1,2,4,8) and threads per CTA (multiple of 32, <= 512) */
int bqp_set_tuning(int tile_nodes, int threads);

/* introspection (tests, roofline arithmetic) */
int bqp_get_dims(bqp_handle h, int *n, int *m, int *npad, long long *factor_bytes, long long *check_bytes);
int bqp_get_scaling(bqp_handle h, double *D, double *E, double *c);
/* guard of the explicit reduced inverse used by the dense-A kernels: *error = largest relative difference between KKT solves
 * through the inverse and through the LDL' factor over 4 probe right-hand sides (NaN: no dense layout was tried);
 * *in_use = 1 when the dense layout passed (error <= 1e-10, BQP_INVERSE_TOL) and the problem runs on the dense kernels,
 * 0 when it was dropped or never built and the problem runs on the LDL' kernels. */
int bqp_get_inverse_guard(bqp_handle h, double *error, int *in_use);
int bqp_handle_device(bqp_handle h);   /* CUDA device ordinal the problem lives on */
int bqp_device_count(void);
const char *bqp_strerror(int code);
const char *bqp_version(void);

/* HOST-ONLY debug hooks for the layout tests (tests/test_host_layout.py): they run the host half of
 * bqp_setup (scaling, rho typing, factor, panel layouts) without touching a device and apply the
 * streamed panels / blocked factor with plain loops.  They are NOT a solve path: no ADMM runs here. */
int bqp_debug_host_setup(const bqp_problem *p, const bqp_settings *s, bqp_handle *out);
int bqp_debug_dump_groups(bqp_handle h);   /* prints the streamed group table to stdout */
int bqp_debug_host_kkt_solve(bqp_handle h, double *rhs_xz /* [n+m], scaled space, in place */);
int bqp_debug_host_stream_kkt_solve(bqp_handle h, double *rhs_xz /* same, through the TMA kernel's streamed layout */);
int bqp_debug_host_panel_kkt_solve(bqp_handle h, double *rhs_xz /* same, through the fused kernel's row panels (explicit reduced inverse) */);
int bqp_debug_host_small_kkt_solve(bqp_handle h, double *rhs_xz /* same, through the shared-memory-resident layout of small problems (mma fragments of the explicit reduced inverse, ELL A and A') */);
int bqp_debug_host_matvec(bqp_handle h, int which /*0: A x, 1: A' y, 2: P x, 3: P x via the streamed layout, 4: P x via the row panels, 5: P x via the fragments of the small layout*/, const double *in, double *out);

#ifdef __cplusplus
}
#endif
#endif
