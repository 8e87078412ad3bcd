/*
 * oracle/osqp_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT THE PRODUCT).
 *
 * Plain-C, scalar, FP64 restatement of the OSQP ADMM algorithm that the
 * reference miOSQP calls for every branch-and-bound node:
 *     /root/reference/miosqp/workspace.py:63-68   osqp.OSQP().setup(P,q,A,l,u,**qp_settings)
 *     /root/reference/miosqp/node.py:102          solver.update(l=,u=)
 *     /root/reference/miosqp/node.py:105          solver.warm_start(x=,y=)
 *     /root/reference/miosqp/node.py:108          solver.solve()
 *     /root/reference/miosqp/solver.py:185        solver.update(q=)
 *
 * PARITY UNPINNED: the arithmetic lives in the PyPI package `osqp`
 * (un-pinned in /root/reference/setup.py:13, not vendored, not installed in
 * this image, no golden vectors in the reference tree).  This file restates
 * the published algorithm (Stellato, Banjac, Goulart, Bemporad, Boyd: "OSQP:
 * an operator splitting solver for quadratic programs", Math. Prog. Comp.
 * 2020) and the LDL^T of T. Davis ("Algorithm 849").
 * It is pinned only against reference-independent known answers
 * (tests/test_oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  Nothing under miosqp_b200/
 * links, imports or executes it.
 *
 * Parity contract (hard part 1): adaptive_rho = 0,
 * rho_vec typed ONCE at setup from the root bounds (never re-typed per
 * node), so a node's result is a pure function of (l,u,x0,y0).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>

#define OSQP_INFTY 1e30
#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4
#define RHO_MIN 1e-6
#define RHO_MAX 1e6
#define RHO_TOL 1e-4
#define RHO_EQ_OVER_RHO_INEQ 1e3
#define DIVISION_TOL (1.0 / OSQP_INFTY)

enum {
  ST_SOLVED = 1, ST_SOLVED_INACC = 2, ST_PINF_INACC = 3, ST_DINF_INACC = 4,
  ST_MAX_ITER = -2, ST_PINF = -3, ST_DINF = -4, ST_NON_CVX = -7, ST_UNSOLVED = -10
};

typedef struct {
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  int max_iter, scaling, check_termination, scaled_termination, eq_rho;
  /* adaptive rho with a FIXED interval: osqp's automatic interval (adaptive_rho_interval = 0) is derived
   * from wall-clock setup time and is not reproducible, so it is not offered.  Every node starts from settings.rho (a node
   * is a pure function of its inputs; the reference's one osqp object carries the adapted rho from node to node). */
  int adaptive_rho, adaptive_rho_interval;
  double adaptive_rho_tolerance;
} OracleSettings;

typedef struct {
  int status, iter;
  double obj_val, pri_res, dua_res, solve_time;
  int rho_updates; double rho_final;
} OracleInfo;

typedef struct {
  int n, m;
  OracleSettings s;
  /* scaled data */
  int *Pp, *Pi; double *Px;       /* triu(P) CSC, scaled */
  int *Ap, *Ai; double *Ax;       /* A CSC, scaled */
  double *q, *l, *u;              /* scaled */
  double *D, *Dinv, *E, *Einv; double c, cinv;
  double *rho_vec, *rho_inv_vec;
  int *ctype0;                    /* osqp constr_type per row as typed at setup: -1 loose, 0 inequality, 1 equality */
  /* factor of permuted KKT */
  int N; int *perm;               /* perm[new] = old */
  int *Lp, *Li; double *Lx, *Dd, *Ddinv; int *etree;
  /* sequential-API state (osqp object semantics) */
  double *x, *y, *z;
} OracleWork;

/* ---------------------------------------------------------------- helpers */
static double vec_norm_inf(const double *v, int n) {
  double r = 0; for (int i = 0; i < n; i++) { double a = fabs(v[i]); if (a > r) r = a; } return r;
}
static void limit_scaling(double *v, int n) {
  for (int i = 0; i < n; i++) {
    v[i] = v[i] < MIN_SCALING ? 1.0 : v[i];
    v[i] = v[i] > MAX_SCALING ? MAX_SCALING : v[i];
  }
}
/* column inf-norms of the symmetric matrix stored as triu CSC */
static void sym_triu_col_norms(int n, const int *Pp, const int *Pi, const double *Px, double *out) {
  for (int j = 0; j < n; j++) out[j] = 0;
  for (int j = 0; j < n; j++)
    for (int p = Pp[j]; p < Pp[j + 1]; p++) {
      int i = Pi[p]; double a = fabs(Px[p]);
      if (a > out[j]) out[j] = a;
      if (i != j && a > out[i]) out[i] = a;
    }
}

/* ------------------------------------------------- Ruiz equilibration (A.2) */
static void scale_data(OracleWork *w) {
  int n = w->n, m = w->m;
  for (int i = 0; i < n; i++) w->D[i] = 1.0;
  for (int i = 0; i < m; i++) w->E[i] = 1.0;
  w->c = 1.0;
  double *Dt = malloc(sizeof(double) * (n > 0 ? n : 1));
  double *Et = malloc(sizeof(double) * (m > 0 ? m : 1));
  for (int it = 0; it < w->s.scaling; it++) {
    sym_triu_col_norms(n, w->Pp, w->Pi, w->Px, Dt);
    for (int i = 0; i < m; i++) Et[i] = 0;
    for (int j = 0; j < n; j++) {
      double cn = 0;
      for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) {
        double a = fabs(w->Ax[p]);
        if (a > cn) cn = a;
        if (a > Et[w->Ai[p]]) Et[w->Ai[p]] = a;
      }
      if (cn > Dt[j]) Dt[j] = cn;
    }
    limit_scaling(Dt, n); limit_scaling(Et, m);
    for (int i = 0; i < n; i++) Dt[i] = 1.0 / sqrt(Dt[i]);
    for (int i = 0; i < m; i++) Et[i] = 1.0 / sqrt(Et[i]);
    for (int j = 0; j < n; j++) {
      for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) w->Px[p] *= Dt[w->Pi[p]] * Dt[j];
      for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) w->Ax[p] *= Et[w->Ai[p]] * Dt[j];
      w->q[j] *= Dt[j];
      w->D[j] *= Dt[j];
    }
    for (int i = 0; i < m; i++) w->E[i] *= Et[i];
    /* cost normalisation */
    sym_triu_col_norms(n, w->Pp, w->Pi, w->Px, Dt);
    double cm = 0; for (int i = 0; i < n; i++) cm += Dt[i];
    cm = n > 0 ? cm / n : 0;
    double qn = vec_norm_inf(w->q, n);
    limit_scaling(&qn, 1);
    double ct = cm > qn ? cm : qn;
    limit_scaling(&ct, 1);
    ct = 1.0 / ct;
    for (int p = 0; p < w->Pp[n]; p++) w->Px[p] *= ct;
    for (int j = 0; j < n; j++) w->q[j] *= ct;
    w->c *= ct;
  }
  for (int i = 0; i < n; i++) w->Dinv[i] = 1.0 / w->D[i];
  for (int i = 0; i < m; i++) w->Einv[i] = 1.0 / w->E[i];
  w->cinv = 1.0 / w->c;
  for (int i = 0; i < m; i++) { w->l[i] *= w->E[i]; w->u[i] *= w->E[i]; }
  free(Dt); free(Et);
}

/* ------------------------------------------------ greedy minimum degree */
static int *min_degree_order(int N, const int *Kp, const int *Ki) {
  /* Kp/Ki: upper-triangular CSC pattern of the (unpermuted) KKT matrix. */
  int W = (N + 63) / 64;
  uint64_t *adj = calloc((size_t)N * W, sizeof(uint64_t));
  int *deg = malloc(sizeof(int) * N), *perm = malloc(sizeof(int) * N);
  char *gone = calloc(N, 1);
  for (int j = 0; j < N; j++)
    for (int p = Kp[j]; p < Kp[j + 1]; p++) {
      int i = Ki[p]; if (i == j) continue;
      adj[(size_t)i * W + j / 64] |= 1ull << (j % 64);
      adj[(size_t)j * W + i / 64] |= 1ull << (i % 64);
    }
  for (int v = 0; v < N; v++) {
    int d = 0; for (int k = 0; k < W; k++) d += __builtin_popcountll(adj[(size_t)v * W + k]);
    deg[v] = d;
  }
  for (int step = 0; step < N; step++) {
    int v = -1, best = N + 1;
    for (int i = 0; i < N; i++) if (!gone[i] && deg[i] < best) { best = deg[i]; v = i; }
    perm[step] = v; gone[v] = 1;
    uint64_t *av = adj + (size_t)v * W;
    for (int k = 0; k < W; k++) {
      uint64_t bits = av[k];
      while (bits) {
        int u = k * 64 + __builtin_ctzll(bits); bits &= bits - 1;
        uint64_t *au = adj + (size_t)u * W;
        int d = 0;
        for (int kk = 0; kk < W; kk++) { au[kk] |= av[kk]; }
        au[u / 64] &= ~(1ull << (u % 64));
        au[v / 64] &= ~(1ull << (v % 64));
        for (int kk = 0; kk < W; kk++) d += __builtin_popcountll(au[kk]);
        deg[u] = d;
      }
    }
  }
  free(adj); free(deg); free(gone);
  return perm;
}

/* -------------------------------------------- KKT assembly + LDL^T factor */
static int factor_kkt(OracleWork *w) {
  int n = w->n, m = w->m, N = n + m;
  w->N = N;
  /* row-wise view of A (CSR) = columns n..N-1 of the upper-triangular KKT */
  int *Rp = calloc(m + 1, sizeof(int));
  for (int p = 0; p < w->Ap[n]; p++) Rp[w->Ai[p] + 1]++;
  for (int i = 0; i < m; i++) Rp[i + 1] += Rp[i];
  int nnzA = w->Ap[n];
  int *Rj = malloc(sizeof(int) * (nnzA > 0 ? nnzA : 1));
  double *Rx = malloc(sizeof(double) * (nnzA > 0 ? nnzA : 1));
  int *cur = malloc(sizeof(int) * (m > 0 ? m : 1));
  for (int i = 0; i < m; i++) cur[i] = Rp[i];
  for (int j = 0; j < n; j++)
    for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) { int i = w->Ai[p]; Rj[cur[i]] = j; Rx[cur[i]++] = w->Ax[p]; }
  /* unpermuted upper-triangular KKT in CSC: [[P+sigma I, A'],[., -1/rho]] */
  int cap = w->Pp[n] + n + nnzA + m;
  int *Kp = malloc(sizeof(int) * (N + 1)), *Ki = malloc(sizeof(int) * cap);
  double *Kx = malloc(sizeof(double) * cap);
  int nz = 0;
  for (int j = 0; j < n; j++) {
    Kp[j] = nz; int have_diag = 0;
    for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) {
      int i = w->Pi[p]; if (i > j) continue;
      Ki[nz] = i; Kx[nz] = w->Px[p];
      if (i == j) { Kx[nz] += w->s.sigma; have_diag = 1; }
      nz++;
    }
    if (!have_diag) { Ki[nz] = j; Kx[nz++] = w->s.sigma; }
  }
  for (int i = 0; i < m; i++) {
    Kp[n + i] = nz;
    for (int p = Rp[i]; p < Rp[i + 1]; p++) { Ki[nz] = Rj[p]; Kx[nz++] = Rx[p]; }
    Ki[nz] = n + i; Kx[nz++] = -w->rho_inv_vec[i];
  }
  Kp[N] = nz;
  free(Rp); free(Rj); free(Rx); free(cur);
  /* ordering */
  w->perm = min_degree_order(N, Kp, Ki);
  int *pinv = malloc(sizeof(int) * N);
  for (int k = 0; k < N; k++) pinv[w->perm[k]] = k;
  /* symmetric permutation into upper-triangular CSC */
  int *Cp = calloc(N + 1, sizeof(int)), *Ci = malloc(sizeof(int) * nz);
  double *Cx = malloc(sizeof(double) * nz);
  for (int j = 0; j < N; j++)
    for (int p = Kp[j]; p < Kp[j + 1]; p++) {
      int a = pinv[Ki[p]], b = pinv[j]; int col = a > b ? a : b; Cp[col + 1]++;
    }
  for (int j = 0; j < N; j++) Cp[j + 1] += Cp[j];
  int *cc = malloc(sizeof(int) * N); for (int j = 0; j < N; j++) cc[j] = Cp[j];
  for (int j = 0; j < N; j++)
    for (int p = Kp[j]; p < Kp[j + 1]; p++) {
      int a = pinv[Ki[p]], b = pinv[j]; int col = a > b ? a : b, row = a > b ? b : a;
      Ci[cc[col]] = row; Cx[cc[col]++] = Kx[p];
    }
  free(cc); free(Kp); free(Ki); free(Kx); free(pinv);
  /* elimination tree + column counts (Liu) */
  int *parent = malloc(sizeof(int) * N), *Lnz = calloc(N, sizeof(int)), *flag = malloc(sizeof(int) * N);
  for (int k = 0; k < N; k++) {
    parent[k] = -1; flag[k] = k;
    for (int p = Cp[k]; p < Cp[k + 1]; p++) {
      int i = Ci[p];
      for (; i < k && flag[i] != k; i = parent[i]) {
        if (parent[i] == -1) parent[i] = k;
        Lnz[i]++; flag[i] = k;
      }
    }
  }
  w->Lp = malloc(sizeof(int) * (N + 1)); w->Lp[0] = 0;
  for (int k = 0; k < N; k++) w->Lp[k + 1] = w->Lp[k] + Lnz[k];
  int nnzL = w->Lp[N];
  w->Li = malloc(sizeof(int) * (nnzL > 0 ? nnzL : 1));
  w->Lx = malloc(sizeof(double) * (nnzL > 0 ? nnzL : 1));
  w->Dd = malloc(sizeof(double) * N); w->Ddinv = malloc(sizeof(double) * N);
  w->etree = parent;
  /* up-looking numeric factorisation */
  double *Y = calloc(N, sizeof(double)); int *pattern = malloc(sizeof(int) * N);
  int ok = 1;
  for (int k = 0; k < N; k++) {
    int top = N; flag[k] = k; Lnz[k] = 0; Y[k] = 0;
    for (int p = Cp[k]; p < Cp[k + 1]; p++) {
      int i = Ci[p]; Y[i] += Cx[p];
      int len = 0;
      for (; flag[i] != k; i = parent[i]) { pattern[len++] = i; flag[i] = k; }
      while (len > 0) pattern[--top] = pattern[--len];
    }
    double dk = Y[k]; Y[k] = 0;
    for (; top < N; top++) {
      int i = pattern[top]; double yi = Y[i]; Y[i] = 0;
      int p2 = w->Lp[i] + Lnz[i];
      for (int p = w->Lp[i]; p < p2; p++) Y[w->Li[p]] -= w->Lx[p] * yi;
      double lki = yi * w->Ddinv[i];
      dk -= lki * yi;
      w->Li[p2] = k; w->Lx[p2] = lki; Lnz[i]++;
    }
    if (dk == 0.0) { ok = 0; dk = 1.0; }
    w->Dd[k] = dk; w->Ddinv[k] = 1.0 / dk;
  }
  free(Y); free(pattern); free(Lnz); free(flag); free(Cp); free(Ci); free(Cx);
  return ok ? 0 : -1;
}

/* solve K sol = b in place (b in original [x;z] order); tmp has length N */
static void kkt_solve(const OracleWork *w, double *b, double *tmp) {
  int N = w->N;
  const int *Lp = w->Lp, *Li = w->Li; const double *Lx = w->Lx;
  for (int k = 0; k < N; k++) tmp[k] = b[w->perm[k]];
  for (int j = 0; j < N; j++) {
    double v = tmp[j];
    for (int p = Lp[j]; p < Lp[j + 1]; p++) tmp[Li[p]] -= Lx[p] * v;
  }
  for (int j = 0; j < N; j++) tmp[j] *= w->Ddinv[j];
  for (int j = N - 1; j >= 0; j--) {
    double v = tmp[j];
    for (int p = Lp[j]; p < Lp[j + 1]; p++) v -= Lx[p] * tmp[Li[p]];
    tmp[j] = v;
  }
  for (int k = 0; k < N; k++) b[w->perm[k]] = tmp[k];
}

/* ---------------------------------------------------------- mat-vec helpers */
static void mat_vec_A(const OracleWork *w, const double *x, double *out) {        /* out = A x */
  for (int i = 0; i < w->m; i++) out[i] = 0;
  for (int j = 0; j < w->n; j++) { double xj = x[j];
    for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) out[w->Ai[p]] += w->Ax[p] * xj; }
}
static void mat_tvec_A(const OracleWork *w, const double *y, double *out) {       /* out = A' y */
  for (int j = 0; j < w->n; j++) { double s = 0;
    for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) s += w->Ax[p] * y[w->Ai[p]];
    out[j] = s; }
}
static void mat_vec_Psym(const OracleWork *w, const double *x, double *out) {     /* out = P x, P from triu */
  for (int j = 0; j < w->n; j++) out[j] = 0;
  for (int j = 0; j < w->n; j++)
    for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) {
      int i = w->Pi[p]; double v = w->Px[p];
      out[i] += v * x[j];
      if (i != j) out[j] += v * x[i];
    }
}

/* --------------------------------------------------------------- lifecycle */
void oracle_free(OracleWork *w) {
  if (!w) return;
  free(w->Pp); free(w->Pi); free(w->Px); free(w->Ap); free(w->Ai); free(w->Ax);
  free(w->q); free(w->l); free(w->u); free(w->D); free(w->Dinv); free(w->E); free(w->Einv);
  free(w->rho_vec); free(w->rho_inv_vec); free(w->ctype0); free(w->perm); free(w->Lp); free(w->Li); free(w->Lx);
  free(w->Dd); free(w->Ddinv); free(w->etree); free(w->x); free(w->y); free(w->z);
  free(w);
}

static void *dup_mem(const void *src, size_t bytes) {
  void *p = malloc(bytes ? bytes : 1); if (bytes) memcpy(p, src, bytes); return p;
}

/* P: upper-triangular CSC (n x n); A: CSC (m x n).  Returns NULL on failure. */
OracleWork *oracle_setup(int n, int m, const int *Pp, const int *Pi, const double *Px, const double *q,
                         const int *Ap, const int *Ai, const double *Ax, const double *l, const double *u,
                         const OracleSettings *s) {
  for (int i = 0; i < m; i++) if (l[i] > u[i]) return NULL;
  OracleWork *w = calloc(1, sizeof(OracleWork));
  w->n = n; w->m = m; w->s = *s;
  w->Pp = dup_mem(Pp, sizeof(int) * (n + 1)); w->Pi = dup_mem(Pi, sizeof(int) * Pp[n]); w->Px = dup_mem(Px, sizeof(double) * Pp[n]);
  w->Ap = dup_mem(Ap, sizeof(int) * (n + 1)); w->Ai = dup_mem(Ai, sizeof(int) * Ap[n]); w->Ax = dup_mem(Ax, sizeof(double) * Ap[n]);
  w->q = dup_mem(q, sizeof(double) * n);
  w->l = malloc(sizeof(double) * (m ? m : 1)); w->u = malloc(sizeof(double) * (m ? m : 1));
  for (int i = 0; i < m; i++) {   /* python wrapper clamps to +-OSQP_INFTY */
    w->l[i] = l[i] < -OSQP_INFTY ? -OSQP_INFTY : l[i];
    w->u[i] = u[i] > OSQP_INFTY ? OSQP_INFTY : u[i];
  }
  w->D = malloc(sizeof(double) * (n ? n : 1)); w->Dinv = malloc(sizeof(double) * (n ? n : 1));
  w->E = malloc(sizeof(double) * (m ? m : 1)); w->Einv = malloc(sizeof(double) * (m ? m : 1));
  scale_data(w);
  w->rho_vec = malloc(sizeof(double) * (m ? m : 1)); w->rho_inv_vec = malloc(sizeof(double) * (m ? m : 1));
  w->ctype0 = calloc(m ? m : 1, sizeof(int));
  for (int i = 0; i < m; i++) {
    double r = s->rho;
    if (s->eq_rho) {
      if (w->l[i] < -OSQP_INFTY * MIN_SCALING && w->u[i] > OSQP_INFTY * MIN_SCALING) { r = RHO_MIN; w->ctype0[i] = -1; }
      else if (w->u[i] - w->l[i] < RHO_TOL) { r = RHO_EQ_OVER_RHO_INEQ * s->rho; w->ctype0[i] = 1; }
    }
    w->rho_vec[i] = r; w->rho_inv_vec[i] = 1.0 / r;
  }
  if (factor_kkt(w) != 0) { oracle_free(w); return NULL; }
  w->x = calloc(n ? n : 1, sizeof(double)); w->y = calloc(m ? m : 1, sizeof(double)); w->z = calloc(m ? m : 1, sizeof(double));
  return w;
}

int oracle_update_lin_cost(OracleWork *w, const double *q) {          /* osqp update(q=) */
  for (int j = 0; j < w->n; j++) w->q[j] = w->c * w->D[j] * q[j];
  return 0;
}
int oracle_update_bounds(OracleWork *w, const double *l, const double *u) {   /* osqp update(l=,u=) */
  for (int i = 0; i < w->m; i++) if (l[i] > u[i]) return 1;
  for (int i = 0; i < w->m; i++) {
    double li = l[i] < -OSQP_INFTY ? -OSQP_INFTY : l[i], ui = u[i] > OSQP_INFTY ? OSQP_INFTY : u[i];
    w->l[i] = w->E[i] * li; w->u[i] = w->E[i] * ui;
  }
  return 0;
}
int oracle_warm_start(OracleWork *w, const double *x, const double *y) {      /* osqp warm_start(x=,y=) */
  for (int j = 0; j < w->n; j++) w->x[j] = w->Dinv[j] * x[j];
  for (int i = 0; i < w->m; i++) w->y[i] = w->c * w->Einv[i] * y[i];
  mat_vec_A(w, w->x, w->z);
  return 0;
}

/* ----------------------------------------------------- the ADMM loop (A.4-6) */
typedef struct {
  double *x, *z, *y, *x_prev, *z_prev, *xz, *tmp, *dx, *dy, *Ax, *Px, *Aty, *wn, *wm;
} Scratch;
static Scratch scratch_new(int n, int m) {
  Scratch s; int N = n + m;
  s.x = calloc(n + 1, 8); s.z = calloc(m + 1, 8); s.y = calloc(m + 1, 8);
  s.x_prev = calloc(n + 1, 8); s.z_prev = calloc(m + 1, 8);
  s.xz = calloc(N + 1, 8); s.tmp = calloc(N + 1, 8);
  s.dx = calloc(n + 1, 8); s.dy = calloc(m + 1, 8);
  s.Ax = calloc(m + 1, 8); s.Px = calloc(n + 1, 8); s.Aty = calloc(n + 1, 8);
  s.wn = calloc(n + 1, 8); s.wm = calloc(m + 1, 8);
  return s;
}
static void scratch_free(Scratch *s) {
  free(s->x); free(s->z); free(s->y); free(s->x_prev); free(s->z_prev); free(s->xz); free(s->tmp);
  free(s->dx); free(s->dy); free(s->Ax); free(s->Px); free(s->Aty); free(s->wn); free(s->wm);
}

typedef struct { double pri_res, dua_res, obj; } Resid;

static void update_info(const OracleWork *w, Scratch *s, const double *q, Resid *r, double *nAx, double *nz,
                        double *nPx, double *nAty, double *nq) {
  int n = w->n, m = w->m;
  int unscaled = w->s.scaling && !w->s.scaled_termination;
  mat_vec_A(w, s->x, s->Ax); mat_vec_Psym(w, s->x, s->Px); mat_tvec_A(w, s->y, s->Aty);
  double pr = 0, a1 = 0, a2 = 0;
  for (int i = 0; i < m; i++) {
    double e = unscaled ? w->Einv[i] : 1.0;
    double d = fabs(e * (s->Ax[i] - s->z[i])); if (d > pr) pr = d;
    double a = fabs(e * s->Ax[i]); if (a > a1) a1 = a;
    double b = fabs(e * s->z[i]); if (b > a2) a2 = b;
  }
  double dr = 0, b1 = 0, b2 = 0, b3 = 0, quad = 0, lin = 0;
  for (int j = 0; j < n; j++) {
    double d = unscaled ? w->Dinv[j] : 1.0;
    double v = fabs(d * (s->Px[j] + q[j] + s->Aty[j])); if (v > dr) dr = v;
    double p = fabs(d * s->Px[j]); if (p > b1) b1 = p;
    double a = fabs(d * s->Aty[j]); if (a > b2) b2 = a;
    double c = fabs(d * q[j]); if (c > b3) b3 = c;
    quad += s->x[j] * s->Px[j]; lin += q[j] * s->x[j];
  }
  double ci = unscaled ? w->cinv : 1.0;
  r->pri_res = pr; r->dua_res = ci * dr;
  *nAx = a1; *nz = a2; *nPx = ci * b1; *nAty = ci * b2; *nq = ci * b3;
  r->obj = (0.5 * quad + lin) * (w->s.scaling ? w->cinv : 1.0);
}

static int is_primal_infeasible(const OracleWork *w, Scratch *s, const double *l, const double *u, double eps) {
  int n = w->n, m = w->m; int unscaled = w->s.scaling && !w->s.scaled_termination;
  double nrm = 0, lhs = 0;
  for (int i = 0; i < m; i++) {
    double d = s->dy[i];
    if (u[i] > OSQP_INFTY * MIN_SCALING) {
      if (l[i] < -OSQP_INFTY * MIN_SCALING) d = 0.0; else d = d < 0 ? d : 0.0;
    } else if (l[i] < -OSQP_INFTY * MIN_SCALING) d = d > 0 ? d : 0.0;
    s->wm[i] = d;
    double a = fabs((unscaled ? w->E[i] : 1.0) * d); if (a > nrm) nrm = a;
  }
  if (nrm > DIVISION_TOL) {
    for (int i = 0; i < m; i++) {
      double d = s->wm[i];
      lhs += u[i] * (d > 0 ? d : 0.0) + l[i] * (d < 0 ? d : 0.0);
    }
    if (lhs < -eps * nrm) {
      mat_tvec_A(w, s->wm, s->wn);
      double t = 0;
      for (int j = 0; j < n; j++) { double a = fabs((unscaled ? w->Dinv[j] : 1.0) * s->wn[j]); if (a > t) t = a; }
      return t < eps * nrm;
    }
  }
  return 0;
}

static int is_dual_infeasible(const OracleWork *w, Scratch *s, const double *q, const double *l, const double *u, double eps) {
  int n = w->n, m = w->m; int unscaled = w->s.scaling && !w->s.scaled_termination;
  double nrm = 0, cs = unscaled ? w->c : 1.0, qdx = 0;
  for (int j = 0; j < n; j++) { double a = fabs((unscaled ? w->D[j] : 1.0) * s->dx[j]); if (a > nrm) nrm = a; qdx += q[j] * s->dx[j]; }
  if (nrm > DIVISION_TOL) {
    if (qdx < -cs * eps * nrm) {
      mat_vec_Psym(w, s->dx, s->wn);
      double t = 0;
      for (int j = 0; j < n; j++) { double a = fabs((unscaled ? w->Dinv[j] : 1.0) * s->wn[j]); if (a > t) t = a; }
      if (t < cs * eps * nrm) {
        mat_vec_A(w, s->dx, s->wm);
        for (int i = 0; i < m; i++) {
          double v = (unscaled ? w->Einv[i] : 1.0) * s->wm[i];
          if ((u[i] < OSQP_INFTY * MIN_SCALING && v > eps * nrm) || (l[i] > -OSQP_INFTY * MIN_SCALING && v < -eps * nrm)) return 0;
        }
        return 1;
      }
    }
  }
  return 0;
}

/* returns 1 and sets *status when a termination condition holds */
static int check_termination(const OracleWork *w, Scratch *s, const double *q, const double *l, const double *u,
                             const Resid *r, double nAx, double nz, double nPx, double nAty, double nq,
                             int approximate, int *status) {
  double k = approximate ? 10.0 : 1.0;
  double eps_abs = w->s.eps_abs * k, eps_rel = w->s.eps_rel * k;
  double eps_pinf = w->s.eps_prim_inf * k, eps_dinf = w->s.eps_dual_inf * k;
  if (r->pri_res > OSQP_INFTY || r->dua_res > OSQP_INFTY) { *status = ST_NON_CVX; return 1; }
  int prim_ok = 0, dual_ok = 0, pinf = 0, dinf = 0;
  if (w->m == 0) prim_ok = 1;
  else {
    double eps_prim = eps_abs + eps_rel * (nAx > nz ? nAx : nz);
    if (r->pri_res < eps_prim) prim_ok = 1; else pinf = is_primal_infeasible(w, s, l, u, eps_pinf);
  }
  double mx = nPx > nAty ? nPx : nAty; if (nq > mx) mx = nq;
  double eps_dual = eps_abs + eps_rel * mx;
  if (r->dua_res < eps_dual) dual_ok = 1; else dinf = is_dual_infeasible(w, s, q, l, u, eps_dinf);
  if (prim_ok && dual_ok) { *status = approximate ? ST_SOLVED_INACC : ST_SOLVED; return 1; }
  if (pinf) { *status = approximate ? ST_PINF_INACC : ST_PINF; return 1; }
  if (dinf) { *status = approximate ? ST_DINF_INACC : ST_DINF; return 1; }
  return 0;
}

/*
 * One node = update(l,u) + warm_start(x0,y0) + solve(), as a PURE function
 * (node.py:96-125).  l,u,x0,y0 and the outputs are UNSCALED.  q_scaled is the
 * current scaled linear cost.  Returns 1 if l>u somewhere (osqp raises).
 */
static int solve_node(const OracleWork *w, Scratch *s, const double *l_in, const double *u_in,
                      const double *x0, const double *y0, double *x_out, double *y_out, OracleInfo *info) {
  int n = w->n, m = w->m; const OracleSettings *S = &w->s;
  struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
  info->rho_updates = 0; info->rho_final = S->rho;
  double *l = malloc(8 * (m + 1)), *u = malloc(8 * (m + 1));
  for (int i = 0; i < m; i++) {
    if (l_in[i] > u_in[i]) { free(l); free(u); return 1; }
    double li = l_in[i] < -OSQP_INFTY ? -OSQP_INFTY : l_in[i], ui = u_in[i] > OSQP_INFTY ? OSQP_INFTY : u_in[i];
    l[i] = w->E[i] * li; u[i] = w->E[i] * ui;
  }
  /* eq_rho == 2: what osqp >= 0.4 does in update_bounds -- re-type every row from THIS node's scaled
   * bounds and, if any type changed against the factor at hand, rebuild rho_vec and refactor numerically.  Not the
   * engine's parity contract (eq_rho == 1: typed once at setup); kept to measure what that contract costs (section 8f2). */
  OracleWork tw; double *rv = NULL, *ri = NULL; int refactored = 0;
  if (S->eq_rho == 2 && m > 0) {
    rv = malloc(8 * (size_t)m); ri = malloc(8 * (size_t)m);
    int changed = 0;
    for (int i = 0; i < m; i++) {
      double r = S->rho;
      if (l[i] < -OSQP_INFTY * MIN_SCALING && u[i] > OSQP_INFTY * MIN_SCALING) r = RHO_MIN;
      else if (u[i] - l[i] < RHO_TOL) r = RHO_EQ_OVER_RHO_INEQ * S->rho;
      rv[i] = r; ri[i] = 1.0 / r;
      if (r != w->rho_vec[i]) changed = 1;
    }
    if (changed) {
      tw = *w; tw.rho_vec = rv; tw.rho_inv_vec = ri;
      tw.perm = NULL; tw.Lp = NULL; tw.Li = NULL; tw.Lx = NULL; tw.Dd = NULL; tw.Ddinv = NULL; tw.etree = NULL;
      if (factor_kkt(&tw) != 0) {
        free(tw.perm); free(tw.Lp); free(tw.Li); free(tw.Lx); free(tw.Dd); free(tw.Ddinv); free(tw.etree);
        free(rv); free(ri); free(l); free(u); return 2;
      }
      w = &tw; S = &w->s; refactored = 1;
    }
  }
  /* adaptive rho: constraint type per row (osqp constr_type: -1 loose, 1 equality, 0 inequality) from the bounds the typing
   * in effect looked at -- the root's at setup (eq_rho 1), this node's (eq_rho 2), none (eq_rho 0) */
  const int adapt = S->adaptive_rho && S->adaptive_rho_interval > 0 && m > 0;
  int *ctype = NULL; double rho_cur = S->rho;
  if (adapt) {
    ctype = calloc((size_t)m, sizeof(int));
    for (int i = 0; i < m && S->eq_rho; i++) {
      if (S->eq_rho != 2) { ctype[i] = w->ctype0[i]; continue; }     /* the stored bounds follow update(l, u): use the setup's typing */
      if (l[i] < -OSQP_INFTY * MIN_SCALING && u[i] > OSQP_INFTY * MIN_SCALING) ctype[i] = -1;
      else if (u[i] - l[i] < RHO_TOL) ctype[i] = 1;
    }
    if (!rv) { rv = malloc(8 * (size_t)m); ri = malloc(8 * (size_t)m); memcpy(rv, w->rho_vec, 8 * (size_t)m); memcpy(ri, w->rho_inv_vec, 8 * (size_t)m); }
  }
  for (int j = 0; j < n; j++) s->x[j] = w->Dinv[j] * x0[j];
  for (int i = 0; i < m; i++) s->y[i] = w->c * w->Einv[i] * y0[i];
  mat_vec_A(w, s->x, s->z);
  const double *q = w->q; double alpha = S->alpha;
  int status = ST_UNSOLVED, iter, checked = 0;
  Resid r = {0, 0, 0}; double nAx = 0, nz = 0, nPx = 0, nAty = 0, nq = 0;
  for (iter = 1; iter <= S->max_iter; iter++) {
    double *t;
    t = s->x; s->x = s->x_prev; s->x_prev = t;
    t = s->z; s->z = s->z_prev; s->z_prev = t;
    for (int j = 0; j < n; j++) s->xz[j] = S->sigma * s->x_prev[j] - q[j];
    for (int i = 0; i < m; i++) s->xz[n + i] = s->z_prev[i] - w->rho_inv_vec[i] * s->y[i];
    kkt_solve(w, s->xz, s->tmp);
    for (int i = 0; i < m; i++) s->xz[n + i] = s->z_prev[i] + w->rho_inv_vec[i] * (s->xz[n + i] - s->y[i]);
    for (int j = 0; j < n; j++) { s->x[j] = alpha * s->xz[j] + (1.0 - alpha) * s->x_prev[j]; s->dx[j] = s->x[j] - s->x_prev[j]; }
    for (int i = 0; i < m; i++) {
      double zr = alpha * s->xz[n + i] + (1.0 - alpha) * s->z_prev[i];
      double zi = zr + w->rho_inv_vec[i] * s->y[i];
      zi = zi < l[i] ? l[i] : (zi > u[i] ? u[i] : zi);
      s->z[i] = zi;
      s->dy[i] = w->rho_vec[i] * (zr - zi);
      s->y[i] += s->dy[i];
    }
    checked = S->check_termination && (iter % S->check_termination == 0);
    if (checked) {
      update_info(w, s, q, &r, &nAx, &nz, &nPx, &nAty, &nq);
      if (check_termination(w, s, q, l, u, &r, nAx, nz, nPx, nAty, nq, 0, &status)) break;
    }
    if (adapt && iter % S->adaptive_rho_interval == 0) {
      /* osqp_solve: update_info if this iteration had no check, then adapt_rho -> compute_rho_estimate on the SCALED
       * residuals and norms (plain inf-norms of the solver's own vectors), update when outside [rho / tol, rho * tol] */
      if (!checked) update_info(w, s, q, &r, &nAx, &nz, &nPx, &nAty, &nq);
      double pr = 0, n1 = 0, n2 = 0, dr = 0, d1 = 0, d2 = 0, d3 = 0;
      for (int i = 0; i < m; i++) {
        double a = fabs(s->Ax[i] - s->z[i]); if (a > pr) pr = a;
        a = fabs(s->z[i]); if (a > n1) n1 = a;
        a = fabs(s->Ax[i]); if (a > n2) n2 = a;
      }
      for (int j = 0; j < n; j++) {
        double a = fabs(s->Px[j] + q[j] + s->Aty[j]); if (a > dr) dr = a;
        a = fabs(q[j]); if (a > d1) d1 = a;
        a = fabs(s->Aty[j]); if (a > d2) d2 = a;
        a = fabs(s->Px[j]); if (a > d3) d3 = a;
      }
      pr /= (fmax(n1, n2) + 1e-10);
      dr /= (fmax(fmax(d1, d2), d3) + 1e-10);
      double rho_new = rho_cur * sqrt(pr / (dr + 1e-10));
      rho_new = fmin(fmax(rho_new, RHO_MIN), RHO_MAX);
      if (rho_new > rho_cur * S->adaptive_rho_tolerance || rho_new < rho_cur / S->adaptive_rho_tolerance) {
        for (int i = 0; i < m; i++) {
          if (ctype[i] == 0) rv[i] = rho_new; else if (ctype[i] == 1) rv[i] = RHO_EQ_OVER_RHO_INEQ * rho_new;
          ri[i] = 1.0 / rv[i];
        }
        OracleWork nw = *w; nw.rho_vec = rv; nw.rho_inv_vec = ri;
        nw.perm = NULL; nw.Lp = NULL; nw.Li = NULL; nw.Lx = NULL; nw.Dd = NULL; nw.Ddinv = NULL; nw.etree = NULL;
        int frc = factor_kkt(&nw);
        if (refactored) { free(tw.perm); free(tw.Lp); free(tw.Li); free(tw.Lx); free(tw.Dd); free(tw.Ddinv); free(tw.etree); refactored = 0; }
        if (frc != 0) {
          free(nw.perm); free(nw.Lp); free(nw.Li); free(nw.Lx); free(nw.Dd); free(nw.Ddinv); free(nw.etree);
          free(rv); free(ri); free(l); free(u); free(ctype); return 2;
        }
        tw = nw; w = &tw; S = &w->s; refactored = 1; rho_cur = rho_new;
        info->rho_updates++;
      }
    }
  }
  if (iter > S->max_iter) iter = S->max_iter;
  if (status == ST_UNSOLVED) {
    if (!checked) {
      update_info(w, s, q, &r, &nAx, &nz, &nPx, &nAty, &nq);
      check_termination(w, s, q, l, u, &r, nAx, nz, nPx, nAty, nq, 0, &status);
    }
    if (status == ST_UNSOLVED && !check_termination(w, s, q, l, u, &r, nAx, nz, nPx, nAty, nq, 1, &status))
      status = ST_MAX_ITER;
  }
  info->status = status; info->iter = iter; info->pri_res = r.pri_res; info->dua_res = r.dua_res;
  int infeas = (status == ST_PINF || status == ST_PINF_INACC || status == ST_DINF || status == ST_DINF_INACC || status == ST_NON_CVX);
  if (status == ST_PINF || status == ST_PINF_INACC) info->obj_val = OSQP_INFTY;
  else if (status == ST_DINF || status == ST_DINF_INACC) info->obj_val = -OSQP_INFTY;
  else if (status == ST_NON_CVX) info->obj_val = NAN;
  else info->obj_val = r.obj;
  for (int j = 0; j < n; j++) x_out[j] = infeas ? NAN : w->D[j] * s->x[j];
  for (int i = 0; i < m; i++) y_out[i] = infeas ? NAN : w->cinv * w->E[i] * s->y[i];
  info->rho_final = rho_cur;
  free(l); free(u); free(ctype);
  if (refactored) { free(tw.perm); free(tw.Lp); free(tw.Li); free(tw.Lx); free(tw.Dd); free(tw.Ddinv); free(tw.etree); }
  free(rv); free(ri);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  info->solve_time = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  return 0;
}

/* sequential osqp-object semantics: uses the stored (scaled) l,u and warm start */
int oracle_solve(OracleWork *w, double *x_out, double *y_out, OracleInfo *info) {
  int n = w->n, m = w->m;
  Scratch s = scratch_new(n, m);
  double *l = malloc(8 * (m + 1)), *u = malloc(8 * (m + 1)), *x0 = malloc(8 * (n + 1)), *y0 = malloc(8 * (m + 1));
  for (int i = 0; i < m; i++) { l[i] = w->Einv[i] * w->l[i]; u[i] = w->Einv[i] * w->u[i]; y0[i] = w->cinv * w->E[i] * w->y[i]; }
  for (int j = 0; j < n; j++) x0[j] = w->D[j] * w->x[j];
  int rc = solve_node(w, &s, l, u, x0, y0, x_out, y_out, info);
  scratch_free(&s); free(l); free(u); free(x0); free(y0);
  return rc;
}

/* stateless single node (pure function of l,u,x0,y0) */
int oracle_solve_node(const OracleWork *w, const double *l, const double *u, const double *x0, const double *y0,
                      double *x, double *y, OracleInfo *info) {
  Scratch s = scratch_new(w->n, w->m);
  int rc = solve_node(w, &s, l, u, x0, y0, x, y, info);
  scratch_free(&s);
  return rc;
}

/*
 * Batches of independent nodes on `threads` host threads (CPU baseline leg):
 * a pthread pool pulling node indices from an atomic counter.  Node b uses
 * workspace ws[b] (the same pointer repeated for nodes of one instance).
 */
typedef struct {
  int B; OracleWork *const *ws; const double *const *l, *const *u, *const *x0, *const *y0;
  double *const *x, *const *y; OracleInfo *infos; int next, bad;
} BatchJob;

static void *batch_worker(void *arg) {
  BatchJob *job = (BatchJob *)arg;
  for (;;) {
    int b = __atomic_fetch_add(&job->next, 1, __ATOMIC_RELAXED);
    if (b >= job->B) break;
    Scratch s = scratch_new(job->ws[b]->n, job->ws[b]->m);
    int rc = solve_node(job->ws[b], &s, job->l[b], job->u[b], job->x0[b], job->y0[b], job->x[b], job->y[b], job->infos + b);
    scratch_free(&s);
    if (rc) __atomic_store_n(&job->bad, 1, __ATOMIC_RELAXED);
  }
  return NULL;
}

int oracle_solve_multi(int B, OracleWork *const *ws, const double *const *l, const double *const *u,
                       const double *const *x0, const double *const *y0, double *const *x, double *const *y,
                       OracleInfo *infos, int threads) {
  BatchJob job = {B, ws, l, u, x0, y0, x, y, infos, 0, 0};
  if (threads < 1) threads = 1;
  if (threads > B) threads = B;
  if (threads <= 1) { batch_worker(&job); return job.bad; }
  pthread_t *tid = malloc(sizeof(pthread_t) * threads);
  for (int t = 0; t < threads; t++) pthread_create(&tid[t], NULL, batch_worker, &job);
  for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
  free(tid);
  return job.bad;
}

/* nodes of ONE instance; row-major arrays: l,u,y0,y are [B][m]; x0,x are [B][n] */
int oracle_solve_batch(OracleWork *w, int B, const double *l, const double *u, const double *x0,
                       const double *y0, double *x, double *y, OracleInfo *infos, int threads) {
  int n = w->n, m = w->m;
  void **p = malloc(sizeof(void *) * 7 * (B + 1));
  OracleWork **ws = (OracleWork **)p;
  const double **pl = (const double **)(p + B), **pu = (const double **)(p + 2 * B), **px0 = (const double **)(p + 3 * B),
               **py0 = (const double **)(p + 4 * B);
  double **px = (double **)(p + 5 * B), **py = (double **)(p + 6 * B);
  for (int b = 0; b < B; b++) {
    ws[b] = w; pl[b] = l + (size_t)b * m; pu[b] = u + (size_t)b * m; px0[b] = x0 + (size_t)b * n; py0[b] = y0 + (size_t)b * m;
    px[b] = x + (size_t)b * n; py[b] = y + (size_t)b * m;
  }
  int rc = oracle_solve_multi(B, ws, pl, pu, px0, py0, px, py, infos, threads);
  free(p);
  return rc;
}

/* introspection for tests */
int oracle_dims(const OracleWork *w, int *n, int *m, int *N, int *nnzL) { *n = w->n; *m = w->m; *N = w->N; *nnzL = w->Lp[w->N]; return 0; }
void oracle_get_scaling(const OracleWork *w, double *D, double *E, double *c) {
  memcpy(D, w->D, 8 * w->n); memcpy(E, w->E, 8 * w->m); *c = w->c;
}
void oracle_get_factor(const OracleWork *w, int *perm, int *Lp, int *Li, double *Lx, double *Dd) {
  memcpy(perm, w->perm, 4 * w->N); memcpy(Lp, w->Lp, 4 * (w->N + 1));
  memcpy(Li, w->Li, 4 * w->Lp[w->N]); memcpy(Lx, w->Lx, 8 * w->Lp[w->N]); memcpy(Dd, w->Dd, 8 * w->N);
}
/* test hook: solve K v = b with the factor (b in [x;z] order, length n+m) */
void oracle_kkt_solve(const OracleWork *w, double *b) { double *t = malloc(8 * (w->N + 1)); kkt_solve(w, b, t); free(t); }
