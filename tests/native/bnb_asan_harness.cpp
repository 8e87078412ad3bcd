// ASan/UBSan harness for bqp_bnb.cpp: separable QP  min 1/2 |x|^2 + q'x  with bounds on the integer variables only,
// so the exact relaxation is x = clip(-q) -- a fake solve function exercises branching, pruning, look-ahead and adoption.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "bqp.h"
static int N, NI; static std::vector<double> Q;
static int fake(void *, int B, const double *const *l, const double *const *u, const double *const *, const double *const *,
                double *const *x, double *const *y, int *status, int *iters) {
  for (int b = 0; b < B; b++) {
    for (int j = 0; j < N; j++) x[b][j] = -Q[j];
    for (int k = 0; k < NI; k++) { double v = -Q[k]; v = std::fmin(std::fmax(v, l[b][k]), u[b][k]); x[b][k] = v; y[b][k] = -(v + Q[k]); }
    status[b] = 1; iters[b] = 25;
  }
  return 0;
}
int main() {
  for (int trial = 0; trial < 200; trial++) {
    srand(trial);
    N = 3 + rand() % 8; NI = 1 + rand() % N;
    Q.assign(N, 0); for (auto &v : Q) v = 4.0 * rand() / RAND_MAX - 2.0;
    std::vector<int> Pp(N + 1), Pi(N), Ap(N + 1), Ai, idx(NI); std::vector<double> Px(N, 1.0), Ax, l(NI), u(NI);
    for (int j = 0; j <= N; j++) Pp[j] = j; for (int j = 0; j < N; j++) Pi[j] = j;
    for (int j = 0; j < N; j++) { Ap[j] = (int)Ai.size(); if (j < NI) { Ai.push_back(j); Ax.push_back(1.0); } } Ap[N] = (int)Ai.size();
    for (int k = 0; k < NI; k++) { idx[k] = k; l[k] = -2; u[k] = 2; }
    bqp_problem p{N, NI, Pp.data(), Pi.data(), Px.data(), Ap.data(), Ai.data(), Ax.data(), Q.data(), l.data(), u.data(), NI, idx.data()};
    double ref_obj = 0; int ref_iter = 0;
    for (int spec : {0, 2, 8, 64}) for (int rule : {0, 1}) for (int lim : {1000, 3}) {
      bqp_bnb_settings s{1e-3, lim, rule, 0, spec, 1e-3};
      std::vector<double> x(N); bqp_bnb_result r; std::vector<int> dec(2 * 1000);
      int rc = bqp_bnb_solve(nullptr, &p, &s, nullptr, INFINITY, fake, nullptr, x.data(), &r, dec.data(), 1000);
      if (rc) { printf("rc %d\n", rc); return 1; }
      if (lim == 1000) {
        // optimum of the separable problem: round each integer variable of clip(-q)
        double o = 0; for (int j = 0; j < N; j++) { double v = -Q[j]; if (j < NI) v = std::nearbyint(std::fmin(std::fmax(v, -2), 2)); o += 0.5 * v * v + Q[j] * v; }
        if (r.status != BQP_MI_SOLVED || std::fabs(r.upper_glob - o) > 5e-3) { printf("trial %d spec %d rule %d: status %d obj %g want %g\n", trial, spec, rule, r.status, r.upper_glob, o); return 1; }
        if (spec == 0 && rule == 0) { ref_obj = r.upper_glob; ref_iter = r.iter_num; }
        (void)ref_obj; (void)ref_iter;
      }
    }
  }
  printf("asan harness ok\n");
  return 0;
}
extern "C" int bqp_solve_multi(int, const bqp_handle *, const double *const *, const double *const *, const double *const *,
                               const double *const *, double *const *, double *const *, const bqp_node_out *) { return BQP_E_CUDA; }
// the engine entry points bqp_bnb.cpp links against (never reached here: the harness drives the replay through its own solve function)
extern "C" int bqp_ctx_solve_multi(bqp_ctx, int, const bqp_handle *, const double *const *, const double *const *, const double *const *,
                                   const double *const *, double *const *, double *const *, const bqp_node_out *) { return BQP_E_CUDA; }
extern "C" int bqp_ctx_create(int, int, bqp_ctx *) { return BQP_E_CUDA; }
extern "C" int bqp_ctx_free(bqp_ctx) { return BQP_OK; }
extern "C" int bqp_handle_device(bqp_handle) { return -1; }
extern "C" int bqp_session_begin(bqp_ctx) { return BQP_E_CUDA; }
extern "C" int bqp_session_append(bqp_ctx, int, const bqp_handle *, const double *const *, const double *const *, const double *const *,
                                  const double *const *, int *) { return BQP_E_CUDA; }
extern "C" int bqp_session_round(bqp_ctx, int *, int, int *, int *) { return BQP_E_CUDA; }
extern "C" int bqp_session_fetch(bqp_ctx, int, double *, double *, const bqp_node_out *) { return BQP_E_CUDA; }
extern "C" int bqp_ctx_set_sm_share(bqp_ctx, int) { return BQP_OK; }
