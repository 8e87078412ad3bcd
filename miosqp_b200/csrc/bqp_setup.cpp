// bqp_setup.cpp -- host side of bqp_setup(): what osqp.OSQP().setup() does once per (P, A)
// (/root/reference/miosqp/workspace.py:63-68), re-designed for the batched sm_100a kernel:
//   1. Ruiz equilibration + cost normalisation (OSQP paper, sec. 5.1)
//   2. per-row rho typing, ONCE, from the root bounds (parity contract: no per-node re-typing)
//   3. LDL^T of the quasi-definite KKT matrix in CONSTRAINTS-FIRST elimination order.  With the
//      (2,2) block -diag(1/rho) eliminated first there is no fill in the first m columns:
//        L = [[I, 0], [L21, L22]],  L21 = -A' diag(rho),  L22 D2 L22' = P + sigma I + A' diag(rho) A,
//      so the sparse columns of L are the rows of A (streamed as the A / A' panels) and the
//      trailing supernode L22 is dense (P = Pt Pt' is dense in every BASELINE config).
//   4. the streaming layouts the kernel reads: 32-row sliced panels for A, A', P and 32x32-blocked
//      column / row panels of L22 for the forward / backward sweeps.
// Nothing here calls into oracle/.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "bqp_internal.h"

namespace bqp {

namespace {

struct Csc {
  int rows = 0, cols = 0;
  std::vector<int> p, i;
  std::vector<double> x;
};

void limit_scaling(double *v, int n) {
  for (int k = 0; k < n; k++) {
    if (v[k] < kMinScaling) v[k] = 1.0;
    if (v[k] > kMaxScaling) v[k] = kMaxScaling;
  }
}

void sym_col_norms(const Csc &P, std::vector<double> &out) {
  std::fill(out.begin(), out.end(), 0.0);
  for (int j = 0; j < P.cols; j++)
    for (int k = P.p[j]; k < P.p[j + 1]; k++) {
      int r = P.i[k];
      double a = std::fabs(P.x[k]);
      out[j] = std::max(out[j], a);
      if (r != j) out[r] = std::max(out[r], a);
    }
}

// Ruiz equilibration of [[P, A'],[A, 0]] with cost normalisation; scales P, A, q in place.
void ruiz(Csc &P, Csc &A, std::vector<double> &q, int passes, HostInstance *h) {
  const int n = h->n, m = h->m;
  h->D.assign(n, 1.0);
  h->E.assign(m, 1.0);
  h->c = 1.0;
  std::vector<double> dt(std::max(n, 1)), et(std::max(m, 1));
  for (int it = 0; it < passes; it++) {
    sym_col_norms(P, dt);
    std::fill(et.begin(), et.end(), 0.0);
    for (int j = 0; j < n; j++) {
      double cn = 0;
      for (int k = A.p[j]; k < A.p[j + 1]; k++) {
        double a = std::fabs(A.x[k]);
        cn = std::max(cn, a);
        et[A.i[k]] = std::max(et[A.i[k]], a);
      }
      dt[j] = std::max(dt[j], cn);
    }
    limit_scaling(dt.data(), n);
    limit_scaling(et.data(), m);
    for (int j = 0; j < n; j++) dt[j] = 1.0 / std::sqrt(dt[j]);
    for (int r = 0; r < m; r++) et[r] = 1.0 / std::sqrt(et[r]);
    for (int j = 0; j < n; j++) {
      for (int k = P.p[j]; k < P.p[j + 1]; k++) P.x[k] *= dt[P.i[k]] * dt[j];
      for (int k = A.p[j]; k < A.p[j + 1]; k++) A.x[k] *= et[A.i[k]] * dt[j];
      q[j] *= dt[j];
      h->D[j] *= dt[j];
    }
    for (int r = 0; r < m; r++) h->E[r] *= et[r];
    sym_col_norms(P, dt);
    double mean = 0;
    for (int j = 0; j < n; j++) mean += dt[j];
    mean = n > 0 ? mean / n : 0.0;
    double qn = 0;
    for (int j = 0; j < n; j++) qn = std::max(qn, std::fabs(q[j]));
    limit_scaling(&qn, 1);
    double ct = std::max(mean, qn);
    limit_scaling(&ct, 1);
    ct = 1.0 / ct;
    for (auto &v : P.x) v *= ct;
    for (auto &v : q) v *= ct;
    h->c *= ct;
  }
  h->Dinv.resize(n);
  h->Einv.resize(m);
  for (int j = 0; j < n; j++) h->Dinv[j] = 1.0 / h->D[j];
  for (int r = 0; r < m; r++) h->Einv[r] = 1.0 / h->E[r];
  h->cinv = 1.0 / h->c;
}

// rows[r] = sorted (col, val) list -> sliced panel format
using RowList = std::vector<std::vector<std::pair<int, double>>>;

void build_panel(const RowList &rows, int ncols, HostMat *M) {
  const int nrows = (int)rows.size();
  M->rows = nrows;
  M->cols = ncols;
  M->nslices = (nrows + 31) / 32;
  M->sptr.assign(M->nslices + 1, 0);
  M->iptr.assign(M->nslices, -1);
  M->vals.clear();
  M->idx.clear();
  for (int s = 0; s < M->nslices; s++) {
    int r0 = s * 32, r1 = std::min(nrows, r0 + 32);
    size_t width = 0, nnz = 0;
    for (int r = r0; r < r1; r++) {
      width = std::max(width, rows[r].size());
      nnz += rows[r].size();
    }
    // dense slice moves 8 B per (row, col); sparse moves 12 B per padded entry
    bool dense = ncols > 0 && (size_t)ncols * 8 <= width * 12;
    if (dense) width = ncols;
    size_t base = M->vals.size();
    M->vals.resize(base + width * 32, 0.0);
    if (dense) {
      for (int r = r0; r < r1; r++)
        for (auto &e : rows[r]) M->vals[base + (size_t)e.first * 32 + (r - r0)] = e.second;
    } else {
      M->iptr[s] = (int)(M->idx.size() / 32);
      size_t ibase = M->idx.size();
      M->idx.resize(ibase + width * 32, 0);
      for (int r = r0; r < r1; r++) {
        size_t j = 0;
        for (auto &e : rows[r]) {
          M->vals[base + j * 32 + (r - r0)] = e.second;
          M->idx[ibase + j * 32 + (r - r0)] = e.first;
          j++;
        }
      }
    }
    M->sptr[s + 1] = (int)(M->vals.size() / 32);
    (void)nnz;
  }
}

}  // namespace


// ------------------------------------------------------------------------------------------------------
// Streamed layout (see bqp_internal.h): groups of <= 16 slices, fixed-size stages in (chunk, quad) order.
namespace {

using Entry = std::pair<int, double>;
using Row = std::vector<Entry>;

// rows: 32*nsl entry lists (absolute columns of the input vector, sorted by column)
void add_group(HostStream &st, int kind, int row0, const std::vector<const Row *> &rows) {
  static const Row empty;
  StreamGroup G{};
  G.kind = kind; G.row0 = row0; G.nsl = (int)rows.size() / 32; G.in_off = 0;
  int clo[4], chi[4], mlen[4];
  long long dense_stages = 0, sparse_stages = 0;
  for (int q = 0; q < 4; q++) {
    clo[q] = 1 << 30; chi[q] = -1; mlen[q] = 0;
    for (int r = q * 128; r < std::min<int>((q + 1) * 128, (int)rows.size()); r++) {
      const Row &R = rows[r] ? *rows[r] : empty;
      if (R.empty()) continue;
      clo[q] = std::min(clo[q], R.front().first);
      chi[q] = std::max(chi[q], R.back().first + 1);
      mlen[q] = std::max(mlen[q], (int)R.size());
    }
    if (chi[q] < 0) { clo[q] = 0; chi[q] = 0; }
    clo[q] &= ~3;   // 32-byte aligned first column (the A' input chunk is fetched by TMA from global memory)
    // a quad that owns slices must still see the group (its warps arrive on the barriers) only if it has data
    dense_stages += (chi[q] - clo[q] + kKC - 1) / kKC;
    sparse_stages += (mlen[q] + kKC - 1) / kKC;
  }
  G.sparse = sparse_stages * (kStageValBytes + kStageIdxBytes) < dense_stages * kStageValBytes ? 1 : 0;
  for (int q = 0; q < 4; q++) {
    G.qcol0[q] = G.sparse ? 0 : clo[q];
    G.qch[q] = G.sparse ? (mlen[q] + kKC - 1) / kKC : (chi[q] - clo[q] + kKC - 1) / kKC;
  }
  G.data_off = (long long)st.data.size();
  const int sb = kStageValBytes + (G.sparse ? kStageIdxBytes : 0);
  if (G.sparse) st.slot_bytes = std::max(st.slot_bytes, sb);
  int maxc = std::max(std::max(G.qch[0], G.qch[1]), std::max(G.qch[2], G.qch[3]));
  std::vector<size_t> cursor(rows.size(), 0);   // dense mode: next entry of each row not yet emitted
  (void)maxc;
  for (int q = 0; q < 4; q++)
    for (int c = 0; c < G.qch[q]; c++) {
      const size_t base = st.data.size();
      st.data.resize(base + sb, 0);
      double *vals = reinterpret_cast<double *>(st.data.data() + base);
      int *idx = reinterpret_cast<int *>(st.data.data() + base + kStageValBytes);
      for (int wq = 0; wq < 4; wq++)
        for (int lane = 0; lane < 32; lane++) {
          const size_t r = (size_t)(q * 4 + wq) * 32 + lane;
          if (r >= rows.size() || !rows[r]) continue;
          const Row &R = *rows[r];
          if (G.sparse) {
            for (int j = 0; j < kKC; j++) {
              const size_t e = (size_t)c * kKC + j;
              if (e >= R.size()) break;
              vals[(wq * kKC + j) * 32 + lane] = R[e].second;
              idx[(wq * kKC + j) * 32 + lane] = R[e].first;
            }
          } else {
            const int col0 = G.qcol0[q] + c * kKC;
            size_t &k = cursor[r];
            // register-blocked dense layout: consumer lane (kp = lane'>>3, r8 = lane'&7) owns rows r8+8i (i<4) of
            // the slice and the columns 4j+kp of the chunk; rows (0,1) and (2,3) are two 16-byte vectors laid out
            // [j][pair][lane] so that each LDS.128 of a warp is bank-conflict free
            while (k < R.size() && R[k].first < col0 + kKC) {
              const int dc = R[k].first - col0, j = dc >> 2, kp = dc & 3;
              const int lane2 = (lane & 7) + 8 * kp, i = lane >> 3;
              vals[(size_t)wq * (kKC * 32) + (((size_t)j * 2 + (i >> 1)) * 32 + lane2) * 2 + (i & 1)] = R[k].second;
              k++;
            }
          }
        }
    }
  st.groups.push_back(G);
}

// split `rows` (a multiple of 32 entries is not required) into groups of <= kStreamWarps slices of similar size
void add_panel(HostStream &st, int kind, int row0, const std::vector<const Row *> &rows_in) {
  std::vector<const Row *> rows(rows_in);
  while (rows.size() % 32) rows.push_back(nullptr);
  const int nsl = (int)rows.size() / 32;
  if (nsl == 0) return;
  const int ng = (nsl + kStreamWarps - 1) / kStreamWarps, spg = (nsl + ng - 1) / ng;
  for (int g = 0; g < ng; g++) {
    const int s0 = g * spg, s1 = std::min(nsl, s0 + spg);
    if (s0 >= s1) break;
    std::vector<const Row *> sub(rows.begin() + (size_t)s0 * 32, rows.begin() + (size_t)s1 * 32);
    add_group(st, kind, row0 + s0 * 32, sub);
  }
}

long long bytes_of(const HostStream &st, int g0, int g1) {
  long long b = 0;
  for (int g = g0; g < g1; g++) {
    const StreamGroup &G = st.groups[g];
    b += (long long)(G.qch[0] + G.qch[1] + G.qch[2] + G.qch[3]) * (kStageValBytes + (G.sparse ? kStageIdxBytes : 0));
  }
  return b;
}

// S: dense column-major npad x npad holding unit-lower L22 strictly below the diagonal
void build_stream(HostInstance *h, const std::vector<Row> &arows, const std::vector<Row> &atrows,
                  const std::vector<Row> &prows, const std::vector<double> &S, int tri_nb) {
  HostStream &st = h->st;
  const int np_ = h->npad;
  st = HostStream();
  st.tri_nb = tri_nb;
  auto L = [&](int r, int c) -> double { return S[(size_t)c * np_ + r]; };
  auto ptrs = [](const std::vector<Row> &v) { std::vector<const Row *> p; for (auto &r : v) p.push_back(&r); return p; };
  st.range[GK_AT][0] = (int)st.groups.size();
  add_panel(st, GK_AT, 0, ptrs(atrows));
  st.range[GK_AT][1] = (int)st.groups.size();
  // triangular sweeps over super-blocks of tri_nb columns with explicitly inverted diagonal super-blocks
  const int nsb = (np_ + tri_nb - 1) / tri_nb;
  std::vector<std::vector<double>> Xs(nsb);
  for (int J = 0; J < nsb; J++) {
    const int r0 = J * tri_nb, nbj = std::min(tri_nb, np_ - r0);
    std::vector<double> &X = Xs[J];
    X.assign((size_t)nbj * nbj, 0.0);   // row-major inverse of the unit-lower diagonal super-block
    for (int c = 0; c < nbj; c++) {
      X[(size_t)c * nbj + c] = 1.0;
      for (int r = c + 1; r < nbj; r++) {
        double acc = 0;
        for (int k = c; k < r; k++) acc += L(r0 + r, r0 + k) * X[(size_t)k * nbj + c];
        X[(size_t)r * nbj + c] = -acc;
      }
    }
  }
  st.fw[0] = (int)st.groups.size();
  for (int J = 0; J < nsb; J++) {
    const int r0 = J * tri_nb, nbj = std::min(tri_nb, np_ - r0);
    std::vector<Row> d(nbj), u(std::max(0, np_ - r0 - nbj));
    for (int r = 0; r < nbj; r++)
      for (int c = 0; c <= r; c++) d[r].push_back({r0 + c, Xs[J][(size_t)r * nbj + c]});
    for (int r = r0 + nbj; r < np_; r++)
      for (int c = 0; c < nbj; c++) u[r - r0 - nbj].push_back({r0 + c, L(r, r0 + c)});
    add_panel(st, GK_FWD_D, r0, ptrs(d));
    add_panel(st, GK_FWD_U, r0 + nbj, ptrs(u));
  }
  st.fw[1] = st.bw[0] = (int)st.groups.size();
  for (int J = nsb - 1; J >= 0; J--) {
    const int r0 = J * tri_nb, nbj = std::min(tri_nb, np_ - r0);
    std::vector<Row> d(nbj), u(r0);
    for (int c = 0; c < nbj; c++)
      for (int rr = c; rr < nbj; rr++) d[c].push_back({r0 + rr, Xs[J][(size_t)rr * nbj + c]});
    for (int c = 0; c < r0; c++)
      for (int rr = 0; rr < nbj; rr++) u[c].push_back({r0 + rr, L(r0 + rr, c)});
    add_panel(st, GK_BWD_D, r0, ptrs(d));
    add_panel(st, GK_BWD_U, 0, ptrs(u));
  }
  st.bw[1] = (int)st.groups.size();
  st.range[GK_AB][0] = (int)st.groups.size();
  add_panel(st, GK_AB, 0, ptrs(arows));
  st.range[GK_AB][1] = st.range[GK_PM][0] = (int)st.groups.size();
  add_panel(st, GK_PM, 0, ptrs(prows));
  st.range[GK_PM][1] = (int)st.groups.size();
  st.iter_bytes = bytes_of(st, st.range[GK_AT][0], st.range[GK_AB][1]);
  st.check_bytes = 2 * (bytes_of(st, st.range[GK_AT][0], st.range[GK_AT][1]) + bytes_of(st, st.range[GK_AB][0], st.range[GK_AB][1]) +
                        bytes_of(st, st.range[GK_PM][0], st.range[GK_PM][1]));
  st.built = true;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
// Row-panel layout of the fused single-pass kernel (bqp_internal.h, bqp_panel.cu).
namespace {

// X: dense row-major rows x npad (rows beyond `rows` and columns beyond the data are zero) -> panels appended to pn.data
// position of X[r][c] inside the panel stream of one matrix (bqp_internal.h: fragment-major tiles)
inline size_t panel_pos(int r, int c, int npad) {
  const int k = r / kPanelRows, rp = r % kPanelRows, h = rp >> 3, rr = rp & 7, w = c / 32, cc = c % 32;
  return (size_t)k * kPanelRows * npad + (size_t)w * (kPanelRows * 32) + (size_t)h * 256 + (size_t)(cc >> 2) * 32 + (size_t)rr * 4 + (cc & 3);
}

void append_panels(HostPanels &pn, const std::vector<double> &X, int rows, int npad) {
  const int npanels = (rows + kPanelRows - 1) / kPanelRows;
  const size_t base = pn.data.size();
  pn.data.resize(base + (size_t)npanels * kPanelRows * npad, 0.0);
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < npad; c++) pn.data[base + panel_pos(r, c, npad)] = X[(size_t)r * npad + c];
}

// S: column-major npad x npad, unit-lower L22 strictly below the diagonal (after the LDL' above); D2inv its inverse pivots.
// Returns M = (P + sigma I + A' rho A)^-1 = L22^-T D2^-1 L22^-1, row-major npad x npad (zero beyond n).
// X = inv(L22) (unit lower, row-major n x n)
std::vector<double> unit_lower_inverse(const HostInstance *h, const std::vector<double> &S) {
  const int n = h->n, np_ = h->npad;
  std::vector<double> X((size_t)n * n, 0.0);
  for (int r = 0; r < n; r++) {
    double *xr = &X[(size_t)r * n];
    for (int k = 0; k < r; k++) {
      const double l = S[(size_t)k * np_ + r];
      if (l == 0.0) continue;
      const double *xk = &X[(size_t)k * n];
      for (int c = 0; c <= k; c++) xr[c] -= l * xk[c];
    }
    xr[r] = 1.0;
  }
  return X;
}

// ---- symmetric eigenproblem: Householder tridiagonalisation + implicit QL (the EISPACK tred2 / tql2 pair).
// Z is row-major n x n, symmetric on input; on output ROW j of Z is the eigenvector of eigenvalue d[j] (kept transposed so
// that the plane rotations of the QL sweep run over contiguous memory).
void sym_eig(int n, std::vector<double> &Z, std::vector<double> &d) {
  std::vector<double> e(n, 0.0), V(Z);
  d.assign(n, 0.0);
  auto A = [&](int r, int c) -> double & { return V[(size_t)r * n + c]; };
  for (int j = 0; j < n; j++) d[j] = A(n - 1, j);
  for (int i = n - 1; i > 0; i--) {
    double scale = 0.0, hsum = 0.0;
    for (int k = 0; k < i; k++) scale += std::fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; j++) { d[j] = A(i - 1, j); A(i, j) = 0.0; A(j, i) = 0.0; }
    } else {
      for (int k = 0; k < i; k++) { d[k] /= scale; hsum += d[k] * d[k]; }
      double f = d[i - 1], g = std::sqrt(hsum);
      if (f > 0) g = -g;
      e[i] = scale * g;
      hsum -= f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; j++) e[j] = 0.0;
      for (int j = 0; j < i; j++) {
        f = d[j];
        A(j, i) = f;
        g = e[j] + A(j, j) * f;
        for (int k = j + 1; k <= i - 1; k++) { g += A(k, j) * d[k]; e[k] += A(k, j) * f; }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; j++) { e[j] /= hsum; f += e[j] * d[j]; }
      const double hh = f / (hsum + hsum);
      for (int j = 0; j < i; j++) e[j] -= hh * d[j];
      for (int j = 0; j < i; j++) {
        f = d[j]; g = e[j];
        for (int k = j; k <= i - 1; k++) A(k, j) -= (f * e[k] + g * d[k]);
        d[j] = A(i - 1, j);
        A(i, j) = 0.0;
      }
    }
    d[i] = hsum;
  }
  for (int i = 0; i < n - 1; i++) {                    // accumulate the transformations
    A(n - 1, i) = A(i, i);
    A(i, i) = 1.0;
    const double hsum = d[i + 1];
    if (hsum != 0.0) {
      for (int k = 0; k <= i; k++) d[k] = A(k, i + 1) / hsum;
      for (int j = 0; j <= i; j++) {
        double g = 0.0;
        for (int k = 0; k <= i; k++) g += A(k, i + 1) * A(k, j);
        for (int k = 0; k <= i; k++) A(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; k++) A(k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; j++) { d[j] = A(n - 1, j); A(n - 1, j) = 0.0; }
  A(n - 1, n - 1) = 1.0;
  e[0] = 0.0;
  for (int r = 0; r < n; r++)                          // Z = V' : eigenvector columns become rows
    for (int c = 0; c < n; c++) Z[(size_t)c * n + r] = V[(size_t)r * n + c];
  V.clear(); V.shrink_to_fit();
  for (int i = 1; i < n; i++) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = std::ldexp(1.0, -52);
  for (int l = 0; l < n; l++) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n) { if (std::fabs(e[m]) <= eps * tst1) break; m++; }
    if (m > l) {
      int sweeps = 0;
      do {
        if (++sweeps > 200) break;                     // never seen; the guard of the inverse catches a bad decomposition
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = std::hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double hh = g - d[l];
        for (int i = l + 2; i < n; i++) d[i] -= hh;
        f += hh;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c, s = 0.0, s2 = 0.0;
        const double el1 = e[l + 1];
        for (int i = m - 1; i >= l; i--) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          hh = c * p;
          r = std::hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = hh + s * (c * g + s * d[i]);
          double *zi = &Z[(size_t)i * n], *zi1 = &Z[(size_t)(i + 1) * n];
          for (int k = 0; k < n; k++) {
            const double t = zi1[k];
            zi1[k] = s * zi[k] + c * t;
            zi[k] = c * zi[k] - s * t;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1);
    }
    d[l] += f;
    e[l] = 0.0;
  }
}

std::vector<double> reduced_inverse(const HostInstance *h, const std::vector<double> &S) {
  const int n = h->n, np_ = h->npad;
  // X = inv(L22) (unit lower, row-major, leading n x n block), row by row: X[r][:] = e_r - sum_{k<r} L[r][k] X[k][:]
  std::vector<double> X((size_t)n * n, 0.0);
  for (int r = 0; r < n; r++) {
    double *xr = &X[(size_t)r * n];
    for (int k = 0; k < r; k++) {
      const double l = S[(size_t)k * np_ + r];
      if (l == 0.0) continue;
      const double *xk = &X[(size_t)k * n];
      for (int c = 0; c <= k; c++) xr[c] -= l * xk[c];
    }
    xr[r] = 1.0;
  }
  // M = X' D2^-1 X  (lower triangle accumulated by rank-1 updates with the rows of X, then mirrored)
  std::vector<double> M((size_t)np_ * np_, 0.0);
  for (int k = 0; k < n; k++) {
    const double *xk = &X[(size_t)k * n];
    const double d = h->D2inv[k];
    for (int i = 0; i <= k; i++) {
      const double f = xk[i] * d;
      if (f == 0.0) continue;
      double *mi = &M[(size_t)i * np_];
      for (int j = 0; j <= i; j++) mi[j] += f * xk[j];
    }
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++) M[(size_t)j * np_ + i] = M[(size_t)i * np_ + j];
  return M;
}

// Adaptive rho (osqp adapt_rho): rho_vec = rho x {1, 1e3} on the inequality / equality rows, RHO_MIN on the
// loose ones, so K(rho) = K0 + (rho - rho0) S1 with K0 = K(rho0) the matrix factorised at setup and S1 = A' diag(type weight) A.
// With K0 = L L' and L^-1 S1 L^-T = Q diag(mu) Q':  K(rho)^-1 = V diag(1 / (1 + (rho - rho0) mu)) V',  V = L^-T Q  (mu < 1/rho0).
// The kernels apply x~ = V (d . (V' b)) with the leaf's own rho in d: ANY rho per leaf, no refactorisation, same stream for
// the 8 leaves of a tile.  At rho = rho0 this is V V' = K0^-1, as well conditioned as the factor itself.
// W = V' and Vm = V come back row-major npad x npad (zero beyond n), mu with npad entries.
void spectral_factors(const HostInstance *h, const std::vector<Row> &arows, const std::vector<double> &S, std::vector<double> &W,
                      std::vector<double> &Vm, std::vector<double> &mu_out) {
  const int n = h->n, np_ = h->npad;
  {
    const int m = h->m;
    std::vector<double> X = unit_lower_inverse(h, S);            // L = L22 D2^(1/2):  L^-1 = D2^(-1/2) X
    std::vector<double> S1((size_t)n * n, 0.0);
    for (int r = 0; r < m; r++) {
      const double wgt = h->rtype[r] == 0 ? 1.0 : (h->rtype[r] == 1 ? kRhoEqFactor : 0.0);
      if (wgt == 0.0) continue;
      const auto &row = arows[r];
      for (size_t a = 0; a < row.size(); a++) {
        const double va = wgt * row[a].second;
        double *dst = &S1[(size_t)row[a].first * n];
        for (size_t b = 0; b < row.size(); b++) dst[row[b].first] += va * row[b].second;
      }
    }
    // T = X S1 (X lower triangular), C = T X' scaled by D2^(-1/2) on both sides
    std::vector<double> Tm((size_t)n * n, 0.0), C((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) {
      double *ti = &Tm[(size_t)i * n];
      for (int k = 0; k <= i; k++) {
        const double x = X[(size_t)i * n + k];
        if (x == 0.0) continue;
        const double *sk = &S1[(size_t)k * n];
        for (int j = 0; j < n; j++) ti[j] += x * sk[j];
      }
    }
    S1.clear(); S1.shrink_to_fit();
    std::vector<double> dh(n);
    for (int i = 0; i < n; i++) dh[i] = std::sqrt(h->D2inv[i]);
    for (int i = 0; i < n; i++)
      for (int j = 0; j <= i; j++) {
        const double *ti = &Tm[(size_t)i * n], *xj = &X[(size_t)j * n];
        double acc = 0.0;
        for (int k = 0; k <= j; k++) acc += ti[k] * xj[k];
        C[(size_t)i * n + j] = C[(size_t)j * n + i] = acc * dh[i] * dh[j];
      }
    Tm.clear(); Tm.shrink_to_fit();
    std::vector<double> mu;
    sym_eig(n, C, mu);                                             // row j of C: eigenvector q_j
    // W = V' = Q' D2^(-1/2) X  (row j of W = (L^-T q_j)')
    W.assign((size_t)np_ * np_, 0.0); Vm.assign((size_t)np_ * np_, 0.0);
    for (int j = 0; j < n; j++) {
      double *wj = &W[(size_t)j * np_];
      const double *qj = &C[(size_t)j * n];
      for (int k = 0; k < n; k++) {
        const double f = qj[k] * dh[k];
        if (f == 0.0) continue;
        const double *xk = &X[(size_t)k * n];
        for (int i = 0; i <= k; i++) wj[i] += f * xk[i];
      }
    }
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) Vm[(size_t)i * np_ + j] = W[(size_t)j * np_ + i];
    mu_out.assign(np_, 0.0);
    for (int j = 0; j < n; j++) mu_out[j] = mu[j];
  }
}

void build_panels(HostInstance *h, const std::vector<Row> &arows, const std::vector<Row> &prows, const std::vector<double> &S) {
  const int n = h->n, m = h->m, np_ = h->npad;
  HostPanels &pn = h->pn;
  pn = HostPanels();
  std::vector<double> M, Vm;
  if (h->s.adaptive_rho) { spectral_factors(h, arows, S, M, Vm, pn.mu); pn.spectral = true; }      // the M slot holds V'
  else M = reduced_inverse(h, S);
  std::vector<double> Ad((size_t)std::max(m, 1) * np_, 0.0), Pd((size_t)np_ * np_, 0.0);
  for (int r = 0; r < m; r++)
    for (auto &e : arows[r]) Ad[(size_t)r * np_ + e.first] = e.second;
  for (int r = 0; r < n; r++)
    for (auto &e : prows[r]) Pd[(size_t)r * np_ + e.first] = e.second;
  pn.nw = np_ / 32;
  pn.panel_doubles = (long long)kPanelRows * np_;
  pn.npm = np_ / kPanelRows;
  pn.npa = (m + kPanelRows - 1) / kPanelRows;
  append_panels(pn, M, np_, np_);
  pn.offA = (long long)pn.data.size();
  append_panels(pn, Ad, m, np_);
  pn.offP = (long long)pn.data.size();
  append_panels(pn, Pd, np_, np_);
  pn.offV = (long long)pn.data.size();
  if (pn.spectral) append_panels(pn, Vm, np_, np_);
  pn.built = true;
}

// whole-GPU layout: M and P as panels (no dense A), A and A' as CSR
void build_grid(HostInstance *h, const std::vector<Row> &arows, const std::vector<Row> &atrows, const std::vector<Row> &prows,
                const std::vector<double> &S) {
  const int n = h->n, np_ = h->npad;
  HostGridL &gd = h->gd;
  gd = HostGridL();
  HostPanels tmp;
  if (h->s.adaptive_rho) {
    std::vector<double> W, Vm;
    spectral_factors(h, arows, S, W, Vm, gd.mu);
    append_panels(tmp, W, np_, np_);
    gd.offV = (long long)tmp.data.size();
    append_panels(tmp, Vm, np_, np_);
    gd.spectral = true;
  } else {
    const std::vector<double> M = reduced_inverse(h, S);
    append_panels(tmp, M, np_, np_);
  }
  gd.offP = (long long)tmp.data.size();
  {
    std::vector<double> Pd((size_t)np_ * np_, 0.0);
    for (int r = 0; r < n; r++)
      for (auto &e : prows[r]) Pd[(size_t)r * np_ + e.first] = e.second;
    append_panels(tmp, Pd, np_, np_);
  }
  gd.MP.swap(tmp.data);
  gd.npm = np_ / kPanelRows;
  auto csr = [](const std::vector<Row> &rows, std::vector<int> &rp, std::vector<int> &ci, std::vector<double> &vl) {
    rp.assign(rows.size() + 1, 0);
    for (size_t r = 0; r < rows.size(); r++) {
      for (auto &e : rows[r]) { ci.push_back(e.first); vl.push_back(e.second); }
      rp[r + 1] = (int)ci.size();
    }
  };
  csr(arows, gd.arp, gd.aci, gd.avl);
  csr(atrows, gd.trp, gd.tci, gd.tvl);
  gd.built = true;
}

// shared-memory-resident layout (bqp_small.cu): M and P as mma A-fragments, A and A' as entry-major ELL, rho and 1 / rho
// per row, in one blob (one TMA bulk copy).  Returns false when the problem does not qualify (the blob plus the kernel's
// vectors must fit the shared memory of one CTA; columns are stored as u16)
bool build_small(HostInstance *h, const std::vector<Row> &arows, const std::vector<Row> &atrows, const std::vector<Row> &prows,
                 const std::vector<double> &S) {
  const int n = h->n, m = h->m, np_ = h->npad;
  HostSmall &sm = h->sm;
  sm = HostSmall();
  if (np_ > 64 || m > 60000) return false;
  int wa = 1, wt = 1;
  for (int r = 0; r < m; r++) wa = std::max(wa, (int)arows[r].size());
  for (int j = 0; j < n; j++) wt = std::max(wt, (int)atrows[j].size());
  // the kernel keeps a thread's ELL entries in registers: 3 rows of A with up to 4 entries, one row of A' with up to 8
  // (config 3: 3 and 5).  Wider or taller problems stay on the direct-load kernel
  if (wa > 4 || wt > 8 || m > 192) return false;
  const int mp = std::max(8, (m + 7) / 8 * 8);
  auto align16 = [](size_t v) { return (v + 15) & ~size_t(15); };
  size_t off = 0;
  const size_t offM = off; off += (size_t)np_ * np_ * 8;
  const size_t offP = off; off += (size_t)np_ * np_ * 8;
  const size_t offAv = off; off += (size_t)wa * mp * 8;
  const size_t offTv = off; off += (size_t)wt * np_ * 8;
  const size_t offRho = off; off += (size_t)mp * 8;
  const size_t offRinv = off; off += (size_t)mp * 8;
  const size_t offE = off; off += (size_t)mp * 8;
  const size_t offEinv = off; off += (size_t)mp * 8;
  const size_t offAc = off; off = align16(off + (size_t)wa * mp * 2);
  const size_t offTc = off; off = align16(off + (size_t)wt * np_ * 2);
  if (small_smem_bytes(np_, m, (int)off) > (size_t)kMaxSmem) return false;
  sm.blob.assign(off, 0);
  const std::vector<double> M = reduced_inverse(h, S);
  std::vector<double> Pd((size_t)np_ * np_, 0.0);
  for (int r = 0; r < n; r++)
    for (auto &e : prows[r]) Pd[(size_t)r * np_ + e.first] = e.second;
  auto frag = [&](const std::vector<double> &X, size_t o) {
    double *dst = reinterpret_cast<double *>(sm.blob.data() + o);
    const int ks = np_ / 4;
    for (int p = 0; p < np_ / 8; p++)
      for (int s = 0; s < ks; s++)
        for (int lane = 0; lane < 32; lane++) dst[((size_t)p * ks + s) * 32 + lane] = X[(size_t)(8 * p + (lane >> 2)) * np_ + 4 * s + (lane & 3)];
  };
  frag(M, offM); frag(Pd, offP);
  double *av = reinterpret_cast<double *>(sm.blob.data() + offAv), *tv = reinterpret_cast<double *>(sm.blob.data() + offTv);
  double *rho = reinterpret_cast<double *>(sm.blob.data() + offRho), *rinv = reinterpret_cast<double *>(sm.blob.data() + offRinv);
  uint16_t *ac = reinterpret_cast<uint16_t *>(sm.blob.data() + offAc), *tc = reinterpret_cast<uint16_t *>(sm.blob.data() + offTc);
  for (int r = 0; r < m; r++)
    for (size_t k = 0; k < arows[r].size(); k++) { av[k * mp + r] = arows[r][k].second; ac[k * mp + r] = (uint16_t)arows[r][k].first; }
  for (int j = 0; j < n; j++)
    for (size_t k = 0; k < atrows[j].size(); k++) { tv[k * np_ + j] = atrows[j][k].second; tc[k * np_ + j] = (uint16_t)atrows[j][k].first; }
  double *es = reinterpret_cast<double *>(sm.blob.data() + offE), *eis = reinterpret_cast<double *>(sm.blob.data() + offEinv);
  for (int r = 0; r < mp; r++) {
    rho[r] = r < m ? h->rho[r] : 1.0; rinv[r] = r < m ? h->rho_inv[r] : 1.0;
    es[r] = r < m ? h->E[r] : 1.0; eis[r] = r < m ? h->Einv[r] : 1.0;
  }
  sm.wa = wa; sm.wt = wt; sm.mp = mp;
  sm.offP = (int)offP; sm.offAv = (int)offAv; sm.offTv = (int)offTv; sm.offRho = (int)offRho; sm.offRinv = (int)offRinv; sm.offE = (int)offE; sm.offEinv = (int)offEinv;
  sm.offAc = (int)offAc; sm.offTc = (int)offTc; sm.bytes = (int)off;
  sm.built = true;
  return true;
}

}  // namespace

void host_rescale_q(HostInstance *h, const double *q) {   // osqp update_lin_cost: q_scaled = c * D * q
  h->nq = 0;
  for (int j = 0; j < h->n; j++) {
    h->q[j] = h->c * h->D[j] * q[j];
    h->nq = std::max(h->nq, std::fabs(h->Dinv[j] * h->q[j]));
  }
}

int host_setup(const bqp_problem *p, const bqp_settings *s, HostInstance *h) {
  if (!p || !s || !h) return BQP_E_ARG;
  const int n = p->n, m = p->m;
  if (n <= 0 || m < 0 || !p->Pp || !p->Ap || !p->q || (m > 0 && (!p->l || !p->u))) return BQP_E_ARG;
  if (p->n_int < 0 || p->n_int > m || (p->n_int > 0 && !p->i_idx)) return BQP_E_ARG;
  if (s->check_termination < 1 || s->max_iter < 1 || s->scaling < 0 || !(s->rho > 0) || !(s->sigma > 0) ||
      !(s->alpha > 0 && s->alpha < 2))
    return BQP_E_ARG;
  for (int r = 0; r < m; r++)
    if (p->l[r] > p->u[r]) return BQP_E_BOUNDS;
  if (s->adaptive_rho) {
    // a fixed interval that coincides with termination checks (osqp rounds its automatic interval to one, too)
    if (s->adaptive_rho_interval <= 0 || s->adaptive_rho_interval % s->check_termination != 0 || !(s->adaptive_rho_tolerance > 1.0)) return BQP_E_ARG;
    if (s->eq_rho == 2) return BQP_E_UNSUPPORTED;
  }
  h->n = n; h->m = m; h->npad = ((n + kNB - 1) / kNB) * kNB; h->n_int = p->n_int; h->s = *s;
  h->i_idx.assign(p->i_idx, p->i_idx + p->n_int);
  for (int k = 0; k < p->n_int; k++)
    if (h->i_idx[k] < 0 || h->i_idx[k] >= n) return BQP_E_ARG;

  Csc P, A;
  P.rows = P.cols = n;
  P.p.assign(p->Pp, p->Pp + n + 1);
  for (int j = 0; j < n; j++)            // keep the upper triangle only
    for (int k = p->Pp[j]; k < p->Pp[j + 1]; k++)
      if (p->Pi[k] < 0 || p->Pi[k] >= n) return BQP_E_ARG;
  {
    std::vector<int> np(n + 1, 0);
    for (int j = 0; j < n; j++) {
      np[j] = (int)P.i.size();
      for (int k = p->Pp[j]; k < p->Pp[j + 1]; k++)
        if (p->Pi[k] <= j) { P.i.push_back(p->Pi[k]); P.x.push_back(p->Px[k]); }
    }
    np[n] = (int)P.i.size();
    P.p = np;
  }
  A.rows = m; A.cols = n;
  A.p.assign(p->Ap, p->Ap + n + 1);
  A.i.assign(p->Ai, p->Ai + p->Ap[n]);
  A.x.assign(p->Ax, p->Ax + p->Ap[n]);
  for (int v : A.i) if (v < 0 || v >= m) return BQP_E_ARG;
  h->q.assign(p->q, p->q + n);

  ruiz(P, A, h->q, s->scaling, h);

  // rho typing on the SCALED, clamped root bounds
  h->rho.resize(m); h->rho_inv.resize(m); h->rtype.assign(m, 0);
  for (int r = 0; r < m; r++) {
    double lo = std::max(p->l[r], -kInfty) * h->E[r], up = std::min(p->u[r], kInfty) * h->E[r];
    double rho = s->rho;
    if (s->eq_rho) {
      if (lo < -kInfty * kMinScaling && up > kInfty * kMinScaling) { rho = kRhoMin; h->rtype[r] = -1; }
      else if (up - lo < kRhoTol) { rho = kRhoEqFactor * s->rho; h->rtype[r] = 1; }
    }
    h->rho[r] = rho; h->rho_inv[r] = 1.0 / rho;
  }

  // row lists of A (m x n), A' (n x m) and full symmetric P (n x n)
  RowList arows(m), atrows(n), prows(n);
  for (int j = 0; j < n; j++)
    for (int k = A.p[j]; k < A.p[j + 1]; k++) {
      arows[A.i[k]].push_back({j, A.x[k]});
      atrows[j].push_back({A.i[k], A.x[k]});
    }
  for (int j = 0; j < n; j++)
    for (int k = P.p[j]; k < P.p[j + 1]; k++) {
      int r = P.i[k];
      prows[r].push_back({j, P.x[k]});
      if (r != j) prows[j].push_back({r, P.x[k]});
    }
  auto by_col = [](const std::pair<int, double> &a, const std::pair<int, double> &b) { return a.first < b.first; };
  for (auto &r : arows) std::sort(r.begin(), r.end(), by_col);
  for (auto &r : atrows) std::sort(r.begin(), r.end(), by_col);
  for (auto &r : prows) std::sort(r.begin(), r.end(), by_col);
  build_panel(arows, n, &h->Ab);
  build_panel(atrows, m, &h->At);
  build_panel(prows, n, &h->Pm);

  // S = P + sigma I + A' diag(rho) A : dense lower triangle, column-major, padded with identity
  const int np_ = h->npad;
  std::vector<double> S((size_t)np_ * np_, 0.0);
  auto Sat = [&](int r, int c) -> double & { return S[(size_t)c * np_ + r]; };
  for (int j = 0; j < n; j++)
    for (int k = P.p[j]; k < P.p[j + 1]; k++) Sat(j, P.i[k]) += P.x[k];   // (row j >= col i) lower copy of triu
  for (int j = 0; j < n; j++) Sat(j, j) += s->sigma;
  for (int j = n; j < np_; j++) Sat(j, j) = 1.0;
  for (int r = 0; r < m; r++) {
    const auto &row = arows[r];
    const double rho = h->rho[r];
    for (size_t a = 0; a < row.size(); a++) {
      const double va = rho * row[a].second;
      double *col = &S[(size_t)row[a].first * np_];
      for (size_t b = a; b < row.size(); b++) col[row[b].first] += va * row[b].second;   // rows >= col
    }
  }
  // in-place dense LDL^T (right-looking); afterwards S holds unit-lower L22 below the diagonal, D2 on it
  h->D2inv.assign(np_, 1.0);
  for (int j = 0; j < np_; j++) {
    double d = Sat(j, j);
    if (!(d > 0.0)) return BQP_E_NONCONVEX;
    double dinv = 1.0 / d;
    h->D2inv[j] = dinv;
    double *cj = &S[(size_t)j * np_];
    for (int k = j + 1; k < np_; k++) {
      const double g = cj[k] * dinv;   // L[k][j]
      if (g == 0.0) continue;
      double *ck = &S[(size_t)k * np_];
      for (int r = k; r < np_; r++) ck[r] -= cj[r] * g;   // S[r][k] -= (L[r][j] d) L[k][j]
    }
    for (int r = j + 1; r < np_; r++) cj[r] *= dinv;
  }
  // inverse of every 32x32 unit-lower diagonal block (strictly-lower part kept; unit diagonal implicit)
  const int nb = np_ / kNB;
  std::vector<double> Linv((size_t)nb * kNB * kNB, 0.0);   // [J][r][c]
  for (int J = 0; J < nb; J++) {
    const int r0 = J * kNB;
    double *X = &Linv[(size_t)J * kNB * kNB];
    for (int c = 0; c < kNB; c++)
      for (int r = c + 1; r < kNB; r++) {
        double acc = Sat(r0 + r, r0 + c);                  // k = c term: L[r][c] * X[c][c]
        for (int k = c + 1; k < r; k++) acc += Sat(r0 + r, r0 + k) * X[k * kNB + c];
        X[r * kNB + c] = -acc;
      }
  }
  // blocked layouts
  h->Lcol.clear(); h->Lrow.clear();
  for (int J = 0; J < nb; J++) {          // block column J: rows J*32..npad, 32 columns, column-major
    int r0 = J * kNB, ld = np_ - r0;
    size_t base = h->Lcol.size();
    h->Lcol.resize(base + (size_t)kNB * ld, 0.0);
    const double *X = &Linv[(size_t)J * kNB * kNB];
    for (int c = 0; c < kNB; c++) {
      for (int r = c + 1; r < kNB; r++) h->Lcol[base + (size_t)c * ld + r] = X[r * kNB + c];
      for (int r = r0 + kNB; r < np_; r++) h->Lcol[base + (size_t)c * ld + (r - r0)] = Sat(r, r0 + c);
    }
  }
  for (int J = 0; J < nb; J++) {          // block row J: 32 rows, columns 0..(J+1)*32, row-major
    int r0 = J * kNB, ld = (J + 1) * kNB;
    size_t base = h->Lrow.size();
    h->Lrow.resize(base + (size_t)kNB * ld, 0.0);
    const double *X = &Linv[(size_t)J * kNB * kNB];
    for (int rr = 0; rr < kNB; rr++) {
      for (int cidx = 0; cidx < r0; cidx++) h->Lrow[base + (size_t)rr * ld + cidx] = Sat(r0 + rr, cidx);
      for (int c = 0; c < rr; c++) h->Lrow[base + (size_t)rr * ld + r0 + c] = X[rr * kNB + c];
    }
  }
  h->nq = 0;
  for (int j = 0; j < n; j++) h->nq = std::max(h->nq, std::fabs(h->Dinv[j] * h->q[j]));
  // streamed layout for the TMA kernel (problems with at least 4 slices of variables)
  h->st = HostStream();
  if (h->At.nslices >= 4) {
    int tri_nb = 256;   // measured best on B200 for n = 500 (fewer, fatter diagonal groups)
    if (const char *e = std::getenv("BQP_TRI_NB")) tri_nb = std::atoi(e);
    tri_nb = std::max(32, std::min(32 * kStreamWarps, (tri_nb / 32) * 32));
    build_stream(h, arows, atrows, prows, S, tri_nb);
  }
  // row-panel layout for the fused single-pass kernel: A stored dense, so only when A is dense enough that one dense
  // pass (8 m n bytes) beats two sparse ones (24 nnz bytes), and the variables fit the 16 column tiles of one CTA
  h->pn = HostPanels();
  {
    double min_density = 0.34;
    if (const char *e = std::getenv("BQP_PANEL_MIN_DENSITY")) min_density = std::atof(e);
    const double density = (m > 0) ? (double)A.x.size() / ((double)m * n) : 1.0;
    if (np_ >= 64 && np_ <= 32 * kPanelMaxWarps && density >= min_density) build_panels(h, arows, prows, S);
  }
  // The panel kernels apply M = (P + sigma I + A' rho A)^-1 explicitly; forming an inverse is not backward stable, and
  // with a tiny sigma, a rank-deficient P or rows typed rho x 1e3 the reduced matrix can be badly conditioned.  Probe it:
  // the same KKT right-hand sides through the panels and through the LDL' substitution (the oracle's path) must agree
  h->pn_inverse_error = NAN; h->pn_rejected = false;
  if (h->pn.built) {
    double tol = 1e-10;
    if (const char *e = std::getenv("BQP_INVERSE_TOL")) tol = std::atof(e);
    double worst = 0.0;
    std::vector<double> r1((size_t)n + m), r2((size_t)n + m);
    for (int probe = 0; probe < 4; probe++) {
      uint64_t st = 0x9E3779B97F4A7C15ull * (uint64_t)(probe + 1);
      for (int k = 0; k < n + m; k++) {
        double v;
        if (probe == 0) v = 1.0;
        else if (probe == 1) v = (k & 1) ? -1.0 : 1.0;
        else { st ^= st << 13; st ^= st >> 7; st ^= st << 17; v = (double)(st >> 11) / 9007199254740992.0 - 0.5; }
        r1[(size_t)k] = r2[(size_t)k] = v;
      }
      host_kkt_solve(h, r1.data());
      host_panel_kkt_solve(h, r2.data());
      double nrm = 0.0, dif = 0.0;
      for (int j = 0; j < n; j++) { nrm = std::max(nrm, std::fabs(r1[(size_t)j])); dif = std::max(dif, std::fabs(r1[(size_t)j] - r2[(size_t)j])); }
      const double rel = dif / std::max(nrm, 1e-300);
      worst = (rel == rel) ? std::max(worst, rel) : INFINITY;
    }
    h->pn_inverse_error = worst;
    if (!(worst <= tol)) { h->pn = HostPanels(); h->pn_rejected = true; }
  }
  // whole-GPU layout: dense-reduced problems too wide for the rows kernel (config 4).  Same guard of the explicit inverse
  h->gd = HostGridL();
  {
    int want = 1, max_np = 2880;                       // x~ of 8 leaves staged in shared memory: 64 npad bytes of the 227 KB
    if (const char *e = std::getenv("BQP_GRID")) want = std::atoi(e);
    // BQP_GRID_ALL=1 (experiments: one config-2 tile on the whole GPU): also for problems the rows kernel serves; run with BQP_KERNEL=grid
    const bool all_sizes = std::getenv("BQP_GRID_ALL") && std::atoi(std::getenv("BQP_GRID_ALL")) != 0;
    // adaptive rho: dense problems the rows kernel serves carry the spectral factors in their panel stream; everything else
    // (wide, sparse or small) runs adaptively on the whole-GPU kernel
    const bool wide = all_sizes ? true : (s->adaptive_rho ? !h->pn.built : (!h->pn.built && np_ > 32 * kPanelMaxWarps));
    if (want && wide && !h->pn_rejected && np_ <= max_np && s->eq_rho != 2) {
      build_grid(h, arows, atrows, prows, S);
      double tol = 1e-10;
      if (const char *e = std::getenv("BQP_INVERSE_TOL")) tol = std::atof(e);
      double worst = 0.0;
      std::vector<double> r1((size_t)n + m), r2((size_t)n + m);
      for (int probe = 0; probe < 3; probe++) {
        uint64_t st = 0x9E3779B97F4A7C15ull * (uint64_t)(probe + 1);
        for (int k = 0; k < n + m; k++) {
          double v;
          if (probe == 0) v = 1.0;
          else if (probe == 1) v = (k & 1) ? -1.0 : 1.0;
          else { st ^= st << 13; st ^= st >> 7; st ^= st << 17; v = (double)(st >> 11) / 9007199254740992.0 - 0.5; }
          r1[(size_t)k] = r2[(size_t)k] = v;
        }
        host_kkt_solve(h, r1.data());
        host_grid_kkt_solve(h, r2.data());
        double nrm = 0.0, dif = 0.0;
        for (int j = 0; j < n; j++) { nrm = std::max(nrm, std::fabs(r1[(size_t)j])); dif = std::max(dif, std::fabs(r1[(size_t)j] - r2[(size_t)j])); }
        const double rel = dif / std::max(nrm, 1e-300);
        worst = (rel == rel) ? std::max(worst, rel) : INFINITY;
      }
      if (!h->pn.built) h->pn_inverse_error = worst;
      if (!(worst <= tol)) { h->gd = HostGridL(); if (!h->pn.built) h->pn_rejected = true; }
    }
  }
  // shared-memory-resident layout for small sparse problems the dense kernels do not serve (config 3); fixed rho typed at
  // setup only.  Same guard of the explicit inverse; BQP_SMALL=0 keeps them on the direct-load LDL' kernel
  h->sm = HostSmall();
  {
    int want = 1;
    if (const char *e = std::getenv("BQP_SMALL")) want = std::atoi(e);
    if (want && np_ <= 64 && !h->pn.built && !h->pn_rejected && !s->adaptive_rho && s->eq_rho != 2 &&
        build_small(h, arows, atrows, prows, S)) {
      double tol = 1e-10;
      if (const char *e = std::getenv("BQP_INVERSE_TOL")) tol = std::atof(e);
      double worst = 0.0;
      std::vector<double> r1((size_t)n + m), r2((size_t)n + m);
      for (int probe = 0; probe < 4; probe++) {
        uint64_t st = 0x9E3779B97F4A7C15ull * (uint64_t)(probe + 1);
        for (int k = 0; k < n + m; k++) {
          double v;
          if (probe == 0) v = 1.0;
          else if (probe == 1) v = (k & 1) ? -1.0 : 1.0;
          else { st ^= st << 13; st ^= st >> 7; st ^= st << 17; v = (double)(st >> 11) / 9007199254740992.0 - 0.5; }
          r1[(size_t)k] = r2[(size_t)k] = v;
        }
        host_kkt_solve(h, r1.data());
        host_small_kkt_solve(h, r2.data());
        double nrm = 0.0, dif = 0.0;
        for (int j = 0; j < n; j++) { nrm = std::max(nrm, std::fabs(r1[(size_t)j])); dif = std::max(dif, std::fabs(r1[(size_t)j] - r2[(size_t)j])); }
        const double rel = dif / std::max(nrm, 1e-300);
        worst = (rel == rel) ? std::max(worst, rel) : INFINITY;
      }
      if (!h->pn.built && !h->gd.built) h->pn_inverse_error = worst;
      if (!(worst <= tol)) { h->sm = HostSmall(); if (!h->pn.built && !h->gd.built) h->pn_rejected = true; }
    }
  }
  if (s->adaptive_rho && !((h->gd.built && h->gd.spectral) || (h->pn.built && h->pn.spectral))) return BQP_E_UNSUPPORTED;
  h->mint.clear();
  if (s->eq_rho == 2) {
    // per-node re-typing corrects the explicit inverse by a Woodbury term over the re-typed integer rows: dense kernels only
    if (!h->pn.built) return BQP_E_UNSUPPORTED;
    h->mint.assign((size_t)std::max(h->n_int, 1) * np_, 0.0);
    for (int k = 0; k < h->n_int; k++)
      for (int c = 0; c < np_; c++) h->mint[(size_t)k * np_ + c] = h->pn.data[panel_pos(h->i_idx[k], c, np_)];
  }
  return BQP_OK;
}

double host_panel_M(const HostInstance *h, int r, int c) { return h->pn.data[panel_pos(r, c, h->npad)]; }

// ---- host-only debug restatements of what the kernel does with the streamed layouts (tests only)
void host_matvec(const HostMat &M, const double *in, double *out) {
  for (int s = 0; s < M.nslices; s++) {
    const int w = M.sptr[s + 1] - M.sptr[s];
    for (int lane = 0; lane < 32; lane++) {
      const int row = s * 32 + lane;
      if (row >= M.rows) break;
      double acc = 0;
      for (int j = 0; j < w; j++) {
        const double a = M.vals[((size_t)M.sptr[s] + j) * 32 + lane];
        const int col = M.iptr[s] < 0 ? j : M.idx[((size_t)M.iptr[s] + j) * 32 + lane];
        acc = std::fma(a, in[col], acc);
      }
      out[row] = acc;
    }
  }
}

void host_kkt_solve(const HostInstance *h, double *rhs) {
  const int n = h->n, m = h->m, np_ = h->npad, nb = np_ / kNB;
  std::vector<double> w(m), b(np_, 0.0), t(std::max(n, m));
  // t2 = rhs_x + A' (rho o rhs_z)
  for (int i = 0; i < m; i++) w[i] = h->rho[i] * rhs[n + i];
  host_matvec(h->At, w.data(), t.data());
  for (int j = 0; j < n; j++) b[j] = rhs[j] + t[j];
  // forward sweep over block columns
  size_t base = 0;
  for (int J = 0; J < nb; J++) {
    const int r0 = J * kNB, ld = np_ - r0;
    const double *Lc = &h->Lcol[base];
    double y[kNB];
    for (int r = 0; r < kNB; r++) {
      double acc = b[r0 + r];
      for (int c = 0; c < r; c++) acc = std::fma(Lc[(size_t)c * ld + r], b[r0 + c], acc);
      y[r] = acc;
    }
    for (int r = 0; r < kNB; r++) b[r0 + r] = y[r];
    for (int r = r0 + kNB; r < np_; r++) {
      double acc = 0;
      for (int c = 0; c < kNB; c++) acc = std::fma(Lc[(size_t)c * ld + (r - r0)], y[c], acc);
      b[r] -= acc;
    }
    base += (size_t)kNB * ld;
  }
  for (int j = 0; j < np_; j++) b[j] *= h->D2inv[j];
  // backward sweep over block rows
  std::vector<size_t> rbase(nb);
  base = 0;
  for (int J = 0; J < nb; J++) { rbase[J] = base; base += (size_t)kNB * (J + 1) * kNB; }
  for (int J = nb - 1; J >= 0; J--) {
    const int r0 = J * kNB, ld = (J + 1) * kNB;
    const double *Lr = &h->Lrow[rbase[J]];
    double x[kNB];
    for (int c = 0; c < kNB; c++) {
      double acc = b[r0 + c];
      for (int rr = c + 1; rr < kNB; rr++) acc = std::fma(Lr[(size_t)rr * ld + r0 + c], b[r0 + rr], acc);
      x[c] = acc;
    }
    for (int c = 0; c < kNB; c++) b[r0 + c] = x[c];
    for (int cidx = 0; cidx < r0; cidx++) {
      double acc = 0;
      for (int rr = 0; rr < kNB; rr++) acc = std::fma(Lr[(size_t)rr * ld + cidx], x[rr], acc);
      b[cidx] -= acc;
    }
  }
  // nu = rho o (A xt - rhs_z)
  host_matvec(h->Ab, b.data(), t.data());
  for (int j = 0; j < n; j++) rhs[j] = b[j];
  for (int i = 0; i < m; i++) rhs[n + i] = h->rho[i] * (t[i] - rhs[n + i]);
}

// ---- host-only debug restatement of the streamed layout (tests only): what consume_group computes
static void host_group_apply(const HostStream &st, const StreamGroup &G, const double *in, double *acc) {
  const int nrows = G.nsl * 32;
  for (int r = 0; r < nrows; r++) acc[r] = 0.0;
  const unsigned char *p = st.data.data() + G.data_off;
  const int sb = kStageValBytes + (G.sparse ? kStageIdxBytes : 0);
  const int maxc = std::max(std::max(G.qch[0], G.qch[1]), std::max(G.qch[2], G.qch[3]));
  (void)maxc;
  for (int q = 0; q < 4; q++)
    for (int c = 0; c < G.qch[q]; c++) {
      const double *vals = reinterpret_cast<const double *>(p);
      const int *idx = reinterpret_cast<const int *>(p + kStageValBytes);
      for (int wq = 0; wq < 4; wq++) {
        if (G.sparse) {
          for (int j = 0; j < kKC; j++)
            for (int lane = 0; lane < 32; lane++) {
              const int r = (q * 4 + wq) * 32 + lane;
              if (r >= nrows) continue;
              acc[r] = std::fma(vals[(wq * kKC + j) * 32 + lane], in[G.in_off + idx[(wq * kKC + j) * 32 + lane]], acc[r]);
            }
        } else {
          for (int j = 0; j < kKC / 4; j++)
            for (int lane2 = 0; lane2 < 32; lane2++)
              for (int i = 0; i < 4; i++) {
                const int r = (q * 4 + wq) * 32 + (lane2 & 7) + 8 * i;
                if (r >= nrows) continue;
                const int col = G.qcol0[q] + c * kKC + 4 * j + (lane2 >> 3);
                acc[r] = std::fma(vals[(size_t)wq * (kKC * 32) + (((size_t)j * 2 + (i >> 1)) * 32 + lane2) * 2 + (i & 1)], in[G.in_off + col], acc[r]);
              }
        }
      }
      p += sb;
    }
}

int host_stream_kkt_solve(const HostInstance *h, double *rhs) {
  const HostStream &st = h->st;
  if (!st.built) return BQP_E_UNSUPPORTED;
  const int n = h->n, m = h->m, np_ = h->npad;
  std::vector<double> w(m + 64, 0.0), b(np_ + 64, 0.0), acc(kStreamWarps * 32), t(m + 64, 0.0);
  for (int i = 0; i < m; i++) w[i] = h->rho[i] * rhs[n + i];
  for (int g = st.range[GK_AT][0]; g < st.range[GK_AT][1]; g++) {
    const StreamGroup &G = st.groups[g];
    host_group_apply(st, G, w.data(), acc.data());
    for (int r = 0; r < G.nsl * 32; r++) if (G.row0 + r < n) b[G.row0 + r] = rhs[G.row0 + r] + acc[r];
  }
  for (int pass = 0; pass < 2; pass++) {
    const int g0 = pass ? st.bw[0] : st.fw[0], g1 = pass ? st.bw[1] : st.fw[1];
    if (pass) for (int j = 0; j < np_; j++) b[j] *= h->D2inv[j];
    for (int g = g0; g < g1; g++) {
      const StreamGroup &G = st.groups[g];
      host_group_apply(st, G, b.data(), acc.data());
      const bool diag = (G.kind == GK_FWD_D || G.kind == GK_BWD_D);
      for (int r = 0; r < G.nsl * 32; r++) {
        if (G.row0 + r >= np_) continue;
        if (diag) b[G.row0 + r] = acc[r]; else b[G.row0 + r] -= acc[r];
      }
    }
  }
  for (int g = st.range[GK_AB][0]; g < st.range[GK_AB][1]; g++) {
    const StreamGroup &G = st.groups[g];
    host_group_apply(st, G, b.data(), acc.data());
    for (int r = 0; r < G.nsl * 32; r++) if (G.row0 + r < m) t[G.row0 + r] = acc[r];
  }
  for (int j = 0; j < n; j++) rhs[j] = b[j];
  for (int i = 0; i < m; i++) rhs[n + i] = h->rho[i] * (t[i] - rhs[n + i]);
  return BQP_OK;
}

// what the panel kernel computes for one KKT solve, from the panel data: b = rhs_x + A'(rho rhs_z); x~ = M b;
// nu = rho (A x~ - rhs_z).  Lane-level summation order is not reproduced (plain loops): a layout check.
int host_panel_kkt_solve(const HostInstance *h, double *rhs) {
  const HostPanels &pn = h->pn;
  if (!pn.built) return BQP_E_UNSUPPORTED;
  const int n = h->n, m = h->m, np_ = h->npad;
  auto at = [&](long long off, int r, int c) -> double { return pn.data[(size_t)off + panel_pos(r, c, np_)]; };
  std::vector<double> b(np_, 0.0), xt(np_, 0.0);
  for (int j = 0; j < n; j++) b[j] = rhs[j];
  for (int i = 0; i < m; i++) {
    const double w = h->rho[i] * rhs[n + i];
    for (int j = 0; j < np_; j++) b[j] = std::fma(at(pn.offA, i, j), w, b[j]);
  }
  for (int r = 0; r < np_; r++) {
    double acc = 0;
    for (int j = 0; j < np_; j++) acc = std::fma(at(0, r, j), b[j], acc);
    xt[r] = acc;
  }
  if (pn.spectral) {       // the M slot held V': x~ = V (V' b) at the setup rho
    std::vector<double> c(xt);
    for (int r = 0; r < np_; r++) {
      double acc = 0;
      for (int j = 0; j < np_; j++) acc = std::fma(at(pn.offV, r, j), c[j], acc);
      xt[r] = acc;
    }
  }
  for (int i = 0; i < m; i++) {
    double acc = 0;
    for (int j = 0; j < np_; j++) acc = std::fma(at(pn.offA, i, j), xt[j], acc);
    rhs[n + i] = h->rho[i] * (acc - rhs[n + i]);
  }
  for (int j = 0; j < n; j++) rhs[j] = xt[j];
  return BQP_OK;
}

// the same through the whole-GPU layout (M panels, CSR A and A')
int host_grid_kkt_solve(const HostInstance *h, double *rhs) {
  const HostGridL &gd = h->gd;
  if (!gd.built) return BQP_E_UNSUPPORTED;
  const int n = h->n, m = h->m, np_ = h->npad;
  std::vector<double> b(np_, 0.0), xt(np_, 0.0);
  for (int j = 0; j < n; j++) {
    double acc = 0;
    for (int k = gd.trp[j]; k < gd.trp[j + 1]; k++) acc = std::fma(gd.tvl[k], h->rho[gd.tci[k]] * rhs[n + gd.tci[k]], acc);
    b[j] = rhs[j] + acc;
  }
  if (gd.spectral) {       // x~ = V (V' b) at the setup rho (d = 1)
    std::vector<double> c(np_, 0.0);
    for (int r = 0; r < n; r++) {
      double acc = 0;
      for (int j = 0; j < n; j++) acc = std::fma(gd.MP[panel_pos(r, j, np_)], b[j], acc);
      c[r] = acc;
    }
    for (int r = 0; r < n; r++) {
      double acc = 0;
      for (int j = 0; j < n; j++) acc = std::fma(gd.MP[(size_t)gd.offV + panel_pos(r, j, np_)], c[j], acc);
      xt[r] = acc;
    }
  } else
  for (int r = 0; r < n; r++) {
    double acc = 0;
    for (int j = 0; j < n; j++) acc = std::fma(gd.MP[panel_pos(r, j, np_)], b[j], acc);
    xt[r] = acc;
  }
  for (int i = 0; i < m; i++) {
    double acc = 0;
    for (int k = gd.arp[i]; k < gd.arp[i + 1]; k++) acc = std::fma(gd.avl[k], xt[gd.aci[k]], acc);
    rhs[n + i] = h->rho[i] * (acc - rhs[n + i]);
  }
  for (int j = 0; j < n; j++) rhs[j] = xt[j];
  return BQP_OK;
}

// the same through the shared-memory-resident layout (fragment-ordered M, ELL A and A'): a layout check
static inline double small_frag_at(const HostSmall &sm, int off, int np_, int r, int c) {
  const double *X = reinterpret_cast<const double *>(sm.blob.data() + off);
  return X[((size_t)(r >> 3) * (np_ / 4) + (c >> 2)) * 32 + ((r & 7) << 2) + (c & 3)];
}
int host_small_kkt_solve(const HostInstance *h, double *rhs) {
  const HostSmall &sm = h->sm;
  if (!sm.built) return BQP_E_UNSUPPORTED;
  const int n = h->n, m = h->m, np_ = h->npad, mp = sm.mp;
  const double *av = reinterpret_cast<const double *>(sm.blob.data() + sm.offAv), *tv = reinterpret_cast<const double *>(sm.blob.data() + sm.offTv);
  const double *rho = reinterpret_cast<const double *>(sm.blob.data() + sm.offRho);
  const uint16_t *ac = reinterpret_cast<const uint16_t *>(sm.blob.data() + sm.offAc), *tc = reinterpret_cast<const uint16_t *>(sm.blob.data() + sm.offTc);
  std::vector<double> b(np_, 0.0), xt(np_, 0.0);
  for (int j = 0; j < n; j++) {
    double acc = 0;
    for (int k = 0; k < sm.wt; k++) { const int i = tc[(size_t)k * np_ + j]; acc = std::fma(tv[(size_t)k * np_ + j], rho[i] * rhs[n + i], acc); }
    b[j] = rhs[j] + acc;
  }
  for (int r = 0; r < np_; r++) {
    double acc = 0;
    for (int j = 0; j < np_; j++) acc = std::fma(small_frag_at(sm, 0, np_, r, j), b[j], acc);
    xt[r] = acc;
  }
  for (int i = 0; i < m; i++) {
    double acc = 0;
    for (int k = 0; k < sm.wa; k++) acc = std::fma(av[(size_t)k * mp + i], xt[ac[(size_t)k * mp + i]], acc);
    rhs[n + i] = rho[i] * (acc - rhs[n + i]);
  }
  for (int j = 0; j < n; j++) rhs[j] = xt[j];
  return BQP_OK;
}

int host_small_matvec_P(const HostInstance *h, const double *in, double *out) {
  const HostSmall &sm = h->sm;
  if (!sm.built) return BQP_E_UNSUPPORTED;
  const int np_ = h->npad;
  for (int r = 0; r < h->n; r++) {
    double acc = 0;
    for (int c = 0; c < h->n; c++) acc = std::fma(small_frag_at(sm, sm.offP, np_, r, c), in[c], acc);
    out[r] = acc;
  }
  return BQP_OK;
}

int host_panel_matvec_P(const HostInstance *h, const double *in, double *out) {
  const HostPanels &pn = h->pn;
  if (!pn.built) return BQP_E_UNSUPPORTED;
  const int np_ = h->npad;
  for (int r = 0; r < h->n; r++) {
    double acc = 0;
    for (int c = 0; c < h->n; c++) acc = std::fma(pn.data[(size_t)pn.offP + panel_pos(r, c, np_)], in[c], acc);
    out[r] = acc;
  }
  return BQP_OK;
}

int host_stream_matvec_P(const HostInstance *h, const double *in, double *out) {
  const HostStream &st = h->st;
  if (!st.built) return BQP_E_UNSUPPORTED;
  std::vector<double> v(h->npad + 64, 0.0), acc(kStreamWarps * 32);
  for (int j = 0; j < h->n; j++) v[j] = in[j];
  for (int g = st.range[GK_PM][0]; g < st.range[GK_PM][1]; g++) {
    const StreamGroup &G = st.groups[g];
    host_group_apply(st, G, v.data(), acc.data());
    for (int r = 0; r < G.nsl * 32; r++) if (G.row0 + r < h->n) out[G.row0 + r] = acc[r];
  }
  return BQP_OK;
}

}  // namespace bqp
