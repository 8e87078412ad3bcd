// bqp_internal.h -- shared declarations of the B200 batched QP engine (host setup <-> kernels <-> C ABI).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "../../include/bqp.h"

namespace bqp {

constexpr double kInfty = 1e30;          // OSQP_INFTY
constexpr double kMinScaling = 1e-4;     // MIN_SCALING
constexpr double kMaxScaling = 1e4;      // MAX_SCALING
constexpr double kRhoMin = 1e-6;         // RHO_MIN
constexpr double kRhoTol = 1e-4;         // RHO_TOL
constexpr double kRhoEqFactor = 1e3;     // RHO_EQ_OVER_RHO_INEQ
constexpr int kNB = 32;                  // dense-tail block size == warp width
constexpr int kMaxTT = 8;                // widest node tile (nodes solved by one CTA)
constexpr int kMaxThreads = 512;         // largest CTA of the ADMM kernel
constexpr int kMaxSmem = 227 * 1024;     // opt-in dynamic shared memory per CTA on sm_100

// Row-sliced matrix ("panel") format streamed by the kernel.  Rows are grouped into slices of
// 32 (one per lane).  Slice s owns width[s] = sptr[s+1]-sptr[s] "entry rows" of 32 doubles:
// vals[(sptr[s]+j)*32 + lane] is the j-th stored entry of row 32*s+lane, so a warp reads 256
// contiguous bytes per step.  iptr[s] < 0 marks a DENSE slice (entry j multiplies column j, no
// index stream); otherwise idx[(iptr[s]+j)*32 + lane] is the column.  Padding entries are 0.0 / column 0.
struct HostMat {
  int rows = 0, cols = 0, nslices = 0;
  std::vector<int> sptr, iptr;
  std::vector<double> vals;
  std::vector<int> idx;
  long long stream_bytes() const { return (long long)vals.size() * 8 + (long long)idx.size() * 4; }
};

struct DevMat {
  int rows, cols, nslices;
  const int *sptr, *iptr;
  const double *vals;
  const int *idx;
};

// ---- streamed layout of the sm_100a TMA kernel (bqp_stream.cu) ------------------------------------------
// Everything one ADMM iteration touches is laid out ONCE, in consumption order, as a sequence of GROUPS.
// A group is up to kStreamWarps row slices (32 rows each; slice i belongs to consumer warp i, warps 4q..4q+3 form
// QUAD q).  Its data is a sequence of fixed-size STAGES, one TMA bulk copy each:
//   stage = vals[4 warps][kKC entry-rows][32 lanes] (f64)  (+ idx[4][kKC][32] (i32) when the group is sparse)
// ordered quad by quad (all qch[0] chunks of quad 0, then quad 1, ...): every quad is an independent stream with
// its own ring of shared-memory slots and its own producer lane.  Dense groups multiply entry-row e of quad q
// with element in_off + qcol0[q] + e of the input vector; sparse groups with in_off + idx.
constexpr int kKC = 16;
constexpr int kStreamWarps = 12;                    // consumer warps of the TMA kernel (3 quads) = slices per group
constexpr int kStageValBytes = 4 * kKC * 32 * 8;   // 8 KiB
constexpr int kStageIdxBytes = 4 * kKC * 32 * 4;   // 4 KiB
enum { GK_AT = 0, GK_FWD_D = 1, GK_FWD_U = 2, GK_BWD_D = 3, GK_BWD_U = 4, GK_AB = 5, GK_PM = 6 };
struct StreamGroup {
  int kind, row0, nsl, in_off, sparse;
  int qch[4], qcol0[4];
  long long data_off;   // bytes into the instance's stream buffer
};
struct HostStream {
  bool built = false;
  int tri_nb = 0;
  std::vector<StreamGroup> groups;
  int range[7][2] = {};              // [kind] -> [first, last) group index (FWD/BWD kinds D and U interleaved: see fw, bw)
  int fw[2] = {0, 0}, bw[2] = {0, 0};
  std::vector<unsigned char> data;
  long long iter_bytes = 0, check_bytes = 0;   // bytes streamed by one iteration / one termination check
  int slot_bytes = kStageValBytes;
};

// ---- row-panel layout of the fused single-pass kernel (bqp_panel.cu) -------------------------------------
// For problems whose A is dense enough to be stored dense (every BASELINE random_miqp config at density 0.7) one ADMM
// iteration is restated as   x~ = M b,  z~ = A x~,  b' = sigma x - q + A'(rho z - y)   with M = (P + sigma I + A' rho A)^-1
// formed explicitly on the host, and A streamed ONCE per iteration: each PANEL (kPanelRows rows x npad columns) is
// used while it sits in shared memory first for its rows of A x~ and then, after the per-row z/y update, for its
// contribution to A' w.  Consumer warp w owns the column tile 32w..32w+31 and multiplies it with FP64 mma.sync.m8n8k4
// (the up to kPanelT nodes of the tile are the N dimension), so a tile of a panel is stored as the 8 A-fragments of
// pass 1:   k*16*npad + w*512 + h*256 + ks*32 + lane   holds   X[16k + 8h + (lane>>2)][32w + 4ks + (lane&3)]
// (a warp-wide 8-byte load is 256 contiguous bytes; the transposed fragments of pass 2 read the same 2 KB without bank
// conflicts).  The tiles of one panel are contiguous in w, so a CTA streaming columns [32 w0, 32 (w0+nwc)) of every
// panel issues one TMA bulk copy per panel.  Matrices in stream order: M (npad rows), A (m rows padded to 8), P full
// symmetric (npad rows; termination checks and the final objective only).
constexpr int kPanelRows = 8;                       // one 8-row mma tile per panel (16 also works: two tiles per hand-off)
constexpr int kPanelT = 8;                          // nodes per tile of the panel kernel (N of the mma)
constexpr int kPanelMaxWarps = 16;                  // column tiles of 32 (npad <= 512), one consumer warp each
constexpr int kPanelCtaWarps = 8;                   // consumer warps per CTA; wider problems run as a cluster pair of CTAs
#ifndef BQP_P1_TILES
#define BQP_P1_TILES 1
#endif
#ifndef BQP_P1_SETS
#define BQP_P1_SETS 1
#endif
constexpr int kPanelP1Tiles = BQP_P1_TILES;         // column tiles per pass-1 warp (pass-2 warps own two)
// pass-1 warp SETS: set s handles the panels whose global panel counter is congruent to s.  With 2 sets the per-panel
// synchronisation tail of one set (mbarrier wait, flow control, partial store + DSMEM copy) overlaps the other set's mma work
constexpr int kPanelP1Sets = BQP_P1_SETS;
constexpr int kPanelUpdWarps = 3;                   // update warps (row-space z / y / x updates), panels dealt round-robin 
// warps of a panel-kernel CTA that streams `nwc` column tiles: pass-1 sets, pass-2 warps, update warps, producer
constexpr int panel_cta_warps(int nwc) { return (nwc + kPanelP1Tiles - 1) / kPanelP1Tiles * kPanelP1Sets + (nwc + 1) / 2 + kPanelUpdWarps + 1; }
struct HostPanels {
  bool built = false;
  int nw = 0, npm = 0, npa = 0;                     // column tiles; panels of M (and P); panels of A
  long long panel_doubles = 0, offA = 0, offP = 0, offV = 0;  // doubles
  std::vector<double> data;
  // adaptive rho: the M slot holds V' and V follows P: K(rho)^-1 = V diag(1 / (1 + (rho - rho0) mu)) V' (bqp_setup.cpp spectral_factors)
  bool spectral = false;
  std::vector<double> mu;
  long long panel_bytes() const { return panel_doubles * 8; }
  long long iter_bytes() const { return (long long)((spectral ? 2 : 1) * npm + npa) * panel_bytes(); }          // M (or V', V) + A
  long long check_bytes() const { return (long long)(2 * npa + npm) * panel_bytes(); }     // A twice + P
  long long launch_bytes() const { return (long long)(npa + npm) * panel_bytes(); }        // prologue A pass + objective P pass
};

// ---- row-split cluster kernel (bqp_rows.cu): same panel stream, CTA r of a cluster of C owns the panels k = r (mod C)
#ifndef BQP_ROWS_GROUPS
#define BQP_ROWS_GROUPS 2
#endif
constexpr int kRowsT = 8;                           // nodes per tile (N of the mma)
constexpr int kRowsGroupWarps = 4;                  // warps per group, each multiplying a quarter of the column tiles
constexpr int kRowsGroups = BQP_ROWS_GROUPS;        // groups per CTA, each taking whole panels through all stages
constexpr int kRowsThreads = (kRowsGroups * kRowsGroupWarps + 4) * 32;   // + the producer warpgroup (one lane issues the TMA copies)

// ---- whole-GPU kernel for ONE large tile (bqp_grid.cu): dense-reduced problems wider than the rows kernel's 512 columns
// (BASELINE config 4: n = 2000, m = 4200, A 5 % dense).  Same restated iteration  x~ = M b,  z~ = A x~,  b' = sigma x - q +
// A'(rho z - y)  with the explicit reduced inverse M, but every SM of the GPU works on the same tile: M and P as 8-row
// fragment-ordered panels (the rows kernel's panel format, npad columns), A and A' as CSR (f64 value + i32 column).
struct HostGridL {
  bool built = false;
  int npm = 0;                                      // 8-row panels of M (and of P)
  std::vector<double> MP;                           // [M panels | P panels], spectral (adaptive rho): [V' panels | V panels | P panels]
  long long offP = 0, offV = 0;                     // doubles
  bool spectral = false;                            // K(rho)^-1 = V diag(1 / (1 + (rho - rho0) mu)) V' (bqp_setup.cpp build_grid)
  std::vector<double> mu;                           // generalised eigenvalues, npad entries
  std::vector<int> arp, aci, trp, tci;              // CSR of A (m rows) and of A' (n rows)
  std::vector<double> avl, tvl;
  long long iter_bytes(int npad) const { return (long long)npad * npad * 8 + 12LL * (long long)(avl.size() + tvl.size()); }
};

// ---- shared-memory-resident kernel for small problems (bqp_small.cu): npad <= 64 (BASELINE config 3: the power-converter
// MPC, n = 60, m = 150; config 5: the max-iter pickles, n = 20, m = 70).  Same restated iteration as the dense kernels
// (x~ = M b with the explicit reduced inverse, z~ = A x~, b' = sigma x - q + A'(rho z - y)); everything the ADMM loop reads
// is ONE contiguous blob that a single TMA bulk copy stages into shared memory at the start of the launch and that stays there:
//   Mf, Pf   M and P as FP64 mma.m8n8k4 A-fragments: [panel p][k-step s][lane] = X[8 p + (lane >> 2)][4 s + (lane & 3)]
//   Av, Ac   A by rows as ELL, entry-major: Av[k * mp + r] (f64), Ac[k * mp + r] (u16 column), k < wa; padding 0.0 / column 0
//   Tv, Tc   A' by rows (= A by columns) the same way: Tv[k * npad + j], Tc[k * npad + j] (u16 row of A), k < wt
//   rho, rinv, E, Einv per row of A (mp entries each)
struct HostSmall {
  bool built = false;
  int wa = 0, wt = 0, mp = 0;                       // ELL widths of A and A'; m rounded up to 8
  int offP = 0, offAv = 0, offTv = 0, offRho = 0, offRinv = 0, offE = 0, offEinv = 0, offAc = 0, offTc = 0, bytes = 0;   // byte offsets into the blob (M at 0)
  std::vector<unsigned char> blob;
  long long iter_bytes() const { return bytes; }    // what one iteration reads -- from shared memory, not from HBM
};

// Everything the host computes once per (P, A): scaled data, rho typing, the LDL^T factor of the
// KKT matrix in constraints-first order (L = [[I,0],[L21,L22]] with L21 = -A' diag(rho)
// streamed as the A panels and the dense trailing supernode L22 D2 L22' = P + sigma I + A' diag(rho) A).
struct HostInstance {
  int n = 0, m = 0, npad = 0, n_int = 0;
  bqp_settings s{};
  std::vector<double> D, Dinv, E, Einv;
  double c = 1, cinv = 1;
  std::vector<double> q;                 // scaled linear cost
  double nq = 0;                         // || Dinv q ||_inf (scaled q), constant between update_q calls
  std::vector<double> rho, rho_inv;
  std::vector<int> rtype;                // osqp constr_type per row at setup: -1 loose (RHO_MIN), 0 inequality (rho), 1 equality (1e3 rho)
  std::vector<int> i_idx;
  HostMat At, Ab, Pm;                    // A' (n x m), A (m x n), P full symmetric (n x n); all scaled
  // Blocked dense tail.  Lcol: block column J = rows J*32..npad of 32 columns, column-major (ld = npad-J*32);
  // Lrow: block row J = 32 rows of columns 0..(J+1)*32, row-major (ld = (J+1)*32).  In BOTH, the 32x32
  // diagonal block holds the strictly-lower part of inv(L22[J,J]) (unit diagonal implicit), so the
  // in-block substitution is a mat-vec.
  std::vector<double> Lcol, Lrow, D2inv;
  HostStream st;                         // streamed layout (built for problems large enough for the TMA kernel)
  HostPanels pn;                         // row-panel layout of the fused single-pass kernel (dense A, npad <= 512)
  HostGridL gd;                          // whole-GPU layout (npad > 512, memory permitting)
  HostSmall sm;                          // shared-memory-resident layout (npad <= 64)
  // guard of the explicit reduced inverse: largest relative difference, over a few probe right-hand sides, between a KKT
  // solve through the panels (x~ = M b) and through the LDL' substitution; NaN when no panel layout was tried.  Above the
  // threshold (1e-10, BQP_INVERSE_TOL) the panel layout is dropped and the problem runs on the LDL' kernels
  double pn_inverse_error = NAN;
  bool pn_rejected = false;
  // eq_rho == 2 (per-node rho typing of the integer-bound rows, what osqp >= 0.4 does in update_bounds): rows i_idx[k] of M
  // in natural order, [n_int][npad] -- the columns of the Woodbury correction of the explicit inverse (bqp_api.cu)
  std::vector<double> mint;
  long long factor_bytes() const {   // bytes one ADMM iteration streams: A', L fwd, L bwd, A, D2inv
    return (long long)(Lcol.size() + Lrow.size() + D2inv.size()) * 8 + At.stream_bytes() + Ab.stream_bytes();
  }
  long long check_bytes() const {    // extra bytes of one termination check: P, A', A twice each (+ certificates)
    return 2 * (Pm.stream_bytes() + At.stream_bytes() + Ab.stream_bytes());
  }
};

int host_setup(const bqp_problem *p, const bqp_settings *s, HostInstance *out);   // bqp_setup.cpp
void host_rescale_q(HostInstance *h, const double *q);
void host_kkt_solve(const HostInstance *h, double *rhs_xz);
void host_matvec(const HostMat &M, const double *in, double *out);
int host_stream_kkt_solve(const HostInstance *h, double *rhs_xz);
int host_stream_matvec_P(const HostInstance *h, const double *in, double *out);
int host_panel_kkt_solve(const HostInstance *h, double *rhs_xz);
int host_grid_kkt_solve(const HostInstance *h, double *rhs_xz);
int host_small_kkt_solve(const HostInstance *h, double *rhs_xz);
int host_small_matvec_P(const HostInstance *h, const double *in, double *out);
int host_panel_matvec_P(const HostInstance *h, const double *in, double *out);
double host_panel_M(const HostInstance *h, int r, int c);      // entry (r, c) of the explicit reduced inverse

struct DevInstance {
  int n, m, npad, n_int;
  // streamed layout (TMA kernel)
  const unsigned char *stream; const StreamGroup *groups;
  int g_at[2], g_fw[2], g_bw[2], g_ab[2], g_pm[2];
  int w_in_stage;   // 1: every A' group is dense -> its input vector chunks ride in the TMA stages (no m x T vector in smem)
  // row-panel layout (fused single-pass kernel)
  const double *pstream; int p_nw, p_npm, p_npa; long long p_panel_doubles, p_offA, p_offP, p_offV;
  const double *p_mint; int eq2; double rho_base;      // eq_rho == 2: rows of M of the integer variables; untyped rho
  // whole-GPU layout (bqp_grid.cu)
  const double *g_M, *g_P; const int *g_arp, *g_aci, *g_trp, *g_tci; const double *g_avl, *g_tvl; int g_npm;
  // adaptive rho (spectral form): g_M = V' panels, g_V = V panels, g_mu = eigenvalues, g_rtype = row types
  const double *g_V, *g_mu; const int *g_rtype; int adaptive, adapt_interval; double adapt_tol;
  // shared-memory-resident layout (bqp_small.cu): the blob and the byte offsets of its sections
  const unsigned char *s_blob; int s_bytes, s_wa, s_wt, s_mp, s_offP, s_offAv, s_offTv, s_offRho, s_offRinv, s_offE, s_offEinv, s_offAc, s_offTc;
  DevMat At, Ab, Pm;
  const double *Lcol, *Lrow, *D2inv;
  const double *rho, *rho_inv, *q, *D, *Dinv, *E, *Einv;
  const int *i_idx;
  double c, cinv, nq;
  double sigma, alpha, eps_abs, eps_rel, eps_pinf, eps_dinf;
  int max_iter, check_every;
};

// One CTA solves one tile = up to kMaxTT nodes of one instance.
struct DevTile {
  int inst, nn;
  int iter_begin, iter_end;      // this launch runs ADMM iterations iter_begin+1 .. iter_end of the tile's nodes (a round)
  int node[kMaxTT];              // caller-side node index (scalar outputs)
  long long in_off[kMaxTT];      // doubles into the packed input buffer: l[m] u[m] x0[n] y0[m]
  long long out_off[kMaxTT];     // doubles into the packed output buffer: x[n] y[m]
  long long state_off[kMaxTT];   // doubles into the ADMM state buffer (scaled x[n] z[m] y[m]) used to resume a node
  long long corr_off[kMaxTT];    // eq_rho == 2: doubles into the correction buffer ([nS][k_s ...][j_s ...][G nS x nS]), -1: none
  long long work_off;            // doubles into the state workspace
};

struct NodeScalars {
  int status, iters;
  double obj, pri_res, dua_res, lower;
};

// state workspace of one tile, [row][T] node-fastest: x, dx (n rows each); z, y, l, u, dy (m rows each); P x scratch
// (n rows); then, 16-byte aligned, the A' input vector w = rho z - y (m + 32 rows, zero padded) for the TMA kernel
#ifdef __CUDACC__
#define BQP_HD __host__ __device__
#else
#define BQP_HD
#endif
BQP_HD inline size_t tile_w_offset(int n, int m, int tt) { return ((size_t)tt * (5 * (size_t)m + 3 * (size_t)n) + 1) & ~size_t(1); }
inline size_t tile_work_doubles(int n, int m, int tt) { return (tile_w_offset(n, m, tt) + (size_t)tt * ((size_t)m + 32) + 1) & ~size_t(1); }
// panel kernel, per CTA, [row][kPanelT]: z, y, l, u, dy (m padded to whole panels); dx, Px, A'y, A'dy, P dx, x, x snapshot (npad rows each)
BQP_HD inline size_t panel_work_doubles(int npad, int m) { return (size_t)kPanelT * (5 * (size_t)((m + kPanelRows - 1) / kPanelRows * kPanelRows) + 7 * (size_t)npad); }
size_t panel_smem_bytes(int npad, int nslots, int cs);                                // bqp_panel.cu
int launch_admm_panel(int cs, int nw_max, int nslots, double *d_state, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles,
                      const double *d_in, double *d_out, double *d_work, NodeScalars *d_ns, int *d_tile_iters,
                      size_t smem_bytes, void *stream);
// rows kernel, per tile, [row][8]: z, y, l, u, dy, A x, A dx (m padded to 8); x, dx, P x, objective operand, x snapshot (npad
// rows each); column-space partials of the three check passes per (CTA, group)
BQP_HD inline size_t rows_work_doubles(int npad, int m, int cs) {
  return (size_t)kRowsT * (7 * (size_t)((m + 7) / 8 * 8) + (5 + 3 * (size_t)cs * kRowsGroups) * (size_t)npad);
}
size_t rows_smem_bytes(int npad, int nslots, int cs);                                 // bqp_rows.cu
// ext != 0: some problem of the launch uses per-node rho typing (eq_rho == 2) or adaptive rho (the instantiation that carries that code)
int launch_admm_rows(int cs, int ext, int nslots, double *d_state, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles,
                     const double *d_in, double *d_out, double *d_work, NodeScalars *d_ns, int *d_tile_iters, size_t smem_bytes,
                     const double *d_corr, void *stream);
// whole-GPU kernel: per tile, [row][8]: b, x~, x, dx, P x, P dx, objective operand, d . (V' b) (npad rows each); w, y, projected
// dy (m rows each); per-CTA partial norms [G][24][8]
BQP_HD inline size_t grid_work_doubles(int npad, int m, int nctas) {
  return (size_t)8 * (8 * (size_t)npad + 3 * (size_t)((m + 7) / 8 * 8) + 24 * (size_t)nctas);
}
size_t grid_smem_bytes(int npad, int m, int n, int nctas);                            // bqp_grid.cu
int grid_max_ctas(int device, size_t smem_bytes);                                     // co-resident CTAs of the cooperative launch
int launch_admm_grid(int nctas, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles, const double *d_in, double *d_out,
                     double *d_work, NodeScalars *d_ns, int *d_tile_iters, unsigned *d_barrier, size_t smem_bytes, void *stream);
size_t small_smem_bytes(int npad, int m, int blob_bytes);                             // bqp_small.cu
int launch_admm_small(int npad_mask /* bit 0: npad 32 present, bit 1: npad 64 */, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles, const double *d_in, double *d_out,
                      NodeScalars *d_ns, int *d_tile_iters, size_t smem_bytes, void *stream);
size_t tile_smem_bytes(int n, int m, int tt, int threads);                           // bqp_kernels.cu
size_t stream_smem_bytes(int n, int m, int tt, int slot_bytes, int nslots, int w_in_stage);          // bqp_stream.cu
int launch_admm_stream(int tt, int slot_bytes, int nslots, int w_in_stage, double *d_state, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles,
                       const double *d_in, double *d_out, double *d_work, NodeScalars *d_ns, int *d_tile_iters,
                       size_t smem_bytes, void *stream);
int launch_admm(int tt, int threads, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles, const double *d_in,
                double *d_out, double *d_work, NodeScalars *d_ns, int *d_tile_iters, size_t smem_bytes, void *stream);

}  // namespace bqp
