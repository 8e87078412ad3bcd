// bqp_rows.cu -- row-split cluster ADMM kernel for sm_100a (dense A, npad <= 512): the config-2 hot path of round 2.
//
// Same tile ownership and iteration as bqp_panel.cu (a tile = up to 8 B&B leaves of one problem, whole OSQP loop
// in-kernel, /root/reference/miosqp/node.py:96-143;  x~ = M b,  z~ = A x~,  z/y update,  b' = sigma x - q + A'(rho z - y)
// with A streamed from HBM once per iteration), but a different decomposition.  bqp_panel.cu splits the COLUMNS of every
// 8-row panel over the two CTAs of a pair and hands each panel through three warp roles (pass-1 warps -> update warps ->
// pass-2 warps); measured (profiles/r02_panel_role_timers.txt) the per-panel hand-off chain, not HBM or the FP64 pipe,
// bounds it: 1200 cycles per half-width panel, 121 us per ADMM iteration whether the GPU runs 1 tile or 74.
// Here the ROWS are split: CTA r of a cluster of C owns the panels k = r (mod C), full width, and inside a CTA a GROUP
// of 4 warps takes one panel through all its stages without handing it to anybody:
//     wait TMA -> pass 1 (warp w multiplies its quarter of the columns: FP64 mma.sync.m8n8k4, the 8 leaves of the tile are
//     the N dimension) -> the 4 partial 8x8 blocks meet in shared memory (one 128-thread named barrier) -> every warp of
//     the group applies the row update (projection, dual update) redundantly in registers -> pass 2 (A_panel' w into the
//     warp's column accumulators) -> slot released.
// G groups work on different panels at the same time, so one group's barrier / update phase overlaps the others' mma work.
// Nothing crosses CTAs per panel.  Per ITERATION the cluster exchanges, over DSMEM (st.async completing transaction
// bytes on the receiver's mbarrier: data and notification in one operation, no cluster barrier in the loop):
//     x~ rows as they are produced during the M pass (all-gather, overlapped with the pass),
//     the column-space partial sums of A'w after the A pass (reduce-scatter by column chunk, fixed rank order), and the
//     finished chunks of b' (all-gather).
// Termination checks (every check_termination iterations) are not on the fast path: the three check passes write their
// raw products (A x, A dx, P x rows; per-(CTA, group) partials of A'y, A'dy, P dx) to the tile's workspace in global
// memory, the cluster synchronises, and every CTA computes all norms redundantly in one canonical order -- identical
// decisions in every CTA, nothing to broadcast.
// Every wait is bounded: a broken protocol traps instead of hanging the GPU.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>

#include <atomic>

#include "bqp_internal.h"

namespace bqp {

namespace {

constexpr int T8 = kRowsT;
constexpr int kW = kRowsGroupWarps;        // warps per group: each owns a quarter of the column tiles
constexpr int kG = kRowsGroups;            // groups per CTA
constexpr int kTPW = 4;                    // column tiles per warp (npad <= 512)
constexpr int kCons = kG * kW;             // consumer warps
constexpr int kConsThreads = kCons * 32;
constexpr int kTileD = 256;                // doubles of one column tile of one panel (8 rows x 32 columns)
constexpr int kFinN = 16;                 // quantities of one canonical reduction
constexpr int kFinAll = 24;               // + the 7 scaled norms of osqp's compute_rho_estimate (adaptive rho)
constexpr int kMaxS = 64;                  // eq_rho == 2: re-typed integer rows per node the scratch holds (n_int <= 64 with the dense kernels' part buffer)
constexpr int kXR = 16;                    // x-iterate elements a consumer thread keeps in registers (npad * 8 / 256 at most)
// Registers: each SM sub-partition holds 16 K registers and hosts every fourth warp, so 12 warps get 168 registers each at
// launch.  The consumer warps need ~200 (B fragments of 128 columns, 32 accumulators, 8 mma chains), the producer warpgroup
// next to nothing: setmaxnreg moves the budget (2 x 240 + 24 registers x 32 lanes per sub-partition).
constexpr int kConsRegs = 240, kProdRegs = 24;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) {
#ifdef BQP_ROWS_DEBUG
      if ((threadIdx.x & 31) == 0) printf("TIMEOUT blk %d warp %d bar %u parity %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), bar, parity);
#endif
      __trap();   // ~10 s at 2 GHz
    }
  }
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void *src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// 16 bytes into a peer CTA's shared memory, completing 16 transaction bytes on the peer's mbarrier when they have landed
__device__ __forceinline__ void st_async_remote_v2(uint32_t raddr, double a, double b, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(raddr), "d"(a), "d"(b), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  // volatile: keeps the issue order as written (independent accumulator chains interleaved); without it ptxas groups the mma
  // of one chain back to back to save fragment registers and exposes the mma latency
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 ldcg2(const double *p) { return __ldcg(reinterpret_cast<const double2 *>(p)); }

struct RowsShared {
  DevInstance I;
  DevTile tile;
  double fin[kFinAll][T8];
  int status[T8], iters[T8], newly[T8];
  double dist[T8];
  int remaining;
  // adaptive rho: the leaf's current rho (osqp settings->rho after osqp_update_rho), 1 / rho and 1 / (1e3 rho); a change is
  // announced to the producer warpgroup through rho_changed (one more A pass rebuilds the right-hand side)
  double rho_t[T8], rinv_t[T8], rinveq_t[T8];
  int rho_changed;
};

enum { RM_A_INIT = 0, RM_A_RESUME, RM_M, RM_A_ITER, RM_CHK_A1, RM_CHK_A2, RM_CHK_P, RM_OBJ_P };

// per-tile workspace in global memory (L2 resident), every vector [row][8 nodes]
struct Work {
  double *gz, *gy, *gl, *gu, *gdy, *gax, *gadx;       // m8 rows
  double *gx, *gdx, *gpx, *gxo;                       // npad rows
  double *gsx;                                        // [8 nodes][npad]: unscaled snapshot of terminated nodes
  double *parts;                                      // [3 check passes][C * kG][npad][8]: column-space partials of the checks
};

#ifdef BQP_ROWS_DEBUG
#define RSTAMP(X, i) do { const long long now_ = clock64(); (X).tacc[MODE == RM_M ? 0 : 1][i] += now_ - (X).tlast; (X).tlast = now_; } while (0)
#else
#define RSTAMP(X, i) do { } while (0)
#endif
template <int CS>
struct Ctx {
#ifdef BQP_ROWS_DEBUG
  long long tacc[2][6] = {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}}, tlast = 0; int npan[2] = {0, 0};
#endif
  RowsShared *S;
  Work Wk;
  uint32_t full, empty, xbar, rbar, cbar, vbar;       // mbarrier addresses (shared window)
  double *colb, *colx, *recv, *part;                  // b (M-pass operand) | x~ (A-pass operand) | reduce-scatter inbox | group partials
  unsigned char *ring;
  uint32_t ring_u32;
  int nslots, slot_bytes;
  int NW, np, n, m, npm, npa;
  int rank;
  int lane, warp, grp, wi;                            // warp in CTA, group, warp in group
  int tile0, ntl;                                     // this warp's column tiles
  int chunk_t0, chunk_nt;                             // column tiles of the chunk this CTA finalises
  int cnt;                                            // CTA-local panel counter across passes (ring position of the next pass)
  int cmine, slot; uint32_t phase;                    // next counter value handled by this warp's group, its ring slot and phase
  int gord;                                           // panels this group has processed (owner-warp rotation, partial double buffer)
  int pset;                                           // which half of this CTA's panels (even / odd local index) the group took in the last pass
  uint32_t xph, rph, cph, vph;                        // phases of the cross-CTA barriers
};

template <int OP>   // 0 max, 1 sum, 2 min
__device__ __forceinline__ double red_op(double v, double w) { return OP == 0 ? fmax(v, w) : (OP == 1 ? v + w : fmin(v, w)); }

template <int CS>
__device__ __forceinline__ void cluster_sync_all() {
  if constexpr (CS > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
}

// B fragments of a column vector v ([column][8 nodes]) for this warp's tiles: bx[tl][ks] = v[32 (tile0 + tl) + 4 ks + (lane & 3)][lane >> 2]
template <bool kGlobal>
__device__ __forceinline__ void load_bx(double (&bx)[kTPW][8], const double *v, int tile0, int ntl, int lane) {
  const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int tl = 0; tl < kTPW; tl++)
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {
      double val = 0.0;
      if (tl < ntl) {
        const double *p = v + (size_t)(32 * (tile0 + tl) + 4 * ks + tq) * T8 + gq;
        val = kGlobal ? __ldcg(p) : *p;
      }
      bx[tl][ks] = val;
    }
}

// One pass over the panels this CTA owns of one matrix.  MODE selects the row functor; passes with a second half
// accumulate A_panel' w into acc (C fragments: column 32 (tile0 + tl) + 8 mt + (lane >> 2), nodes 2 (lane & 3) + {0, 1}).
// EXT: the launch holds problems with per-node rho typing (eq_rho == 2) or adaptive rho; the plain instantiation carries none of
// that code (registers, branches in the row functor)
template <int CS, int MODE, bool EXT>
__device__ __forceinline__ void rows_pass(Ctx<CS> &X, int npanels, const double (&bx)[kTPW][8], double (&acc)[kTPW][4][2], bool do_check,
                                          double *mdst = nullptr, bool dscale = false, uint32_t mbar = 0) {
  // RM_M only: mdst = the vector the product rows go to (this CTA's copy and every peer's); dscale = multiply row j of
  // leaf t by 1 / (1 + (rho_t - rho0) mu_j) (adaptive rho: the first of the two passes of x~ = V (d . (V' b))); mbar = the
  // peers' mbarrier the remote rows complete their bytes on
  const DevInstance &I = X.S->I;
  const Work &W = X.Wk;
  constexpr bool kIsA = (MODE == RM_A_INIT || MODE == RM_A_RESUME || MODE == RM_A_ITER || MODE == RM_CHK_A1 || MODE == RM_CHK_A2);
  constexpr bool kPass2 = kIsA || MODE == RM_CHK_P;
  constexpr bool kPass1 = MODE != RM_A_RESUME;
  const int lane = X.lane, gq = lane >> 2, tq = lane & 3, r = gq;
  const int a2 = (gq >> 2) * 32 + tq * 4 + (gq & 3);
  const double alpha = I.alpha, oma = 1.0 - I.alpha;
  const int m = X.m;
  if constexpr (kPass2) {
#pragma unroll
    for (int tl = 0; tl < kTPW; tl++)
#pragma unroll
      for (int mt = 0; mt < 4; mt++) acc[tl][mt][0] = acc[tl][mt][1] = 0.0;
  }
  const int npc = (npanels - X.rank + CS - 1) / CS;     // panels of this CTA: k = rank + jl * CS
  // the two groups alternate over the CTA's panels by a counter that runs across passes: which group takes the even local
  // panels depends on how many panels came before.  Results that are summed per group are filed under the HALF, not the group,
  // so that their order of addition does not depend on the history of the launch (rounds, extra passes)
  X.pset = (X.cmine - X.cnt) & 1;
  const int src_b = (lane & 3) * 4 + (lane >> 3);        // shuffle source of the B fragment of w (rows 0..3); rows 4..7: + 16
  for (; X.cmine < X.cnt + npc; X.cmine += kG) {
    const int jl = X.cmine - X.cnt;
    const int k = X.rank + jl * CS;
    const int slot = X.slot;
    const uint32_t phase = X.phase;
    X.slot += kG;
    if (X.slot >= X.nslots) { X.slot -= X.nslots; X.phase ^= 1u; }     // kG <= nslots
    const int row = k * 8 + r;
    const bool live = kIsA ? row < m : row < X.np;
    const bool owner = (X.gord % kW) == X.wi;           // the warp of the group that stores this panel's row results
    const int e2 = row * (T8 / 2) + tq;                 // double2 index of (row, nodes 2 tq, 2 tq + 1)
    // operands of the row functor that do not depend on the products: fetched before the mma work so that their L2 latency hides
    double2 s0 = make_double2(0, 0), s1 = s0, s2 = s0, s3 = s0;
    double rho = 0.0, rinv = 0.0;
    if (live) {
      auto ld2 = [&](const double *v) { return __ldcg(reinterpret_cast<const double2 *>(v) + e2); };
      if constexpr (MODE == RM_A_ITER) { s0 = ld2(W.gz); s1 = ld2(W.gy); s2 = ld2(W.gl); s3 = ld2(W.gu); rho = __ldg(I.rho + row); rinv = __ldg(I.rho_inv + row); }
      if constexpr (MODE == RM_A_INIT) { s1 = ld2(W.gy); rho = __ldg(I.rho + row); }
      if constexpr (MODE == RM_A_RESUME) { s0 = ld2(W.gz); s1 = ld2(W.gy); rho = __ldg(I.rho + row); }
      if constexpr (MODE == RM_CHK_A1) { s1 = ld2(W.gy); }
      if constexpr (MODE == RM_CHK_A2) { s0 = ld2(W.gdy); s2 = ld2(W.gl); s3 = ld2(W.gu); }
      if constexpr (MODE == RM_CHK_P) { s0 = ld2(W.gdx); }
      if constexpr (MODE == RM_M) {
        if (EXT && dscale && owner) {
          const double mu = __ldg(I.g_mu + row);
          s0.x = 1.0 / (1.0 + (X.S->rho_t[2 * tq] - I.rho_base) * mu); s0.y = 1.0 / (1.0 + (X.S->rho_t[2 * tq + 1] - I.rho_base) * mu);
        }
      }
    }
    RSTAMP(X, 5);
    double rho2[2] = {rho, rho}, rinv2[2] = {rinv, rinv};
    if constexpr (MODE == RM_A_ITER || MODE == RM_A_INIT || MODE == RM_A_RESUME) {
      // eq_rho == 2: the integer-bound rows are typed per node from the node's own (scaled) bounds, as osqp >= 0.4 does
      if (EXT && I.eq2 && live && row >= m - I.n_int) {
        double2 lo2, up2;
        if constexpr (MODE == RM_A_ITER) { lo2 = s2; up2 = s3; }
        else { lo2 = __ldcg(reinterpret_cast<const double2 *>(W.gl) + e2); up2 = __ldcg(reinterpret_cast<const double2 *>(W.gu) + e2); }
        const double lo_[2] = {lo2.x, lo2.y}, up_[2] = {up2.x, up2.y};
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double rr = I.rho_base;
          if (lo_[i] < -kInfty * kMinScaling && up_[i] > kInfty * kMinScaling) rr = kRhoMin;
          else if (up_[i] - lo_[i] < kRhoTol) rr = kRhoEqFactor * I.rho_base;
          rho2[i] = rr; rinv2[i] = 1.0 / rr;
        }
      }
      // adaptive rho: inequality / equality rows follow the leaf's own rho (osqp_update_rho), loose rows stay at RHO_MIN
      if (EXT && I.adaptive && live) {
        const int ty = __ldg(I.g_rtype + row);
#pragma unroll
        for (int i = 0; i < 2; i++) {
          const int t = 2 * tq + i;
          if (ty == 0) { rho2[i] = X.S->rho_t[t]; rinv2[i] = X.S->rinv_t[t]; }
          else if (ty == 1) { rho2[i] = kRhoEqFactor * X.S->rho_t[t]; rinv2[i] = X.S->rinveq_t[t]; }
        }
      }
    }
    mbar_wait(X.full + 8u * slot, phase);
    RSTAMP(X, 0);
    const double *slotp = reinterpret_cast<const double *>(X.ring + (size_t)slot * X.slot_bytes) + X.tile0 * kTileD;
    double sum[2] = {0.0, 0.0};
    if constexpr (kPass1) {
      double cc[8][2];
#pragma unroll
      for (int ks = 0; ks < 8; ks++) cc[ks][0] = cc[ks][1] = 0.0;
      const double *sp = slotp + lane;
      if (X.ntl == kTPW) {
        // all four tiles (npad = 512): no branches, the fragments of tile tl + 1 are loaded while tile tl multiplies
        double a[2][8];
#pragma unroll
        for (int ks = 0; ks < 8; ks++) a[0][ks] = sp[ks * 32];
#pragma unroll
        for (int tl = 0; tl < kTPW; tl++) {
          if (tl + 1 < kTPW) {
#pragma unroll
            for (int ks = 0; ks < 8; ks++) a[(tl + 1) & 1][ks] = sp[(tl + 1) * kTileD + ks * 32];
          }
#pragma unroll
          for (int ks = 0; ks < 8; ks++) dmma(cc[ks], a[tl & 1][ks], bx[tl][ks]);
        }
      } else {
#pragma unroll
        for (int tl = 0; tl < kTPW; tl++) {
          if (tl < X.ntl) {
#pragma unroll
            for (int ks = 0; ks < 8; ks++) dmma(cc[ks], sp[tl * kTileD + ks * 32], bx[tl][ks]);
          }
        }
      }
      double c0[2];
#pragma unroll
      for (int i = 0; i < 2; i++) c0[i] = ((cc[0][i] + cc[1][i]) + (cc[2][i] + cc[3][i])) + ((cc[4][i] + cc[5][i]) + (cc[6][i] + cc[7][i]));
      // the four warps' partial blocks meet in shared memory (double-buffered by the group's panel parity)
      double2 *pb = reinterpret_cast<double2 *>(X.part) + ((X.grp * 2 + (X.gord & 1)) * kW) * 32;
      pb[X.wi * 32 + lane] = make_double2(c0[0], c0[1]);
      if constexpr (!kPass2) {   // no second half: the slot is free as soon as the fragments are in registers
        __syncwarp();
        if (lane == 0) mbar_arrive(X.empty + 8u * slot);
      }
      RSTAMP(X, 1);
      named_bar(1 + X.grp, kW * 32);
      RSTAMP(X, 2);
      const double2 p0 = pb[lane], p1 = pb[32 + lane], p2 = pb[64 + lane], p3 = pb[96 + lane];
      sum[0] = (p0.x + p1.x) + (p2.x + p3.x);
      sum[1] = (p0.y + p1.y) + (p2.y + p3.y);
    }
    // ---- row functor (every warp of the group redundantly; only `owner` stores)
    double u[2] = {0.0, 0.0};
    if (live) {
      auto st2 = [&](double *v, double a, double b) { if (owner) reinterpret_cast<double2 *>(v)[e2] = make_double2(a, b); };
      const double a0[2] = {s0.x, s0.y}, a1[2] = {s1.x, s1.y}, lo[2] = {s2.x, s2.y}, up[2] = {s3.x, s3.y};
      if constexpr (MODE == RM_M) {
        // x~ rows: into this CTA's operand vector and every peer's (all-gather riding along the pass)
        if (owner) {
          if (EXT && dscale) { sum[0] *= s0.x; sum[1] *= s0.y; }
          reinterpret_cast<double2 *>(mdst)[e2] = make_double2(sum[0], sum[1]);
          if constexpr (CS > 1) {
            const uint32_t off = smem_u32(mdst) + 16u * (uint32_t)e2;
#pragma unroll
            for (int p = 0; p < CS; p++)
              if (p != X.rank) st_async_remote_v2(mapa(off, (uint32_t)p), sum[0], sum[1], mapa(mbar, (uint32_t)p));
          }
        }
      } else if constexpr (MODE == RM_A_ITER) {
        double zn[2], yn[2], dy[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
          const double zr = alpha * sum[i] + oma * a0[i];
          double z = zr + rinv2[i] * a1[i];
          z = fmin(fmax(z, lo[i]), up[i]);
          dy[i] = rho2[i] * (zr - z); yn[i] = a1[i] + dy[i]; zn[i] = z;
          u[i] = fma(rho2[i], z, -yn[i]);
        }
        st2(W.gz, zn[0], zn[1]); st2(W.gy, yn[0], yn[1]);
        if (do_check) st2(W.gdy, dy[0], dy[1]);
      } else if constexpr (MODE == RM_A_INIT) {
        st2(W.gz, sum[0], sum[1]);
        u[0] = fma(rho2[0], sum[0], -a1[0]); u[1] = fma(rho2[1], sum[1], -a1[1]);
      } else if constexpr (MODE == RM_A_RESUME) {
        u[0] = fma(rho2[0], a0[0], -a1[0]); u[1] = fma(rho2[1], a0[1], -a1[1]);
      } else if constexpr (MODE == RM_CHK_A1) {
        st2(W.gax, sum[0], sum[1]);
        u[0] = a1[0]; u[1] = a1[1];
      } else if constexpr (MODE == RM_CHK_A2) {
        st2(W.gadx, sum[0], sum[1]);
#pragma unroll
        for (int i = 0; i < 2; i++) {   // dy projected on the recession directions of [l, u] (OSQP is_primal_infeasible)
          double d = a0[i];
          if (up[i] > kInfty * kMinScaling) {
            if (lo[i] < -kInfty * kMinScaling) d = 0.0; else d = fmin(d, 0.0);
          } else if (lo[i] < -kInfty * kMinScaling) d = fmax(d, 0.0);
          u[i] = d;
        }
      } else if constexpr (MODE == RM_CHK_P) {
        st2(W.gpx, sum[0], sum[1]);
        u[0] = a0[0]; u[1] = a0[1];
      } else if constexpr (MODE == RM_OBJ_P) {
        st2(W.gpx, sum[0], sum[1]);
      }
    }
    RSTAMP(X, 3);
    if constexpr (kPass2) {
      // w from the C-fragment layout (row lane >> 2, nodes 2 (lane & 3) + e) to the B fragments of the two k halves
      const double v00 = __shfl_sync(0xffffffffu, u[0], src_b), v01 = __shfl_sync(0xffffffffu, u[1], src_b);
      const double v10 = __shfl_sync(0xffffffffu, u[0], src_b + 16), v11 = __shfl_sync(0xffffffffu, u[1], src_b + 16);
      const bool odd = (lane >> 2) & 1;
      const double bu0 = odd ? v01 : v00, bu1 = odd ? v11 : v10;
      const double *sp = slotp + a2;
      if (X.ntl == kTPW) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const double bu = h ? bu1 : bu0;
#pragma unroll
          for (int tl = 0; tl < kTPW; tl++) {
            double a[4];
#pragma unroll
            for (int mt = 0; mt < 4; mt++) a[mt] = sp[tl * kTileD + mt * 64 + 16 * h];
#pragma unroll
            for (int mt = 0; mt < 4; mt++) dmma(acc[tl][mt], a[mt], bu);
          }
        }
      } else {
#pragma unroll
        for (int tl = 0; tl < kTPW; tl++) {
          if (tl < X.ntl) {
#pragma unroll
            for (int mt = 0; mt < 4; mt++) dmma(acc[tl][mt], sp[tl * kTileD + mt * 64], bu0);
          }
        }
#pragma unroll
        for (int tl = 0; tl < kTPW; tl++) {
          if (tl < X.ntl) {
#pragma unroll
            for (int mt = 0; mt < 4; mt++) dmma(acc[tl][mt], sp[tl * kTileD + mt * 64 + 16], bu1);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(X.empty + 8u * slot);
    }
    RSTAMP(X, 4);
#ifdef BQP_ROWS_DEBUG
    X.npan[MODE == RM_M ? 0 : 1]++;
#endif
    X.gord++;
  }
  X.cnt += npc;
}

// 8 bytes into a peer CTA's shared memory, completing 8 transaction bytes on the peer's mbarrier
__device__ __forceinline__ void st_async_remote_f64(uint32_t raddr, double v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(raddr), "d"(v), "r"(rbar) : "memory");
}

// ---- column-space reduction of the pass-2 accumulators on the iteration fast path:
// groups > 0 stage their accumulators in colb, group 0 adds them and sends every column chunk to the CTA that finalises it
// (reduce-scatter into `recv`); the chunk owner adds the contributions in rank order; then ALL consumer threads of the
// owner form b' = sigma x - q + sum for the chunk (thread <-> element mapping of xr: element e0 + tid + 256 i, the x
// iterate lives in registers) and write it into every CTA's colb (all-gather).  Returns with colb complete and visible.
template <int CS>
__device__ __forceinline__ void reduce_b(Ctx<CS> &X, double (&acc)[kTPW][4][2], const double (&xr)[kXR]) {
  const DevInstance &I = X.S->I;
  const int lane = X.lane, gq = lane >> 2, tq = lane & 3;
  double2 *cb2 = reinterpret_cast<double2 *>(X.colb);
  if constexpr (CS > 1) {
    if (X.warp == 0 && lane == 0) {
      mbar_expect_tx(X.rbar, (uint32_t)((CS - 1) * X.chunk_nt * kTileD * 8));
      mbar_expect_tx(X.cbar, (uint32_t)((X.NW - X.chunk_nt) * kTileD * 8));
    }
  }
  if (kG > 1) {
    for (int g = 1; g < kG; g++) {
      if (X.grp == g) {
#pragma unroll
        for (int tl = 0; tl < kTPW; tl++)
          if (tl < X.ntl)
#pragma unroll
            for (int mt = 0; mt < 4; mt++) cb2[(32 * (X.tile0 + tl) + 8 * mt + gq) * (T8 / 2) + tq] = make_double2(acc[tl][mt][0], acc[tl][mt][1]);
      }
      named_bar(kG + 1, kConsThreads);
      if (X.grp == 0) {
#pragma unroll
        for (int tl = 0; tl < kTPW; tl++)
          if (tl < X.ntl)
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
              const double2 v = cb2[(32 * (X.tile0 + tl) + 8 * mt + gq) * (T8 / 2) + tq];
              acc[tl][mt][0] += v.x; acc[tl][mt][1] += v.y;
            }
      }
      if (g + 1 < kG) named_bar(kG + 1, kConsThreads);
    }
  }
  if (X.grp == 0) {
    if constexpr (CS > 1) {
      // contributions for chunks finalised elsewhere
#pragma unroll
      for (int tl = 0; tl < kTPW; tl++) {
        if (tl < X.ntl) {
          const int tile = X.tile0 + tl;
          const int owner = tile / X.chunk_nt;
          if (owner != X.rank) {
            const int src_slot = X.rank < owner ? X.rank : X.rank - 1;
            const uint32_t base = smem_u32(X.recv) + (uint32_t)(((src_slot * X.chunk_nt + (tile - owner * X.chunk_nt)) * kTileD) * 8);
            const uint32_t rb = mapa(X.rbar, (uint32_t)owner);
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
              st_async_remote_v2(mapa(base + (uint32_t)(((8 * mt + gq) * T8 + 2 * tq) * 8), (uint32_t)owner), acc[tl][mt][0], acc[tl][mt][1], rb);
          }
        }
      }
      bool mine = false;
#pragma unroll
      for (int tl = 0; tl < kTPW; tl++)
        if (tl < X.ntl && (X.tile0 + tl) / X.chunk_nt == X.rank) mine = true;
      if (mine) {
        mbar_wait(X.rbar, X.rph);
        const double2 *rv = reinterpret_cast<const double2 *>(X.recv);
#pragma unroll
        for (int tl = 0; tl < kTPW; tl++) {
          const int tile = X.tile0 + tl;
          if (tl < X.ntl && tile / X.chunk_nt == X.rank) {
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
              double t0 = 0.0, t1 = 0.0;
#pragma unroll
              for (int src = 0; src < CS; src++) {     // fixed rank order, whoever does the adding
                if (src == X.rank) { t0 += acc[tl][mt][0]; t1 += acc[tl][mt][1]; }
                else {
                  const int ss = src < X.rank ? src : src - 1;
                  const double2 v = rv[((ss * X.chunk_nt + (tile - X.rank * X.chunk_nt)) * kTileD + (8 * mt + gq) * T8) / 2 + tq];
                  t0 += v.x; t1 += v.y;
                }
              }
              cb2[(32 * tile + 8 * mt + gq) * (T8 / 2) + tq] = make_double2(t0, t1);
            }
          }
        }
      }
    } else {
#pragma unroll
      for (int tl = 0; tl < kTPW; tl++)
        if (tl < X.ntl)
#pragma unroll
          for (int mt = 0; mt < 4; mt++) cb2[(32 * (X.tile0 + tl) + 8 * mt + gq) * (T8 / 2) + tq] = make_double2(acc[tl][mt][0], acc[tl][mt][1]);
    }
  }
  named_bar(kG + 1, kConsThreads);                       // the sums of this CTA's chunk are in colb
  {
    const int tid = X.warp * 32 + lane;
    const int e0 = 32 * X.chunk_t0 * T8 + tid;
    const double sigma = I.sigma;
#pragma unroll
    for (int i = 0; i < kXR; i++) {
      if (i < X.chunk_nt) {
        const int e = e0 + i * kConsThreads, j = e >> 3;
        double b = 0.0;
        if (j < X.n) b = sigma * xr[i] - __ldg(I.q + j) + X.colb[e];
        X.colb[e] = b;
        if constexpr (CS > 1) {
          const uint32_t off = smem_u32(X.colb) + 8u * (uint32_t)e;
#pragma unroll
          for (int p = 0; p < CS; p++)
            if (p != X.rank) st_async_remote_f64(mapa(off, (uint32_t)p), b, mapa(X.cbar, (uint32_t)p));
        }
      }
    }
  }
  if constexpr (CS > 1) { mbar_wait(X.cbar, X.cph); X.rph ^= 1u; X.cph ^= 1u; }
  named_bar(kG + 1, kConsThreads);
}

// pass-2 accumulators of a check pass -> the tile's workspace (slot of this CTA and group); summed in the norms phase
template <int CS>
__device__ __forceinline__ void store_parts(Ctx<CS> &X, int which, const double (&acc)[kTPW][4][2]) {
  const int lane = X.lane, gq = lane >> 2, tq = lane & 3;
  double2 *dst = reinterpret_cast<double2 *>(X.Wk.parts + ((size_t)which * CS * kG + (size_t)X.rank * kG + (kG == 2 ? X.pset : X.grp)) * X.np * T8);
#pragma unroll
  for (int tl = 0; tl < kTPW; tl++)
    if (tl < X.ntl)
#pragma unroll
      for (int mt = 0; mt < 4; mt++) dst[(32 * (X.tile0 + tl) + 8 * mt + gq) * (T8 / 2) + tq] = make_double2(acc[tl][mt][0], acc[tl][mt][1]);
}

template <int CS, bool EXT>
__global__ void __launch_bounds__(kRowsThreads, 1)
admm_rows_kernel(const DevInstance *__restrict__ insts, const DevTile *__restrict__ tiles, const double *__restrict__ in,
                 double *__restrict__ out, double *__restrict__ work, NodeScalars *__restrict__ ns,
                 int *__restrict__ tile_iters, int nslots, double *__restrict__ state, int prefetch_panels,
                 const double *__restrict__ corr) {
  constexpr int T = T8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = CS > 1 ? (int)cluster_ctarank() : 0;
  const int tile_id = blockIdx.x / CS;
  RowsShared &S = *reinterpret_cast<RowsShared *>(smem_raw);
  if (tid == 0) {
    S.tile = tiles[tile_id];
    S.I = insts[S.tile.inst];
    S.remaining = S.tile.nn;
  }
  if (tid < T) { S.status[tid] = BQP_UNSOLVED; S.iters[tid] = 0; S.newly[tid] = 0; S.dist[tid] = NAN; }
  __syncthreads();
  if (tid < T) {
    // every leaf starts from the setup rho; a resumed round continues with the rho it had adapted to
    double r0 = S.I.rho_base;
    if (EXT && S.I.adaptive && tid < S.tile.nn && S.tile.iter_begin > 0) r0 = state[S.tile.state_off[tid] + S.I.n + 2 * (size_t)S.I.m];
    S.rho_t[tid] = r0; S.rinv_t[tid] = 1.0 / r0; S.rinveq_t[tid] = 1.0 / (kRhoEqFactor * r0);
  }
  if (tid == 0) S.rho_changed = 0;
  __syncthreads();
  const DevInstance &I = S.I;
  const int n = I.n, m = I.m, np = I.npad, nn = S.tile.nn, NW = I.p_nw;
  const int npm = I.p_npm, npa = I.p_npa;
  const int iter_begin = S.tile.iter_begin, iter_end = S.tile.iter_end;
  const int m8 = (m + 7) / 8 * 8;
  const bool is_producer = warp >= kCons;               // the producer WARPGROUP: warp kCons lane 0 issues the TMA copies, the other
                                                        // three warps only keep the CTA- and cluster-wide barriers company

  Ctx<CS> X;
  X.S = &S;
  size_t off = (sizeof(RowsShared) + 15) & ~size_t(15);
  X.full = smem_u32(smem_raw + off);
  X.empty = X.full + 8u * nslots; X.xbar = X.empty + 8u * nslots; X.rbar = X.xbar + 8u; X.cbar = X.rbar + 8u; X.vbar = X.cbar + 8u;
  off += sizeof(uint64_t) * (2 * (size_t)nslots + 4);
  off = (off + 127) & ~size_t(127);
  X.colb = reinterpret_cast<double *>(smem_raw + off); off += (size_t)np * T * 8;
  X.colx = reinterpret_cast<double *>(smem_raw + off); off += (size_t)np * T * 8;
  X.part = reinterpret_cast<double *>(smem_raw + off); off += (size_t)kG * 2 * kW * 32 * 16;
  X.recv = reinterpret_cast<double *>(smem_raw + off);
  const int chunk_nt = NW / CS;                        // the host launches CS > 1 only when CS divides NW
  off += CS > 1 ? (size_t)(CS - 1) * chunk_nt * kTileD * 8 : 0;
  off = (off + 127) & ~size_t(127);
  X.ring = smem_raw + off;
  X.ring_u32 = smem_u32(X.ring);
  X.nslots = nslots; X.slot_bytes = NW * kTileD * 8;
  X.NW = NW; X.np = np; X.n = n; X.m = m; X.npm = npm; X.npa = npa; X.rank = rank;
  X.lane = lane; X.warp = warp; X.grp = warp / kW; X.wi = warp % kW;
  const int tpw = (NW + kW - 1) / kW;
  X.tile0 = X.wi * tpw; X.ntl = max(0, min(tpw, NW - X.tile0));
  X.chunk_t0 = rank * chunk_nt; X.chunk_nt = chunk_nt;
  X.cnt = 0; X.gord = 0; X.xph = X.rph = X.cph = X.vph = 0;
  X.cmine = X.grp; X.slot = X.grp % nslots; X.phase = 0;
  {
    double *p = work + S.tile.work_off;
    Work &W = X.Wk;
    W.gz = p; p += (size_t)m8 * T; W.gy = p; p += (size_t)m8 * T; W.gl = p; p += (size_t)m8 * T; W.gu = p; p += (size_t)m8 * T;
    W.gdy = p; p += (size_t)m8 * T; W.gax = p; p += (size_t)m8 * T; W.gadx = p; p += (size_t)m8 * T;
    W.gx = p; p += (size_t)np * T; W.gdx = p; p += (size_t)np * T; W.gpx = p; p += (size_t)np * T; W.gxo = p; p += (size_t)np * T;
    W.gsx = p; p += (size_t)np * T; W.parts = p;
  }
  const Work &W = X.Wk;
  if (tid == 0) {
    for (int s = 0; s < nslots; s++) { mbar_init(X.full + 8u * s, 1); mbar_init(X.empty + 8u * s, kW); }
    mbar_init(X.xbar, 1); mbar_init(X.rbar, 1); mbar_init(X.cbar, 1); mbar_init(X.vbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  const int max_iter = I.max_iter, check_every = I.check_every;
  const int gtid = rank * kConsThreads + tid, gthreads = CS * kConsThreads;   // consumer threads of the whole cluster

  // ---- prologue (node.py:102-105): scaled bounds, warm start (or the saved state of a resumed round); split over the cluster
  if (!is_producer) {
    for (int e = gtid; e < m8 * T; e += gthreads) {
      const int i = e / T, t = e - i * T;
      double lo = -kInfty, up = kInfty, yv = 0.0, zv = 0.0, ei = 1.0;
      if (i < m) {
        if (t < nn) {
          const double *p = in + S.tile.in_off[t];
          lo = fmax(p[i], -kInfty);
          up = fmin(p[m + i], kInfty);
          if (iter_begin == 0) yv = I.c * __ldg(I.Einv + i) * p[2 * (size_t)m + n + i];
          else { const double *sp = state + S.tile.state_off[t] + n; zv = sp[i]; yv = sp[m + i]; }
        }
        ei = __ldg(I.E + i);
      }
      W.gl[e] = ei * lo; W.gu[e] = ei * up; W.gy[e] = yv; W.gz[e] = zv;
    }
    for (int e = gtid; e < np * T; e += gthreads) {
      const int j = e / T, t = e - j * T;
      double xv = 0.0;
      if (j < n && t < nn)
        xv = iter_begin == 0 ? __ldg(I.Dinv + j) * in[S.tile.in_off[t] + 2 * (size_t)m + j] : state[S.tile.state_off[t] + j];
      W.gx[e] = xv;
    }
  }
  cluster_sync_all<CS>();     // barriers of every CTA initialised, prologue vectors visible to the cluster

  const double *pM = I.pstream, *pA = I.pstream + I.p_offA, *pP = I.pstream + I.p_offP;

  // =============================================================== producer warpgroup: mirror of the pass sequence
  if (is_producer) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProdRegs));
    int pslot = 0; uint32_t pphase = 0;
    const uint32_t slot_bytes = (uint32_t)X.slot_bytes;
    auto produce = [&](const double *src, int npanels) {
      const int npc = (npanels - rank + CS - 1) / CS;
      if (warp == kCons && lane == 0) {
        for (int jl = 0; jl < npc; jl++) {
          const int slot = pslot;
          const uint32_t phase = pphase;
          if (++pslot == nslots) { pslot = 0; pphase ^= 1u; }
          mbar_wait(X.empty + 8u * slot, phase ^ 1u);       // passes at once on the first lap
          mbar_expect_tx(X.full + 8u * slot, slot_bytes);
          tma_load_1d(X.ring_u32 + (uint32_t)slot * slot_bytes, src + (size_t)(rank + jl * CS) * I.p_panel_doubles, slot_bytes, X.full + 8u * slot);
          const int jp = jl + prefetch_panels;
          if (prefetch_panels > 0 && jp < npc) l2_prefetch(src + (size_t)(rank + jp * CS) * I.p_panel_doubles, slot_bytes);
        }
      }
      __syncwarp();
    };
    produce(pA, npa);
    for (int iter = iter_begin + 1; iter <= iter_end; iter++) {
      const bool do_check = (iter % check_every == 0) || iter == max_iter;
      produce(pM, npm);
      if (EXT && I.adaptive) produce(I.pstream + I.p_offV, npm);   // x~ = V (d . (V' b)): the M slot holds V'
      produce(pA, npa);
      if (!do_check) continue;
      cluster_sync_all<CS>();                            // iterates of the check in global memory
      produce(pA, npa); produce(pA, npa); produce(pP, npm);
      cluster_sync_all<CS>();                            // raw products in global memory
      named_bar(kG + 2, kRowsThreads);                   // decision published
      if (S.remaining == 0 || iter == iter_end) break;
      if (EXT && S.rho_changed) produce(pA, npa);               // a leaf adapted its rho: the right-hand side is rebuilt
    }
    cluster_sync_all<CS>();                              // final iterates of every CTA's rows in global memory
    cluster_sync_all<CS>();                              // epilogue operand ready
    produce(pP, npm);
    cluster_sync_all<CS>();
    cluster_sync_all<CS>();                              // nobody leaves while a peer may still write into its shared memory
    return;
  }

  // =============================================================== consumer warps
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kConsRegs));
  double bx[kTPW][8];
  double acc[kTPW][4][2];

  // canonical reduction of per-thread values over the 256 consumer threads: lane tree, then the warps in order.
  // v: this thread's value for node (tid & 7); returns the total in every thread with the same node
  double *red = X.part;     // the group partial buffers are idle whenever this runs (kCons * 8 doubles per quantity)
  auto block_reduce = [&](double (&v)[kFinN], const int (&op)[kFinN], int nq, int q0 = 0) {
#pragma unroll
    for (int q = 0; q < kFinN; q++) {
      if (q < nq) {
        double x = v[q];
        for (int o = 16; o >= 8; o >>= 1) {
          const double y = __shfl_xor_sync(0xffffffffu, x, o);
          x = op[q] == 0 ? fmax(x, y) : (op[q] == 1 ? x + y : fmin(x, y));
        }
        v[q] = x;
      }
    }
    named_bar(kG + 1, kConsThreads);                     // previous users of `red` are done
    if (lane < 8)
      for (int q = 0; q < nq; q++) red[(q * kCons + warp) * 8 + lane] = v[q];
    named_bar(kG + 1, kConsThreads);
    if (tid < nq * 8) {
      const int q = tid >> 3, t = tid & 7;
      double x = red[(q * kCons) * 8 + t];
      for (int w = 1; w < kCons; w++) {
        const double y = red[(q * kCons + w) * 8 + t];
        x = op[q] == 0 ? fmax(x, y) : (op[q] == 1 ? x + y : fmin(x, y));
      }
      S.fin[q0 + q][t] = x;
    }
    named_bar(kG + 1, kConsThreads);
  };

  // scalar decision (optimality / infeasibility tests of OSQP) for node t at iteration `iter`
  auto decide = [&](int t, int iter) {
    S.newly[t] = 0;
    if (!(t < nn && S.status[t] == BQP_UNSOLVED)) return;
    const double cinv = I.cinv, c = I.c;
    const double pri = S.fin[5][t], dua = cinv * S.fin[0][t];
    const double nAx = S.fin[6][t], nz = S.fin[7][t], nPx = cinv * S.fin[1][t], nAty = cinv * S.fin[2][t], nq = cinv * I.nq;
    const double obj = (0.5 * S.fin[3][t] + S.fin[4][t]) * cinv;
    int status = BQP_UNSOLVED;
    const int passes = (iter == max_iter) ? 2 : 1;
    for (int pass = 0; pass < passes && status == BQP_UNSOLVED; pass++) {
      const double k = pass ? 10.0 : 1.0;
      const double eps_abs = I.eps_abs * k, eps_rel = I.eps_rel * k, eps_pinf = I.eps_pinf * k, eps_dinf = I.eps_dinf * k;
      if (pri > kInfty || dua > kInfty) { status = BQP_NON_CVX; break; }
      bool prim_ok = false, dual_ok = false, pinf = false, dinf = false;
      if (m == 0) prim_ok = true;
      else {
        const double eps_prim = eps_abs + eps_rel * fmax(nAx, nz);
        if (pri < eps_prim) prim_ok = true;
        else {
          const double nrm = S.fin[8][t];
          if (nrm > 1.0 / kInfty && S.fin[9][t] < -eps_pinf * nrm) pinf = S.fin[12][t] < eps_pinf * nrm;
        }
      }
      const double eps_dual = eps_abs + eps_rel * fmax(fmax(nPx, nAty), nq);
      if (dua < eps_dual) dual_ok = true;
      else {
        const double nrm = S.fin[10][t];
        if (nrm > 1.0 / kInfty && S.fin[11][t] < -c * eps_dinf * nrm && S.fin[13][t] < c * eps_dinf * nrm)
          dinf = !(S.fin[14][t] > eps_dinf * nrm) && !(S.fin[15][t] < -eps_dinf * nrm);
      }
      if (prim_ok && dual_ok) status = pass ? BQP_SOLVED_INACCURATE : BQP_SOLVED;
      else if (pinf) status = pass ? BQP_PRIMAL_INFEASIBLE_INACCURATE : BQP_PRIMAL_INFEASIBLE;
      else if (dinf) status = pass ? BQP_DUAL_INFEASIBLE_INACCURATE : BQP_DUAL_INFEASIBLE;
    }
    if (status == BQP_UNSOLVED && iter == max_iter) status = BQP_MAX_ITER_REACHED;
    if (status == BQP_UNSOLVED) {   // still running: distance to the (first-pass) tolerances, a scheduling hint for the host
      const double eps_prim = I.eps_abs + I.eps_rel * fmax(nAx, nz), eps_dual = I.eps_abs + I.eps_rel * fmax(fmax(nPx, nAty), nq);
      S.dist[t] = fmax(pri / eps_prim, dua / eps_dual);
    }
    if (status != BQP_UNSOLVED) {
      S.status[t] = status; S.iters[t] = iter; S.newly[t] = 1;
      NodeScalars r;
      r.status = status; r.iters = iter; r.pri_res = pri; r.dua_res = dua;
      r.obj = (status == BQP_PRIMAL_INFEASIBLE || status == BQP_PRIMAL_INFEASIBLE_INACCURATE) ? kInfty
              : (status == BQP_DUAL_INFEASIBLE || status == BQP_DUAL_INFEASIBLE_INACCURATE) ? -kInfty
              : (status == BQP_NON_CVX ? NAN : obj);
      r.lower = NAN;
      if (rank == 0) ns[S.tile.node[t]] = r;
      atomicSub(&S.remaining, 1);
    }
  };
  // unscaled iterates of the nodes that terminated at this check: the caller's buffers and the epilogue's copy (CTA 0 writes)
  auto snapshot = [&]() {
    if (rank != 0) return;
    for (int t = 0; t < nn; t++) {
      if (!S.newly[t]) continue;
      const int st = S.status[t];
      const bool bad = !(st == BQP_SOLVED || st == BQP_SOLVED_INACCURATE || st == BQP_MAX_ITER_REACHED);
      double *ox = out + S.tile.out_off[t], *oy = ox + n, *sx = W.gsx + (size_t)t * np;
      for (int j = tid; j < n; j += kConsThreads) {
        const double v = bad ? NAN : __ldg(I.D + j) * __ldcg(W.gx + (size_t)j * T + t);
        sx[j] = v; ox[j] = v;
      }
      for (int i = tid; i < m; i += kConsThreads) oy[i] = bad ? NAN : I.cinv * __ldg(I.E + i) * __ldcg(W.gy + (size_t)i * T + t);
    }
  };

  auto cons_bar = [&]() { named_bar(kG + 1, kConsThreads); };
  // start of a reduction / M pass: post the expected remote bytes of the phase
  auto post_x = [&](uint32_t bar) {
    if constexpr (CS > 1) {
      if (warp == 0 && lane == 0) {
        const int own = (npm - rank + CS - 1) / CS;
        mbar_expect_tx(bar, (uint32_t)((npm - own) * 8 * T * 8));
      }
    }
  };

  // the x iterate of the columns this CTA finalises, in registers: element 256 (8 chunk_t0 + i) + tid of the [column][8] layout
  double xr[kXR];
  const int xe0 = 32 * X.chunk_t0 * T + tid;
#pragma unroll
  for (int i = 0; i < kXR; i++) xr[i] = i < X.chunk_nt ? __ldcg(W.gx + xe0 + i * kConsThreads) : 0.0;

  // ---- first pass of the launch: z = A x0 (fresh nodes) and b' of the starting point
  load_bx<true>(bx, W.gx, X.tile0, X.ntl, lane);
  if (iter_begin == 0) rows_pass<CS, RM_A_INIT, EXT>(X, npa, bx, acc, false); else rows_pass<CS, RM_A_RESUME, EXT>(X, npa, bx, acc, false);
  reduce_b<CS>(X, acc, xr);

#ifdef BQP_ROWS_DEBUG
  long long ph[6] = {0, 0, 0, 0, 0, 0}, pl = clock64();
#define PSTAMP(i) do { const long long now_ = clock64(); ph[i] += now_ - pl; pl = now_; } while (0)
#else
#define PSTAMP(i) do { } while (0)
#endif
  const bool adaptive = EXT && I.adaptive != 0;
  double *const xt = adaptive ? X.colb : X.colx;            // where x~ of the iteration lands
  int iter;
  for (iter = iter_begin + 1; iter <= iter_end; iter++) {
    const bool do_check = (iter % check_every == 0) || iter == max_iter;
    // x~ = M b
    post_x(X.xbar);
    load_bx<false>(bx, X.colb, X.tile0, X.ntl, lane);
    PSTAMP(0);
    rows_pass<CS, RM_M, EXT>(X, npm, bx, acc, do_check, X.colx, adaptive, X.xbar);
    PSTAMP(1);
    cons_bar();                                           // local x~ rows visible
    if constexpr (CS > 1) { mbar_wait(X.xbar, X.xph); X.xph ^= 1u; }
    if (adaptive) {
      // second pass of x~ = V (d . (V' b)): the scaled coefficients are complete in colx (every CTA's copy); the product rows go
      // to colb, whose b every warp of the cluster has long loaded into registers (a peer reaches this pass only after it has
      // received this CTA's last coefficient row).  Its rows complete on a barrier of their own: with more than two CTAs a fast
      // peer's product rows can arrive while this CTA still waits for a third CTA's coefficients
      post_x(X.vbar);
      load_bx<false>(bx, X.colx, X.tile0, X.ntl, lane);
      rows_pass<CS, RM_M, EXT>(X, npm, bx, acc, do_check, X.colb, false, X.vbar);
      cons_bar();
      if constexpr (CS > 1) { mbar_wait(X.vbar, X.vph); X.vph ^= 1u; }
    }
    PSTAMP(2);
    if (EXT && I.eq2) {
      // eq_rho == 2: Woodbury correction of the explicit inverse over the re-typed integer rows of every node,
      //   x~ <- x~ - M[:,S] G (x~[S])     (warp w <-> node w; every CTA of the cluster corrects its full copy identically)
      if (warp < nn && S.tile.corr_off[warp] >= 0) {
        const double *cb = corr + S.tile.corr_off[warp];
        const int nS = (int)cb[0];
        const double *ksd = cb + 1, *jsd = cb + 1 + nS, *G = cb + 1 + 2 * nS;
        double *sc = reinterpret_cast<double *>(X.part) + warp * 2 * kMaxS;       // c = x~[S], then t = G c (group partial buffers are idle)
        for (int a = lane; a < nS; a += 32) sc[a] = X.colx[(size_t)((int)jsd[a]) * T + warp];
        __syncwarp();
        for (int a = lane; a < nS; a += 32) {
          double t_a = 0.0;
          for (int c = 0; c < nS; c++) t_a = fma(__ldg(G + (size_t)c * nS + a), sc[c], t_a);      // G symmetric: column-wise, coalesced
          sc[kMaxS + a] = t_a;
        }
        __syncwarp();
        for (int j = lane; j < np; j += 32) {
          double acc_j = 0.0;
          for (int a = 0; a < nS; a++) acc_j = fma(__ldg(I.p_mint + (size_t)((int)ksd[a]) * np + j), sc[kMaxS + a], acc_j);
          X.colx[(size_t)j * T + warp] -= acc_j;
        }
      }
      cons_bar();
    }
    // x = alpha x~ + (1 - alpha) x_prev for the columns this CTA finalises; then z~ = A x~, update, b'
    load_bx<false>(bx, xt, X.tile0, X.ntl, lane);
    {
      const double alpha = I.alpha, oma = 1.0 - I.alpha;
      const bool publish = do_check || iter == iter_end;     // the check passes and the epilogue read x from global memory
#pragma unroll
      for (int i = 0; i < kXR; i++) {
        if (i < X.chunk_nt) {
          const int e = xe0 + i * kConsThreads;
          const double xp = xr[i], xn = alpha * xt[e] + oma * xp;
          xr[i] = xn;
          if (publish) { W.gx[e] = xn; W.gdx[e] = xn - xp; }
        }
      }
    }
    cons_bar();
    PSTAMP(3);
    rows_pass<CS, RM_A_ITER, EXT>(X, npa, bx, acc, do_check);
    PSTAMP(4);
    reduce_b<CS>(X, acc, xr);
    PSTAMP(5);
    if (!do_check) continue;

    // ---- termination check (update_info + check_termination): A x | A' y,  A dx | A' dy_proj,  P x | P dx
    cluster_sync_all<CS>();
    load_bx<true>(bx, W.gx, X.tile0, X.ntl, lane);
    rows_pass<CS, RM_CHK_A1, EXT>(X, npa, bx, acc, true); store_parts<CS>(X, 0, acc);
    load_bx<true>(bx, W.gdx, X.tile0, X.ntl, lane);
    rows_pass<CS, RM_CHK_A2, EXT>(X, npa, bx, acc, true); store_parts<CS>(X, 1, acc);
    load_bx<true>(bx, W.gx, X.tile0, X.ntl, lane);
    rows_pass<CS, RM_CHK_P, EXT>(X, npm, bx, acc, true); store_parts<CS>(X, 2, acc);
    cluster_sync_all<CS>();
    {
      // norms, every CTA redundantly, one canonical order: thread <-> (node tid & 7, rows/columns (tid >> 3) + 32 i)
      const int t = tid & 7, s = tid >> 3;
      double v[kFinN];
      const int op[kFinN] = {0, 0, 0, 1, 1, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 2};
      v[0] = v[1] = v[2] = v[5] = v[6] = v[7] = v[8] = v[10] = v[12] = v[13] = 0.0;
      v[3] = v[4] = v[9] = v[11] = 0.0; v[14] = -INFINITY; v[15] = INFINITY;
      const size_t pstride = (size_t)np * T;
      for (int j = s; j < n; j += kConsThreads / 8) {
        const size_t e = (size_t)j * T + t;
        double aty = 0.0, atd = 0.0, pdx = 0.0;
        for (int p = 0; p < CS * kG; p++) {
          aty += __ldcg(W.parts + (size_t)p * pstride + e);
          atd += __ldcg(W.parts + ((size_t)CS * kG + p) * pstride + e);
          pdx += __ldcg(W.parts + ((size_t)2 * CS * kG + p) * pstride + e);
        }
        const double px = __ldcg(W.gpx + e), xj = __ldcg(W.gx + e), dxj = __ldcg(W.gdx + e);
        const double di = __ldg(I.Dinv + j), dj = __ldg(I.D + j), qj = __ldg(I.q + j);
        v[0] = fmax(v[0], fabs(di * (px + qj + aty)));
        v[1] = fmax(v[1], fabs(di * px));
        v[2] = fmax(v[2], fabs(di * aty));
        v[3] += xj * px;
        v[4] += qj * xj;
        v[12] = fmax(v[12], fabs(di * atd));
        v[13] = fmax(v[13], fabs(di * pdx));
        v[10] = fmax(v[10], fabs(dj * dxj));
        v[11] += qj * dxj;
      }
      for (int i = s; i < m; i += kConsThreads / 8) {
        const size_t e = (size_t)i * T + t;
        const double ax = __ldcg(W.gax + e), adx = __ldcg(W.gadx + e), z = __ldcg(W.gz + e), dy = __ldcg(W.gdy + e),
                     lo = __ldcg(W.gl + e), up = __ldcg(W.gu + e);
        const double ei = __ldg(I.Einv + i), Ei = __ldg(I.E + i);
        v[5] = fmax(v[5], fabs(ei * (ax - z)));
        v[6] = fmax(v[6], fabs(ei * ax));
        v[7] = fmax(v[7], fabs(ei * z));
        const double w = ei * adx;
        if (up < kInfty * kMinScaling) v[14] = fmax(v[14], w);
        if (lo > -kInfty * kMinScaling) v[15] = fmin(v[15], w);
        double d = dy;
        if (up > kInfty * kMinScaling) {
          if (lo < -kInfty * kMinScaling) d = 0.0; else d = fmin(d, 0.0);
        } else if (lo < -kInfty * kMinScaling) d = fmax(d, 0.0);
        v[8] = fmax(v[8], fabs(Ei * d));
        v[9] += up * fmax(d, 0.0) + lo * fmin(d, 0.0);
      }
      block_reduce(v, op, kFinN);
      if (adaptive && iter % I.adapt_interval == 0) {
        // the scaled norms of osqp's compute_rho_estimate: |Ax - z|, |z|, |Ax|, |Px + q + A'y|, |q|, |A'y|, |Px| (same canonical order)
        const int op2[kFinN] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int q = 0; q < kFinN; q++) v[q] = 0.0;
        for (int j = s; j < n; j += kConsThreads / 8) {
          const size_t e = (size_t)j * T + t;
          double aty = 0.0;
          for (int p = 0; p < CS * kG; p++) aty += __ldcg(W.parts + (size_t)p * pstride + e);
          const double px = __ldcg(W.gpx + e), qj = __ldg(I.q + j);
          v[3] = fmax(v[3], fabs(px + qj + aty)); v[4] = fmax(v[4], fabs(qj)); v[5] = fmax(v[5], fabs(aty)); v[6] = fmax(v[6], fabs(px));
        }
        for (int i = s; i < m; i += kConsThreads / 8) {
          const size_t e = (size_t)i * T + t;
          const double ax = __ldcg(W.gax + e), z = __ldcg(W.gz + e);
          v[0] = fmax(v[0], fabs(ax - z)); v[1] = fmax(v[1], fabs(z)); v[2] = fmax(v[2], fabs(ax));
        }
        block_reduce(v, op2, 8, kFinN);
      }
    }
    if (tid < T) decide(tid, iter);
    cons_bar();
    snapshot();
    if (adaptive && iter % I.adapt_interval == 0) {
      // osqp adapt_rho: rho_new = rho sqrt(pri / dua) on the normalised scaled residuals, adopted outside [rho / tol, rho tol]
      if (tid < nn && S.status[tid] == BQP_UNSOLVED) {
        const int t = tid;
        const double pri = S.fin[kFinN + 0][t] / (fmax(S.fin[kFinN + 1][t], S.fin[kFinN + 2][t]) + 1e-10);
        const double dua = S.fin[kFinN + 3][t] / (fmax(fmax(S.fin[kFinN + 4][t], S.fin[kFinN + 5][t]), S.fin[kFinN + 6][t]) + 1e-10);
        const double rho = S.rho_t[t];
        double rn = rho * sqrt(pri / (dua + 1e-10));
        rn = fmin(fmax(rn, kRhoMin), 1e6);
        if (rn > rho * I.adapt_tol || rn < rho / I.adapt_tol) {
          S.rho_t[t] = rn; S.rinv_t[t] = 1.0 / rn; S.rinveq_t[t] = 1.0 / (kRhoEqFactor * rn);
          S.rho_changed = 1;
        }
      }
      cons_bar();
    }
    named_bar(kG + 2, kRowsThreads);                      // ... and the producer sees the decision
    if (S.remaining == 0 || iter == iter_end) break;
    if (EXT && S.rho_changed) {
      // the right-hand side of the next iteration carries rho: u = rho z - y and b' = sigma x - q + A' u again
      rows_pass<CS, RM_A_RESUME, EXT>(X, npa, bx, acc, false);
      reduce_b<CS>(X, acc, xr);
      if (tid == 0) S.rho_changed = 0;
      cons_bar();
    }
  }

  // ---- end of the launch: save the state of unfinished nodes, clip (node.py:128-143), objective at the clipped point
  cluster_sync_all<CS>();                                 // final iterates of every CTA's rows in global memory
  if (rank == 0) {
    if (tid == 0) tile_iters[tile_id] = (iter > iter_end ? iter_end : iter) - iter_begin;
    for (int t = 0; t < nn; t++) {
      if (S.status[t] != BQP_UNSOLVED) continue;
      double *sp = state + S.tile.state_off[t];
      for (int j = tid; j < n; j += kConsThreads) sp[j] = __ldcg(W.gx + (size_t)j * T + t);
      for (int i = tid; i < m; i += kConsThreads) { sp[n + i] = __ldcg(W.gz + (size_t)i * T + t); sp[n + m + i] = __ldcg(W.gy + (size_t)i * T + t); }
      if (EXT && tid == 0) sp[n + 2 * (size_t)m] = S.rho_t[t];
      if (tid == 0) {   // pri_res of a node that is still running carries its distance to the tolerance (scheduling hint)
        NodeScalars r; r.status = BQP_UNSOLVED; r.iters = iter_end; r.obj = r.dua_res = r.lower = NAN; r.pri_res = S.dist[t];
        ns[S.tile.node[t]] = r;
      }
    }
    for (int t = 0; t < nn; t++) {
      const int st = S.status[t];
      if (!(st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED)) continue;
      double *ox = out + S.tile.out_off[t], *sx = W.gsx + (size_t)t * np;
      const double *p = in + S.tile.in_off[t];
      for (int k = tid; k < I.n_int; k += kConsThreads) {
        const int j = __ldg(I.i_idx + k), row = m - I.n_int + k;
        const double v = fmin(fmax(sx[j], p[row]), p[m + row]);
        sx[j] = v; ox[j] = v;
      }
    }
    cons_bar();
    for (int e = tid; e < np * T; e += kConsThreads) {
      const int j = e / T, t = e - j * T;
      double v = 0.0;
      if (j < n && t < nn) {
        const int st = S.status[t];
        if (st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED) v = __ldg(I.Dinv + j) * W.gsx[(size_t)t * np + j];
      }
      W.gxo[e] = v;
    }
  }
  cluster_sync_all<CS>();
  load_bx<true>(bx, W.gxo, X.tile0, X.ntl, lane);
  rows_pass<CS, RM_OBJ_P, EXT>(X, npm, bx, acc, false);
  cluster_sync_all<CS>();
  {
    const int t = tid & 7, s = tid >> 3;
    double v[kFinN];
    const int op[kFinN] = {1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    v[0] = v[1] = 0.0;
    for (int j = s; j < n; j += kConsThreads / 8) {
      const size_t e = (size_t)j * T + t;
      const double xo = __ldcg(W.gxo + e);
      v[0] += xo * __ldcg(W.gpx + e);
      v[1] += __ldg(I.q + j) * xo;
    }
    block_reduce(v, op, 2);
    if (rank == 0 && tid < nn) {
      const int st = S.status[tid];
      if (st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED) ns[S.tile.node[tid]].lower = (0.5 * S.fin[0][tid] + S.fin[1][tid]) * I.cinv;
    }
  }
#ifdef BQP_ROWS_DEBUG
  if (lane == 0 && blockIdx.x < 2 && (warp == 0 || warp == 5))
    printf("PHASES blk %d warp %d per iteration: bxload %.0f Mpass %.0f xwait %.0f xupd %.0f Apass %.0f reduce %.0f\n", (int)blockIdx.x, warp,
           (double)ph[0] / (iter_end - iter_begin), (double)ph[1] / (iter_end - iter_begin), (double)ph[2] / (iter_end - iter_begin),
           (double)ph[3] / (iter_end - iter_begin), (double)ph[4] / (iter_end - iter_begin), (double)ph[5] / (iter_end - iter_begin));
  if (lane == 0 && blockIdx.x == 0 && (warp == 0 || warp == 5))
    for (int k = 0; k < 2; k++)
      printf("ROWS warp %d %s panels %d: full-wait %.0f p1 %.0f bar %.0f update %.0f p2 %.0f between %.0f\n", warp, k ? "A+other" : "M", X.npan[k],
             (double)X.tacc[k][0] / X.npan[k], (double)X.tacc[k][1] / X.npan[k], (double)X.tacc[k][2] / X.npan[k], (double)X.tacc[k][3] / X.npan[k],
             (double)X.tacc[k][4] / X.npan[k], (double)X.tacc[k][5] / X.npan[k]);
#endif
  cluster_sync_all<CS>();                                 // nobody leaves while a peer may still write into its shared memory
}

}  // namespace

size_t rows_smem_bytes(int npad, int nslots, int cs) {
  const int nw = npad / 32;
  size_t off = (sizeof(RowsShared) + 15) & ~size_t(15);
  off += sizeof(uint64_t) * (2 * (size_t)nslots + 4);
  off = (off + 127) & ~size_t(127);
  off += (size_t)2 * npad * T8 * 8 + (size_t)kG * 2 * kW * 32 * 16;
  off += cs > 1 ? (size_t)(cs - 1) * (nw / cs) * kTileD * 8 : 0;
  off = (off + 127) & ~size_t(127);
  return off + (size_t)nslots * nw * kTileD * 8;
}

template <int CS, bool EXT>
static int launch_r(int nslots, double *d_state, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles, const double *d_in,
                    double *d_out, double *d_work, NodeScalars *d_ns, int *d_tile_iters, size_t smem, const double *d_corr, cudaStream_t st) {
  // many host threads launch concurrently (one context each): raise the attribute only when it has to grow
  static std::atomic<size_t> smem_set{0};
  cudaError_t e = cudaSuccess;
  if (smem_set.load() < smem) {
    e = cudaFuncSetAttribute(admm_rows_kernel<CS, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    if (e != cudaSuccess) return BQP_E_CUDA;
    smem_set.store((size_t)kMaxSmem);
  }
  int prefetch_panels = 4;   // L2 prefetch distance of the producer, in panels of this CTA
  if (const char *pk = getenv("BQP_ROWS_PREFETCH")) prefetch_panels = atoi(pk);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ntiles * CS), 1, 1);
  cfg.blockDim = dim3((unsigned)kRowsThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, admm_rows_kernel<CS, EXT>, d_insts, d_tiles, d_in, d_out, d_work, d_ns, d_tile_iters, nslots, d_state,
                         prefetch_panels, d_corr);
  return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? BQP_OK : BQP_E_CUDA;
}

int launch_admm_rows(int cs, int ext, int nslots, double *d_state, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles,
                     const double *d_in, double *d_out, double *d_work, NodeScalars *d_ns, int *d_tile_iters, size_t smem_bytes,
                     const double *d_corr, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (nslots < 2) return BQP_E_ARG;
#define BQP_ROWS_LAUNCH(C)                                                                                                              \
  if (cs == C)                                                                                                                          \
    return ext ? launch_r<C, true>(nslots, d_state, d_insts, d_tiles, ntiles, d_in, d_out, d_work, d_ns, d_tile_iters, smem_bytes, d_corr, st) \
               : launch_r<C, false>(nslots, d_state, d_insts, d_tiles, ntiles, d_in, d_out, d_work, d_ns, d_tile_iters, smem_bytes, d_corr, st)
  BQP_ROWS_LAUNCH(1); BQP_ROWS_LAUNCH(2); BQP_ROWS_LAUNCH(4); BQP_ROWS_LAUNCH(8);
#undef BQP_ROWS_LAUNCH
  return BQP_E_ARG;
}

}  // namespace bqp
