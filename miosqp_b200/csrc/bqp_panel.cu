// bqp_panel.cu -- fused single-pass batched ADMM kernel for sm_100a (dense A, npad <= 512).
//
// Same node-tile ownership as the other kernels (a tile = up to 8 B&B leaves of one problem, whole OSQP loop in-kernel;
// /root/reference/miosqp/node.py:96-143), but the iteration is restated so that A is streamed from HBM ONCE per ADMM
// iteration instead of twice (A' then A) and the triangular sweeps disappear:
//
//     x~ = M b                         M = (P + sigma I + A' rho A)^-1, explicit, one dependency-free mat-vec
//     z~ = A x~ ; z,y update ; b' = sigma x - q + A'(rho z - y)     ONE pass over A
//
// Every matrix is cut into row PANELS (kPanelRows = 8 rows x npad columns, bqp_internal.h), streamed by TMA bulk copies
// into a ring of shared-memory slots.  While panel k of A sits in shared memory it is used twice:
//   pass 1   z~_I = A_I x~            a PASS-1 WARP owns one column tile (32 columns); the [8 rows x 32 cols] x [32 cols x
//                                     8 nodes] product is 8 FP64 mma.sync.m8n8k4 (the 8 leaves of the tile are the N
//                                     dimension, so the hardware does the reduction over columns: no shuffles); the 8x8
//                                     partial sum of every warp goes to the
//   update   z_I, y_I, w_I            UPDATE WARPS (one lane per (row, node pair)), which add the warp partials in a fixed
//                                     order, apply the projection / dual update and publish w_I = rho z_I - y_I;
//   pass 2   b' += A_I' w_I           PASS-2 WARPS (two column tiles each) follow the update warps panel by panel and read
//                                     the same slot transposed: 8 more mma.sync per tile, the [32 cols x 8 nodes]
//                                     accumulators stay in registers for the whole pass.
// The consumer warps are specialised by pass so that the warps sharing an SM sub-partition (and its FP64 mma pipe) are in
// different phases of the panel loop.  (FP64 mma.sync issues at the same 64 FMA/clk/SM as DFMA on B200 --
// tools/micro/dmma_rate.cu -- but with 1/8 of the instructions; the vector-FMA version of this kernel was
// instruction-issue bound at a quarter of the HBM roofline.)
//
// Problems wider than 8 column tiles run as a CLUSTER OF TWO CTAs (one SM each): CTA r streams and multiplies only
// its half of the columns of every panel (half the HBM stream, shared-memory traffic and FP64 work per SM), the per-warp
// partial sums of pass 1 are copied into the peer CTA's shared memory with one bulk DSMEM copy per warp and panel
// (cp.async.bulk shared::cta -> shared::cluster, completing transaction bytes on the peer's mbarrier: no cluster-scope
// fence anywhere in the loop -- a release.cluster arrive costs a MEMBAR.ALL.GPU), and both CTAs run the (cheap)
// row-space update redundantly on identical inputs, so nothing else crosses the pair.
// Roles per CTA: pass-1 warps, pass-2 warps, kPanelUpdWarps update warps, one TMA producer warp (one lane).
// Synchronisation inside a pass: mbarriers (full/empty per ring slot; "partials full", "update done", "u consumed" per
// hand-off buffer, each buffer owned by one update warp so that every barrier is waited on strictly phase by phase) and,
// for the pass-1 warps' flow control, polled progress counters in shared memory (a completed mbarrier try_wait costs
// ~200 cycles).  Named barriers only at termination checks and once per iteration between pass-2 and pass-1 warps.
// Every wait is bounded: a broken protocol traps instead of hanging the GPU.  -DBQP_PANEL_DEBUG adds per-role phase timers.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>

#include "bqp_internal.h"

namespace bqp {

namespace {

constexpr int T8 = kPanelT;       // nodes per tile = N of the mma
constexpr int kPR = kPanelRows;
constexpr int kColQ = 9;          // column-space reductions per termination check
constexpr int kFin = 16;
constexpr int kFinP = 9;          // row-space quantities each update warp accumulates
constexpr int kHBmul = 2;
constexpr int kHB = kHBmul * kPanelUpdWarps;   // hand-off buffers; buffer b = panel % kHB always belongs to update warp b % kPanelUpdWarps
constexpr int kPH = kPanelRows / 8;            // 8-row mma tiles per panel
constexpr int kP1T = kPanelP1Tiles;
constexpr int kSets = kPanelP1Sets;
constexpr int kPanelThreads = panel_cta_warps(kPanelCtaWarps) * 32;
static_assert(kPanelThreads <= 1024, "panel kernel CTA too large");
constexpr int kTileDoubles = kPR * 32;   // one column tile of one panel

// ------------------------------------------------------------------ mbarrier / TMA / cluster wrappers (PTX)
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
// Bounded wait: a broken protocol traps (reported as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
#ifdef BQP_PANEL_DEBUG
    if (clock64() - t0 > 400000000LL) {
      if ((threadIdx.x & 31) == 0) printf("TIMEOUT blk %d warp %d bar %u parity %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), bar, parity);
      __trap();
    }
#else
    if (clock64() - t0 > 20000000000LL) __trap();   // ~10 s at 2 GHz
#endif
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
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// DSMEM store that signals the peer's mbarrier when it has landed (complete_tx of 8 bytes): data and notification in one
// asynchronous operation, so the sender needs no cluster-scope fence
__device__ __forceinline__ void st_remote_u32(uint32_t raddr, uint32_t v) {
  asm volatile("st.relaxed.cluster.shared::cluster.u32 [%0], %1;" ::"r"(raddr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_async_remote_f64(uint32_t raddr, double v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(raddr), "d"(v), "r"(rbar) : "memory");
}
// bulk DSMEM copy: `bytes` of this CTA's shared memory into the peer's, completing transaction bytes on the peer's mbarrier.
// The source was written with ordinary stores: the writers fence the async proxy first (fence_async_smem).
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_remote, uint32_t src_local, uint32_t bytes, uint32_t rbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_remote),
               "r"(src_local), "r"(bytes), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// D (8x8) += A (8x4, row) * B (4x8, col), FP64.  Fragments: A: lane holds A[lane>>2][lane&3]; B: B[lane&3][lane>>2];
// C/D: rows lane>>2, columns 2*(lane&3), 2*(lane&3)+1.
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

struct PanelShared {
  DevInstance I;
  DevTile tile;
  double fin[kFin][T8];
  double finp[kPanelUpdWarps][kFinP][T8];
  int status[T8], iters[T8], newly[T8];
  double dist[T8];          // max(pri_res / eps_prim, dua_res / eps_dual) at the last check: how far a running node is from done
  int remaining;
  // panels finished by each update warp: [0] of this CTA, [1] of the peer CTA (written there over DSMEM).  The pass-1 warps
  // poll these with plain shared-memory loads for flow control (a completed mbarrier try_wait costs ~200 cycles)
  volatile int upd_cnt[2][kPanelUpdWarps];
};

// everything a role needs to find its way around shared memory (u32 = shared-window addresses; *_r = the same object in
// the peer CTA of the pair, mapped with mapa)
struct Lay {
  uint32_t full, empty, pf, ud, uc, ck;      // barrier arrays: [nslots], [nslots], [kHB] x3, [1]
  uint32_t pf_r, ck_r, part_r, red_r, part_u32, cnt_r;          // cnt_r: upd_cnt[1] of the peer
  volatile int *cnt;                                           // upd_cnt[0] of this CTA ([kPanelUpdWarps], then the peer's copy)
  double *xts, *vs;                      // x~ and b for THIS CTA's columns: [32 NWc][8]
  double *part, *ubuf, *red;             // [kHB][NW][32 lanes][2], [kHB][8 rows][8], [kColQ][NW][8]
  unsigned char *ring;
  uint32_t ring_u32;
  int nslots, slot_bytes, nw, nwc, w0, np;    // nw: column tiles of the problem; nwc, w0: this CTA's share
  int npt, np1;                               // pass-1 warps (= partial sums per panel) of the pair / of this CTA
};

// LAG = how many panels pass 2 runs behind pass 1 (the update latency it hides), LAG < kHB.  CS = CTAs per tile.
// Consumer warps are specialised by pass, so that the two warps sharing an SM sub-partition (and its FP64 mma pipe) are
// naturally in different phases of the panel loop:
//   P1 warps: pass 1 only.  Each owns up to two column tiles; its C fragment (8 rows x 8 nodes per mma tile) sums over
//             both, so a panel yields one partial per P1 warp.
//   P2 warps: pass 2 only, following the update warps panel by panel; each owns the accumulators of up to two column tiles.
// Both walk the same ring of slots (a slot is released when every P1 and every P2 warp has arrived on its empty barrier).
#ifdef BQP_PANEL_DEBUG
#define TSTAMP(obj, i) do { const long long now_ = clock64(); (obj).tacc[i] += now_ - (obj).tlast; (obj).tlast = now_; } while (0)
#else
#define TSTAMP(obj, i) do { } while (0)
#endif
template <int CS>
struct ConsumerBase {
  Lay L;
#ifdef BQP_PANEL_DEBUG
  long long tacc[6] = {0, 0, 0, 0, 0, 0}, tlast = 0;
#endif
  int lane, ntl, wg0;                // lane; column tiles of this warp (1 or 2) and the first of them (global index)
  int tl0;                           // first tile, index inside this CTA's slot
  int slot; uint32_t phase;          // ring position of the next panel
  int g, gb;                         // global panel counter (same sequence in the update warps), g % kHB
  int ud_g, ud_b; uint32_t ud_ph;    // next panel whose "update done" barrier this thread has not observed yet

  __device__ __forceinline__ void init(const Lay &lay, int lane_, int tile0_local, int ntiles) {
    L = lay; lane = lane_; tl0 = tile0_local; ntl = ntiles; wg0 = lay.w0 + tile0_local;
    slot = 0; phase = 0; g = 0; gb = 0; ud_g = 0; ud_b = 0; ud_ph = 0;
  }
  __device__ __forceinline__ void wait_ud(int target) {
    while (ud_g <= target) {
      mbar_wait(L.ud + 8u * ud_b, ud_ph);
      ud_g++;
      if (++ud_b == kHB) { ud_b = 0; ud_ph ^= 1u; }
    }
  }
  __device__ __forceinline__ void advance() {
    if (++slot == L.nslots) { slot = 0; phase ^= 1u; }
    g++;
    if (++gb == kHB) gb = 0;
  }
  // flow control by the update warps' progress counters: every panel < `upto` has been updated (in both CTAs of a pair)
  __device__ __forceinline__ void wait_progress(int upto) {
    if (upto <= 0) return;
    const long long t0 = clock64();
#pragma unroll
    for (int u = 0; u < kPanelUpdWarps; u++) {
      const int need = (upto - u + kPanelUpdWarps - 1) / kPanelUpdWarps;     // panels u, u + KU, ... below upto
      while (L.cnt[u] < need || (CS == 2 && L.cnt[kPanelUpdWarps + u] < need)) {
        if (clock64() - t0 > 20000000000LL) __trap();
      }
    }
  }
  // a pass a pass-2 warp has no work in: keep its place in the ring, in the panel sequence and in the barrier phases
  __device__ __forceinline__ void skip(int npanels) {
    for (int k = 0; k < npanels; k++) {
      wait_ud(g);
      mbar_wait(L.full + 8u * slot, phase);
      if (lane == 0) { mbar_arrive(L.empty + 8u * slot); mbar_arrive(L.uc + 8u * gb); }
      advance();
    }
  }
};

template <int CS>
struct P1Warp : ConsumerBase<CS> {
  using B = ConsumerBase<CS>;
  int pidx;                          // this warp's partial slot (index over the P1 warps of ONE set of both CTAs)
  int set;                           // this warp handles the panels with global counter g % kSets == set
  // pass 1 over `npanels` panels with the column vector src ([column - src_col0][8 nodes])
  __device__ __forceinline__ void pass(int npanels, const double *src, int src_col0) {
    const Lay &L = B::L;
    const int lane = B::lane, gq = lane >> 2, tq = lane & 3;
    double bx[2][8];
#pragma unroll
    for (int tl = 0; tl < 2; tl++)
#pragma unroll
      for (int ks = 0; ks < 8; ks++)
        bx[tl][ks] = (tl < kP1T && tl < B::ntl) ? src[(size_t)(32 * (B::wg0 + tl) + 4 * ks + tq - src_col0) * T8 + gq] : 0.0;
    for (int k = 0; k < npanels; k++) {
      if (kSets > 1 && (B::g % kSets) != set) { B::advance(); continue; }     // the other set's panel
      TSTAMP(*this, 5);
      mbar_wait(L.full + 8u * B::slot, B::phase);
      TSTAMP(*this, 0);
      const double *sp = reinterpret_cast<const double *>(L.ring + (size_t)B::slot * L.slot_bytes) + B::tl0 * kTileDoubles + lane;
      double c0[kPH][2];
#pragma unroll
      for (int h = 0; h < kPH; h++) {
        double cc[8][2];     // independent mma chains (two tiles deep), then a fixed tree
#pragma unroll
        for (int ks = 0; ks < 8; ks++) { cc[ks][0] = 0.0; cc[ks][1] = 0.0; dmma(cc[ks], sp[h * 256 + ks * 32], bx[0][ks]); }
        if (kP1T == 2 && B::ntl == 2) {
#pragma unroll
          for (int ks = 0; ks < 8; ks++) dmma(cc[ks], sp[kTileDoubles + h * 256 + ks * 32], bx[1][ks]);
        }
#pragma unroll
        for (int i = 0; i < 2; i++) c0[h][i] = ((cc[0][i] + cc[1][i]) + (cc[2][i] + cc[3][i])) + ((cc[4][i] + cc[5][i]) + (cc[6][i] + cc[7][i]));
      }
      TSTAMP(*this, 1);
      if (B::g >= kHB) {   // the update warps of both CTAs have consumed this partials buffer (panel g - kHB)
        const int pg = B::g - kHB, u = pg % kPanelUpdWarps, need = pg / kPanelUpdWarps + 1;
        const long long t0 = clock64();
        while (L.cnt[u] < need || (CS == 2 && L.cnt[kPanelUpdWarps + u] < need)) {
          if (clock64() - t0 > 20000000000LL) __trap();
        }
      }
      TSTAMP(*this, 2);
      const int pw0 = ((B::gb * L.npt + pidx) * kPH) * 64;     // this warp's partial block (doubles): kPH x 32 lanes x 2
#pragma unroll
      for (int h = 0; h < kPH; h++)
        *reinterpret_cast<double2 *>(L.part + pw0 + (h * 32 + lane) * 2) = make_double2(c0[h][0], c0[h][1]);
      if constexpr (CS == 2) fence_async_smem();    // the bulk copy below reads these stores through the async proxy
      __syncwarp();
      if (lane == 0) {
        // the same block into the peer CTA's buffer: one bulk DSMEM copy, completing bytes on the peer's barrier
        if constexpr (CS == 2) dsmem_bulk_copy(L.part_r + 8u * pw0, L.part_u32 + 8u * pw0, (uint32_t)(kPH * 32 * 16), L.pf_r + 8u * B::gb);
        // pair: the barrier also counts the bytes the peer's P1 warps store into our buffer (posted by P1 warp 0)
        if (CS == 2 && B::tl0 == 0) mbar_expect_tx(L.pf + 8u * B::gb, (uint32_t)((L.npt - L.np1) * kPH * 32 * 16));
        else mbar_arrive(L.pf + 8u * B::gb);
        mbar_arrive(L.empty + 8u * B::slot);
      }
      B::advance();
      TSTAMP(*this, 3);
    }
  }
};

template <int CS>
struct P2Warp : ConsumerBase<CS> {
  using B = ConsumerBase<CS>;
  // pass 2 over `npanels` panels: acc[tile][mt] (C fragments: column 32 wg + 8 mt + (lane >> 2), nodes 2 (lane & 3) + {0, 1})
  // += A_panel' u, panel by panel as the update warps publish u
  __device__ __forceinline__ void pass(int npanels, double (&acc)[2][4][2]) {
    const Lay &L = B::L;
    const int lane = B::lane, gq = lane >> 2, tq = lane & 3;
    // A fragment of A' (row operand): element (row 4 kk + tq, column 8 mt + gq) of the tile
    const int a2 = (gq >> 2) * 32 + tq * 4 + (gq & 3);
    double acc1[2][4][2];     // second accumulator set (rows 4..7 of every mma tile): independent chains
#pragma unroll
    for (int tl = 0; tl < 2; tl++)
#pragma unroll
      for (int mt = 0; mt < 4; mt++) { acc[tl][mt][0] = acc[tl][mt][1] = 0.0; acc1[tl][mt][0] = acc1[tl][mt][1] = 0.0; }
    for (int k = 0; k < npanels; k++) {
      TSTAMP(*this, 5);
      B::wait_ud(B::g);
      TSTAMP(*this, 0);
      mbar_wait(L.full + 8u * B::slot, B::phase);     // long complete: makes the TMA data visible to this warp
      TSTAMP(*this, 1);
      const double *up = L.ubuf + B::gb * (kPR * T8);
      const double *sp = reinterpret_cast<const double *>(L.ring + (size_t)B::slot * L.slot_bytes) + B::tl0 * kTileDoubles + a2;
#pragma unroll
      for (int h = 0; h < kPH; h++) {
        const double bu0 = up[(h * 8 + tq) * T8 + gq], bu1 = up[(h * 8 + 4 + tq) * T8 + gq];
#pragma unroll
        for (int tl = 0; tl < 2; tl++) {
          if (tl < B::ntl) {
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
              dmma(acc[tl][mt], sp[tl * kTileDoubles + h * 256 + mt * 64], bu0);
              dmma(acc1[tl][mt], sp[tl * kTileDoubles + h * 256 + mt * 64 + 16], bu1);
            }
          }
        }
      }
      __syncwarp();
      TSTAMP(*this, 2);
      if (lane == 0) { mbar_arrive(L.empty + 8u * B::slot); mbar_arrive(L.uc + 8u * B::gb); }   // slot and u buffer consumed
      B::advance();
      TSTAMP(*this, 3);
    }
#pragma unroll
    for (int tl = 0; tl < 2; tl++)
#pragma unroll
      for (int mt = 0; mt < 4; mt++) { acc[tl][mt][0] += acc1[tl][mt][0]; acc[tl][mt][1] += acc1[tl][mt][1]; }
  }
};

template <int OP>   // OP 0: max, 1: sum, 2: min
__device__ __forceinline__ double red_op(double v, double w) { return OP == 0 ? fmax(v, w) : (OP == 1 ? v + w : fmin(v, w)); }
// reduce over the 8 rows of a panel: lanes with the same lane & 3 (xor masks 16, 8, 4)
template <int OP>
__device__ __forceinline__ double reduce_rows(double v) {
#pragma unroll
  for (int o = 16; o >= 4; o >>= 1) v = red_op<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int OP>
__device__ __forceinline__ double reduce_warp(double v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = red_op<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

enum { PM_M = 0, PM_A_INIT, PM_A_RESUME, PM_A_ITER, PM_A_CHK1, PM_A_CHK2, PM_P_CHK, PM_P_OBJ };

struct WorkPtrs {   // per-CTA workspace in global memory (L2 resident), every vector [row][8 nodes]
  double *gz, *gy, *gl, *gu, *gdy, *gdx, *gpx, *gaty, *gatd, *gpdx, *gxs, *gsx;
};

// accumulators of an update warp lane (one row of every panel it handles, two nodes), reduced over rows at decision time
struct RowAcc {
  double pr[2], a1[2], a2[2], vu[2], vl[2], ndy[2], lhs[2], quad[2], lin[2];
};

// Update warps: warp uw handles the panels whose hand-off buffer is uw (global panel counter % kHB), so it waits on its
// "partials full" barrier strictly phase by phase.  Within one pass that is every kHB-th panel starting at some class
// c = k % kHB: `cls` (set by pass()) names it, so that sums over rows can be combined class by class -- a fixed order
// whatever ran earlier in the launch.  With CS == 2 both CTAs of the pair run this redundantly on identical inputs.
// Lane <-> (row lane >> 2 of the panel, nodes 2 (lane & 3) and 2 (lane & 3) + 1): the C fragment layout of pass 1.
template <int CS>
struct Updater {
  Lay L;
  int lane, uw;
  int gb, cls;                       // global panel counter % kPanelUpdWarps at the start of the next pass; class of the last pass
  int bsel; uint32_t ph;             // this warp's next hand-off buffer is uw + kPanelUpdWarps * bsel, its phase parity ph
  int done;                          // panels this warp has finished (published in upd_cnt)
  struct Pre { double2 s0, s1, s2, s3; double rho, rinv, ei; };
#ifdef BQP_PANEL_DEBUG
  long long tacc[6] = {0, 0, 0, 0, 0, 0}, tlast = 0; int npan = 0;
#endif

  template <int MODE>
  __device__ __forceinline__ void pass(const PanelShared &S, const WorkPtrs &W, int npanels, bool do_check, RowAcc &R) {
    const DevInstance &I = S.I;
    const int m = I.m, n = I.n;
    constexpr bool kIsA = (MODE == PM_A_INIT || MODE == PM_A_RESUME || MODE == PM_A_ITER || MODE == PM_A_CHK1 || MODE == PM_A_CHK2);
    constexpr bool kPass2 = kIsA || MODE == PM_P_CHK;
    const int r = lane >> 2, tp = lane & 3;
    const double alpha = I.alpha, oma = 1.0 - I.alpha;
    const int jlo = 32 * L.w0, jhi = 32 * (L.w0 + L.nwc);     // rows of x~ this CTA's consumers read
    constexpr int KU = kPanelUpdWarps;
    cls = uw - gb; if (cls < 0) cls += KU;   // this warp's panels of the pass: k = cls, cls + KU, ...
    // operands that do not depend on the partial sums.  They are fetched one owned panel AHEAD (right after the previous
    // panel is handed over), so their L2 / HBM latency never sits between "partials full" and "update done".
    auto load_state = [&](int k, int h) {
      Pre p;
      p.s0 = p.s1 = p.s2 = p.s3 = make_double2(0, 0); p.rho = p.rinv = p.ei = 0.0;
      const int row = k * kPR + h * 8 + r;
      if (k < npanels && (kIsA ? row < m : row < L.np)) {
        const int e2 = row * (T8 / 2) + tp;
        auto ld2 = [&](const double *v) { return reinterpret_cast<const double2 *>(v)[e2]; };
        if constexpr (MODE == PM_A_ITER) { p.s0 = ld2(W.gz); p.s1 = ld2(W.gy); p.s2 = ld2(W.gl); p.s3 = ld2(W.gu); p.rho = __ldg(I.rho + row); p.rinv = __ldg(I.rho_inv + row); }
        if constexpr (MODE == PM_A_INIT) { p.s1 = ld2(W.gy); p.rho = __ldg(I.rho + row); }
        if constexpr (MODE == PM_A_RESUME) { p.s0 = ld2(W.gz); p.s1 = ld2(W.gy); p.rho = __ldg(I.rho + row); }
        if constexpr (MODE == PM_A_CHK1) { p.s0 = ld2(W.gz); p.s1 = ld2(W.gy); p.ei = __ldg(I.Einv + row); }
        if constexpr (MODE == PM_A_CHK2) { p.s0 = ld2(W.gdy); p.s2 = ld2(W.gl); p.s3 = ld2(W.gu); p.ei = __ldg(I.Einv + row); p.rho = __ldg(I.E + row); }
        if constexpr (MODE == PM_M) { p.s0 = ld2(W.gxs); }
        if constexpr (MODE == PM_P_CHK) { p.s0 = ld2(W.gdx); }
        if constexpr (MODE == PM_P_OBJ) { p.s0 = ld2(W.gxs); p.rho = row < n ? __ldg(I.q + row) : 0.0; }
      }
      return p;
    };
    Pre nxt[kPH], nxt2[kPH];     // state of the next two owned panels (two deep: the loads may come from HBM)
#pragma unroll
    for (int h = 0; h < kPH; h++) { nxt[h] = load_state(cls, h); nxt2[h] = load_state(cls + KU, h); }
    for (int k = cls; k < npanels; k += KU) {
      Pre cur[kPH];
#pragma unroll
      for (int h = 0; h < kPH; h++) { cur[h] = nxt[h]; nxt[h] = nxt2[h]; nxt2[h] = load_state(k + 2 * KU, h); }
      const int hb = uw + KU * bsel;              // hand-off buffer of this panel
      TSTAMP(*this, 5);
      mbar_wait(L.pf + 8u * hb, ph);
      TSTAMP(*this, 0);
      mbar_wait(L.uc + 8u * hb, ph ^ 1u);
      TSTAMP(*this, 1);   // the pass-2 warps are done with this buffer's previous panel (u values, "update done" phase)
#pragma unroll
      for (int h = 0; h < kPH; h++) {
        const int row = k * kPR + h * 8 + r;
        const bool live = kIsA ? row < m : row < L.np;
        const int e2 = row * (T8 / 2) + tp;         // double2 index of (row, nodes 2tp, 2tp+1) in a [row][8] vector
        auto st2 = [&](double *v, double a, double b) { reinterpret_cast<double2 *>(v)[e2] = make_double2(a, b); };
        const double rho = cur[h].rho, rinv = cur[h].rinv, ei = cur[h].ei;
        double sum[2];
        {   // the pass-1 warps' partials in a fixed order: four interleaved chains, then a fixed tree
          const double2 *pp = reinterpret_cast<const double2 *>(L.part) + (hb * L.npt * kPH + h) * 32 + lane;
          double2 c0 = make_double2(0, 0), c1 = c0, c2 = c0, c3 = c0;
          int w = 0;
          for (; w + 4 <= L.npt; w += 4) {
            const double2 p0 = pp[(w + 0) * (kPH * 32)], p1 = pp[(w + 1) * (kPH * 32)], p2 = pp[(w + 2) * (kPH * 32)], p3 = pp[(w + 3) * (kPH * 32)];
            c0.x += p0.x; c0.y += p0.y; c1.x += p1.x; c1.y += p1.y; c2.x += p2.x; c2.y += p2.y; c3.x += p3.x; c3.y += p3.y;
          }
          for (; w < L.npt; w++) { const double2 p0 = pp[w * (kPH * 32)]; c0.x += p0.x; c0.y += p0.y; }
          sum[0] = (c0.x + c1.x) + (c2.x + c3.x);
          sum[1] = (c0.y + c1.y) + (c2.y + c3.y);
        }
        double u[2] = {0.0, 0.0};
        if (live) {
          const double a0[2] = {cur[h].s0.x, cur[h].s0.y}, a1[2] = {cur[h].s1.x, cur[h].s1.y}, a2[2] = {cur[h].s2.x, cur[h].s2.y},
                       a3[2] = {cur[h].s3.x, cur[h].s3.y};
          if constexpr (MODE == PM_M) {
            const double xn0 = alpha * sum[0] + oma * a0[0], xn1 = alpha * sum[1] + oma * a0[1];
            st2(W.gxs, xn0, xn1);
            if (row >= jlo && row < jhi) reinterpret_cast<double2 *>(L.xts)[(row - jlo) * (T8 / 2) + tp] = make_double2(sum[0], sum[1]);
            if (do_check) st2(W.gdx, xn0 - a0[0], xn1 - a0[1]);
          } else if constexpr (MODE == PM_A_ITER) {
            double zn[2], yn[2], dy[2];
#pragma unroll
            for (int i = 0; i < 2; i++) {
              const double zr = alpha * sum[i] + oma * a0[i];
              double z = zr + rinv * a1[i];
              z = fmin(fmax(z, a2[i]), a3[i]);
              dy[i] = rho * (zr - z); yn[i] = a1[i] + dy[i]; zn[i] = z;
              u[i] = fma(rho, z, -yn[i]);
            }
            st2(W.gz, zn[0], zn[1]); st2(W.gy, yn[0], yn[1]);
            if (do_check) st2(W.gdy, dy[0], dy[1]);
          } else if constexpr (MODE == PM_A_INIT) {
            st2(W.gz, sum[0], sum[1]);
            u[0] = fma(rho, sum[0], -a1[0]); u[1] = fma(rho, sum[1], -a1[1]);
          } else if constexpr (MODE == PM_A_RESUME) {
            u[0] = fma(rho, a0[0], -a1[0]); u[1] = fma(rho, a0[1], -a1[1]);
          } else if constexpr (MODE == PM_A_CHK1) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
              R.pr[i] = fmax(R.pr[i], fabs(ei * (sum[i] - a0[i])));
              R.a1[i] = fmax(R.a1[i], fabs(ei * sum[i]));
              R.a2[i] = fmax(R.a2[i], fabs(ei * a0[i]));
              u[i] = a1[i];
            }
          } else if constexpr (MODE == PM_A_CHK2) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
              const double v = ei * sum[i];
              if (a3[i] < kInfty * kMinScaling) R.vu[i] = fmax(R.vu[i], v);
              if (a2[i] > -kInfty * kMinScaling) R.vl[i] = fmin(R.vl[i], v);
              double d = a0[i];
              if (a3[i] > kInfty * kMinScaling) {
                if (a2[i] < -kInfty * kMinScaling) d = 0.0; else d = fmin(d, 0.0);
              } else if (a2[i] < -kInfty * kMinScaling) d = fmax(d, 0.0);
              R.ndy[i] = fmax(R.ndy[i], fabs(rho * d));       // rho holds E[row] in this mode
              R.lhs[i] += a3[i] * fmax(d, 0.0) + a2[i] * fmin(d, 0.0);
              u[i] = d;
            }
          } else if constexpr (MODE == PM_P_CHK) {
            st2(W.gpx, sum[0], sum[1]);
            u[0] = a0[0]; u[1] = a0[1];
          } else if constexpr (MODE == PM_P_OBJ) {
#pragma unroll
            for (int i = 0; i < 2; i++) { R.quad[i] += a0[i] * sum[i]; R.lin[i] += rho * a0[i]; }   // rho holds q[row]
          }
        }
        if constexpr (kPass2) reinterpret_cast<double2 *>(L.ubuf)[hb * (kPR * T8 / 2) + h * 32 + lane] = make_double2(u[0], u[1]);
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(L.ud + 8u * hb);
        __threadfence_block();
        done++;
        L.cnt[uw] = done;
        if constexpr (CS == 2) st_remote_u32(L.cnt_r + 4u * uw, (uint32_t)done);
      }
      if constexpr (kHBmul == 1) { ph ^= 1u; } else { if (bsel) ph ^= 1u; bsel ^= 1; }
      TSTAMP(*this, 2);
#ifdef BQP_PANEL_DEBUG
      npan++;
#endif
    }
    gb = (gb + npanels) % KU;
  }
};

template <int CS>
__global__ void __launch_bounds__(kPanelThreads, 1)
admm_panel_kernel(const DevInstance *__restrict__ insts, const DevTile *__restrict__ tiles, const double *__restrict__ in,
                  double *__restrict__ out, double *__restrict__ work, NodeScalars *__restrict__ ns,
                  int *__restrict__ tile_iters, int nslots, double *__restrict__ state, int prefetch_panels) {
  constexpr int KU = kPanelUpdWarps;
  constexpr int T = T8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = CS == 2 ? cluster_ctarank() : 0u, peer = rank ^ 1u;
  const int tile_id = blockIdx.x / CS;
  PanelShared &S = *reinterpret_cast<PanelShared *>(smem_raw);
  if (tid == 0) {
    S.tile = tiles[tile_id];
    S.I = insts[S.tile.inst];
    S.remaining = S.tile.nn;
  }
  if (tid < T) { S.status[tid] = BQP_UNSOLVED; S.iters[tid] = 0; S.newly[tid] = 0; S.dist[tid] = NAN; }
  if (tid < 2 * kPanelUpdWarps) S.upd_cnt[tid / kPanelUpdWarps][tid % kPanelUpdWarps] = 0;
  __syncthreads();
  const DevInstance &I = S.I;
  const int n = I.n, m = I.m, np = I.npad, nn = S.tile.nn, NW = I.p_nw;
  const int npm = I.p_npm, npa = I.p_npa;
  const int iter_begin = S.tile.iter_begin, iter_end = S.tile.iter_end;
  // this CTA's share of the column tiles
  const int nwh = CS == 2 ? (NW + 1) / 2 : NW;
  const int w0 = rank == 0 ? 0 : nwh, NWc = rank == 0 ? nwh : NW - nwh;
  const int slot_bytes = NWc * (kTileDoubles * 8);
  const int nwslots = (int)(blockDim.x >> 5) - KU - 1;     // consumer warp slots of this launch
  // pass-1 warps of ONE set / of all sets / pass-2 warps of this CTA
  const int np1 = (NWc + kP1T - 1) / kP1T, np1w = np1 * kSets, np2 = (NWc + 1) / 2, ncons = np1w + np2;
  const int np1_r0 = (nwh + kP1T - 1) / kP1T, npt = np1_r0 + (CS == 2 ? (NW - nwh + kP1T - 1) / kP1T : 0);
  const int nthr_cu = (ncons + KU) * 32, nthr_all = (ncons + KU + 1) * 32;
  const bool is_consumer = warp < ncons, is_update = warp >= nwslots && warp < nwslots + KU, is_producer = warp == nwslots + KU;

  Lay L;
  size_t off = (sizeof(PanelShared) + 15) & ~size_t(15);
  L.full = smem_u32(smem_raw + off);
  L.empty = L.full + 8u * nslots; L.pf = L.empty + 8u * nslots; L.ud = L.pf + 8u * kHB; L.uc = L.ud + 8u * kHB;
  L.ck = L.uc + 8u * kHB;
  off += sizeof(uint64_t) * (2 * (size_t)nslots + 3 * kHB + 1);
  off = (off + 15) & ~size_t(15);
  L.xts = reinterpret_cast<double *>(smem_raw + off);
  L.vs = L.xts + (size_t)nwh * 32 * T;
  L.part = L.vs + (size_t)nwh * 32 * T;
  L.ubuf = L.part + (size_t)kHB * npt * 64 * kPH;
  L.red = L.ubuf + (size_t)kHB * kPR * T;
  off += ((size_t)2 * nwh * 32 * T + (size_t)kHB * npt * 64 * kPH + (size_t)kHB * kPR * T + (size_t)kColQ * NW * T) * 8;
  off = (off + 127) & ~size_t(127);
  L.ring = smem_raw + off;
  L.ring_u32 = smem_u32(L.ring);
  L.nslots = nslots; L.slot_bytes = slot_bytes; L.nw = NW; L.nwc = NWc; L.w0 = w0; L.np = np; L.npt = npt; L.np1 = np1;
  if constexpr (CS == 2) {
    L.pf_r = mapa(L.pf, peer); L.ck_r = mapa(L.ck, peer);
    L.part_r = mapa(smem_u32(L.part), peer); L.red_r = mapa(smem_u32(L.red), peer);
    L.part_u32 = smem_u32(L.part);
    L.cnt_r = mapa(smem_u32(const_cast<int *>(&S.upd_cnt[1][0])), peer);
  } else {
    L.pf_r = L.ck_r = L.part_r = L.red_r = L.part_u32 = L.cnt_r = 0;
  }
  L.cnt = &S.upd_cnt[0][0];
  if (tid == 0) {
    for (int s = 0; s < nslots; s++) { mbar_init(L.full + 8u * s, 1); mbar_init(L.empty + 8u * s, np1 + np2); }
    // "partials full" / check barriers: one arrival per LOCAL pass-1 / pass-2 warp; the peer's share arrives as transaction bytes
    for (int s = 0; s < kHB; s++) {
      mbar_init(L.pf + 8u * s, np1); mbar_init(L.ud + 8u * s, 1); mbar_init(L.uc + 8u * s, np2);
    }
    mbar_init(L.ck, np2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
#ifdef BQP_PANEL_DEBUG
  if (tid == 0 && blockIdx.x == 0) printf("LAYOUT full %u empty %u pf %u ud %u uc %u ck %u nslots %d np1 %d npt %d ncons %d NWc %d\n", L.full, L.empty, L.pf, L.ud, L.uc, L.ck, nslots, np1, npt, ncons, NWc);
#endif
  if constexpr (CS == 2) cluster_sync_all(); else __syncthreads();   // barriers of both CTAs exist before anyone arrives
  if (!is_consumer && !is_update && !is_producer) return;            // CTA sized for the widest problem of the launch

  const int max_iter = I.max_iter, check_every = I.check_every;
  const double *pM = I.pstream + (size_t)w0 * kTileDoubles, *pA = I.pstream + I.p_offA + (size_t)w0 * kTileDoubles,
               *pP = I.pstream + I.p_offP + (size_t)w0 * kTileDoubles;

  // =============================================================== producer warp: mirror of the pass sequence
  if (is_producer) {
    int slot = 0; uint32_t phase = 0;
    // `nxt`/`nnxt`: the matrix that follows this one in the stream, for the L2 prefetch running `prefetch_panels` ahead
    auto produce = [&](const double *src, int npanels, const double *nxt, int nnxt) {
      if (lane == 0) {
        for (int k = 0; k < npanels; k++) {
          mbar_wait(L.empty + 8u * slot, phase ^ 1u);       // passes at once on the first lap
          mbar_expect_tx(L.full + 8u * slot, (uint32_t)slot_bytes);
          tma_load_1d(L.ring_u32 + (uint32_t)slot * slot_bytes, src + (size_t)k * I.p_panel_doubles, (uint32_t)slot_bytes, L.full + 8u * slot);
          if (prefetch_panels > 0) {
            const int kp = k + prefetch_panels;
            if (kp < npanels) l2_prefetch(src + (size_t)kp * I.p_panel_doubles, (uint32_t)slot_bytes);
            else if (kp - npanels < nnxt) l2_prefetch(nxt + (size_t)(kp - npanels) * I.p_panel_doubles, (uint32_t)slot_bytes);
          }
          if (++slot == nslots) { slot = 0; phase ^= 1u; }
        }
      }
      __syncwarp();
    };
    produce(pA, npa, pM, npm);
    for (int iter = iter_begin + 1; iter <= iter_end; iter++) {
      const bool do_check = (iter % check_every == 0) || iter == max_iter;
      produce(pM, npm, pA, npa); produce(pA, npa, do_check ? pA : pM, do_check ? npa : npm);
      if (!do_check) continue;
      produce(pA, npa, pA, npa); produce(pA, npa, pP, npm); produce(pP, npm, pM, npm);
      named_bar(1, nthr_all);                          // decision published
      if (S.remaining == 0 || iter == iter_end) break;
    }
    named_bar(1, nthr_all);                            // epilogue operands ready
    produce(pP, npm, pP, 0);
    return;
  }

  // =============================================================== consumers + update warps
  WorkPtrs W;
  {
    const size_t m8 = (size_t)((m + kPR - 1) / kPR * kPR);
    double *p = work + S.tile.work_off + (size_t)rank * panel_work_doubles(np, m);   // each CTA of a pair keeps its own copy
    W.gz = p; p += m8 * T; W.gy = p; p += m8 * T; W.gl = p; p += m8 * T; W.gu = p; p += m8 * T; W.gdy = p; p += m8 * T;
    W.gdx = p; p += (size_t)np * T; W.gpx = p; p += (size_t)np * T; W.gaty = p; p += (size_t)np * T;
    W.gatd = p; p += (size_t)np * T; W.gpdx = p; p += (size_t)np * T; W.gxs = p; p += (size_t)np * T; W.gsx = p;
  }
  const double sigma = I.sigma;
  const int ctid = is_consumer ? tid : ncons * 32 + (tid - nwslots * 32);   // dense index over consumer + update threads

  // ---- prologue (node.py:102-105): bounds, warm start (or the saved state of a resumed round)
  for (int e = ctid; e < (m + kPR - 1) / kPR * kPR * T; e += nthr_cu) {
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
  for (int e = ctid; e < np * T; e += nthr_cu) {
    const int j = e / T, t = e - j * T;
    double xv = 0.0;
    if (j < n && t < nn)
      xv = iter_begin == 0 ? __ldg(I.Dinv + j) * in[S.tile.in_off[t] + 2 * (size_t)m + j] : state[S.tile.state_off[t] + j];
    W.gxs[e] = xv;
  }
  named_bar(2, nthr_cu);

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
  // unscaled iterates of the nodes that terminated at this check (consumers + update warps): a private copy per CTA (the
  // objective operand of the epilogue), and the caller's buffers from CTA 0
  auto snapshot = [&]() {
    for (int t = 0; t < nn; t++) {
      if (!S.newly[t]) continue;
      const int st = S.status[t];
      const bool bad = !(st == BQP_SOLVED || st == BQP_SOLVED_INACCURATE || st == BQP_MAX_ITER_REACHED);
      double *ox = out + S.tile.out_off[t], *oy = ox + n, *sx = W.gsx + (size_t)t * np;
      for (int j = ctid; j < n; j += nthr_cu) {
        const double v = bad ? NAN : __ldg(I.D + j) * W.gxs[(size_t)j * T + t];
        sx[j] = v;
        if (rank == 0) ox[j] = v;
      }
      if (rank == 0)
        for (int i = ctid; i < m; i += nthr_cu) oy[i] = bad ? NAN : I.cinv * __ldg(I.E + i) * W.gy[(size_t)i * T + t];
    }
  };
  // end of the launch, both roles: save the state of unfinished nodes, clip + stage the objective operand
  auto finish_common = [&](int iter) {
    if (ctid == 0 && rank == 0) tile_iters[tile_id] = (iter > iter_end ? iter_end : iter) - iter_begin;
    if (rank == 0) {
      for (int t = 0; t < nn; t++) {
        if (S.status[t] != BQP_UNSOLVED) continue;
        double *sp = state + S.tile.state_off[t];
        for (int j = ctid; j < n; j += nthr_cu) sp[j] = W.gxs[(size_t)j * T + t];
        for (int i = ctid; i < m; i += nthr_cu) { sp[n + i] = W.gz[(size_t)i * T + t]; sp[n + m + i] = W.gy[(size_t)i * T + t]; }
        if (ctid == 0) {   // pri_res of a node that is still running carries its distance to the tolerance (scheduling hint)
          NodeScalars r; r.status = BQP_UNSOLVED; r.iters = iter_end; r.obj = r.dua_res = r.lower = NAN; r.pri_res = S.dist[t];
          ns[S.tile.node[t]] = r;
        }
      }
    }
    // epilogue (node.py:128-143): clip integer entries, lower = 1/2 x'Px + q'x at the clipped point
    for (int t = 0; t < nn; t++) {
      const int st = S.status[t];
      if (!(st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED)) continue;
      double *ox = out + S.tile.out_off[t], *sx = W.gsx + (size_t)t * np;
      const double *p = in + S.tile.in_off[t];
      for (int k = ctid; k < I.n_int; k += nthr_cu) {
        const int j = __ldg(I.i_idx + k), row = m - I.n_int + k;
        const double v = fmin(fmax(sx[j], p[row]), p[m + row]);
        sx[j] = v;
        if (rank == 0) ox[j] = v;
      }
    }
    named_bar(2, nthr_cu);
    for (int e = ctid; e < np * T; e += nthr_cu) {   // the x state is saved: its buffer takes the objective operand
      const int j = e / T, t = e - j * T;
      double v = 0.0;
      if (j < n && t < nn) {
        const int st = S.status[t];
        if (st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED) v = __ldg(I.Dinv + j) * W.gsx[(size_t)t * np + j];
      }
      W.gxs[e] = v;
    }
    named_bar(1, nthr_all);                            // with the producer: epilogue operands ready
  };

  if (is_update) {
    // ============================================================= update warps
    Updater<CS> U;
    U.L = L; U.lane = lane; U.uw = warp - nwslots; U.gb = 0; U.cls = 0; U.ph = 0; U.bsel = 0; U.done = 0;
    uint32_t ck_ph = 0;
    RowAcc R;
    auto reset = [&]() {
#pragma unroll
      for (int i = 0; i < 2; i++) { R.pr[i] = R.a1[i] = R.a2[i] = R.ndy[i] = R.lhs[i] = R.quad[i] = R.lin[i] = 0.0; R.vu[i] = -INFINITY; R.vl[i] = INFINITY; }
    };
    // this warp's row-space accumulators -> finp (one value per node)
    auto publish_rows = [&](int cls_sum) {
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const double v[kFinP] = {reduce_rows<0>(R.pr[i]), reduce_rows<0>(R.a1[i]), reduce_rows<0>(R.a2[i]), reduce_rows<0>(R.ndy[i]),
                                 reduce_rows<1>(R.lhs[i]), reduce_rows<0>(R.vu[i]), reduce_rows<2>(R.vl[i]), reduce_rows<1>(R.quad[i]),
                                 reduce_rows<1>(R.lin[i])};
        if (lane < 4) {   // maxima / minima: any order; sums (q = 4 lhs, 7 quad, 8 lin): by the class of the pass that made them
#pragma unroll
          for (int q = 0; q < kFinP; q++) S.finp[(q == 4 || q == 7 || q == 8) ? cls_sum : U.uw][q][2 * lane + i] = v[q];
        }
      }
    };
    reset();
    if (iter_begin == 0) U.template pass<PM_A_INIT>(S, W, npa, false, R); else U.template pass<PM_A_RESUME>(S, W, npa, false, R);
    int iter;
    for (iter = iter_begin + 1; iter <= iter_end; iter++) {
      const bool do_check = (iter % check_every == 0) || iter == max_iter;
      U.template pass<PM_M>(S, W, npm, do_check, R);
      U.template pass<PM_A_ITER>(S, W, npa, do_check, R);
      if (!do_check) continue;
      // CHK1 prefetches z, y of its first 2 KU panels before it waits for anything, and the panel -> warp ownership shifts from
      // pass to pass: those rows were written in the A pass just finished, possibly by a warp that is still up to kHB panels
      // behind.  With fewer than 2 KU + kHB + KU - 1 = 14 panels in A that prefetch read stale rows (found at m = 60: dual
      // residual and termination decision depended on timing).  Join the update warps first; once per check, so free.
      named_bar(3, KU * 32);
      reset();
      U.template pass<PM_A_CHK1>(S, W, npa, true, R);
      U.template pass<PM_A_CHK2>(S, W, npa, true, R);
      const int cls_chk2 = U.cls;
      U.template pass<PM_P_CHK>(S, W, npm, true, R);
      publish_rows(cls_chk2);
      named_bar(3, KU * 32);     // row-space partials of every update warp are in finp
      if (U.uw == 0) {
        mbar_wait(L.ck, ck_ph);   // every consumer warp (of both CTAs) has delivered its column-space partials
        for (int idx = lane; idx < kColQ * T; idx += 32) {
          const int q = idx / T, t = idx - q * T;
          const bool is_sum = (q == 3 || q == 4 || q == 8);
          double rr = L.red[((size_t)q * NW) * T + t];
          for (int w = 1; w < NW; w++) {
            const double v = L.red[((size_t)q * NW + w) * T + t];
            rr = is_sum ? rr + v : fmax(rr, v);
          }
          // slots: 0 dr, 1 |Px|, 2 |A'y|, 3 quad, 4 lin, 5 t1 -> fin[12], 6 t2 -> fin[13], 7 ndx -> fin[10], 8 qdx -> fin[11]
          const int slot = q < 5 ? q : (q == 5 ? 12 : (q == 6 ? 13 : (q == 7 ? 10 : 11)));
          S.fin[slot][t] = rr;
        }
        if (lane < T) {   // row-space quantities: maxima in any order, the sum class by class
          double pr = S.finp[0][0][lane], a1 = S.finp[0][1][lane], a2 = S.finp[0][2][lane], ndy = S.finp[0][3][lane],
                 lhs = S.finp[0][4][lane], vu = S.finp[0][5][lane], vl = S.finp[0][6][lane];
          for (int w = 1; w < KU; w++) {
            pr = fmax(pr, S.finp[w][0][lane]); a1 = fmax(a1, S.finp[w][1][lane]); a2 = fmax(a2, S.finp[w][2][lane]);
            ndy = fmax(ndy, S.finp[w][3][lane]); lhs += S.finp[w][4][lane];
            vu = fmax(vu, S.finp[w][5][lane]); vl = fmin(vl, S.finp[w][6][lane]);
          }
          S.fin[5][lane] = pr; S.fin[6][lane] = a1; S.fin[7][lane] = a2; S.fin[8][lane] = ndy; S.fin[9][lane] = lhs;
          S.fin[14][lane] = vu; S.fin[15][lane] = vl;
        }
        __syncwarp();
        if (lane < T) decide(lane, iter);
      }
      ck_ph ^= 1u;
      named_bar(2, nthr_cu);     // decision visible to the consumers
      snapshot();
      named_bar(1, nthr_all);    // ... and to the producer
      if (S.remaining == 0 || iter == iter_end) break;
    }
    finish_common(iter);
    reset();
    U.template pass<PM_P_OBJ>(S, W, npm, false, R);
#ifdef BQP_PANEL_DEBUG
    if (lane == 0 && blockIdx.x == 0 && U.uw == 0)
      printf("UPD panels %d: pf-wait %.0f uc-wait %.0f compute %.0f loop %.0f\n", U.npan, (double)U.tacc[0] / U.npan, (double)U.tacc[1] / U.npan,
             (double)U.tacc[2] / U.npan, (double)U.tacc[5] / U.npan);
#endif
    publish_rows(U.cls);
    named_bar(3, KU * 32);
    if (U.uw == 0 && lane < nn && rank == 0) {
      const int st = S.status[lane];
      if (st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED) {
        double qd = S.finp[0][7][lane], ln = S.finp[0][8][lane];
        for (int w = 1; w < KU; w++) { qd += S.finp[w][7][lane]; ln += S.finp[w][8][lane]; }
        ns[S.tile.node[lane]].lower = (0.5 * qd + ln) * I.cinv;
      }
    }
    return;
  }

  // =============================================================== consumer warps (pass-1 warps, then pass-2 warps)
  const int jcol0 = 32 * w0;     // first column held in xts / vs
  const int gq = lane >> 2, tq = lane & 3;
  const int ncthr = ncons * 32;
  int iter;
  if (warp < np1w) {
    // ------------------------------------------------------------- pass-1 warp
    P1Warp<CS> C;
    const int wi = warp % np1;                         // position inside its set
    C.init(L, lane, kP1T * wi, min(kP1T, NWc - kP1T * wi));
    C.pidx = (rank == 0 ? 0 : np1_r0) + wi;
    C.set = warp / np1;
    C.pass(npa, W.gxs, 0);                            // z = A x0 (first round)
    named_bar(4, ncthr);                              // b' of the pass-2 warps is in vs
    for (iter = iter_begin + 1; iter <= iter_end; iter++) {
      const bool do_check = (iter % check_every == 0) || iter == max_iter;
      C.pass(npm, L.vs, jcol0);                       // x~ = M b
      C.wait_progress(C.g); __threadfence_block();    // every row of x~ is in xts
      C.pass(npa, L.xts, jcol0);                      // z~ = A x~
      named_bar(4, ncthr);
      if (!do_check) continue;
      // termination check (update_info + check_termination): A x | A dx | P x
      C.pass(npa, W.gxs, 0);
      C.pass(npa, W.gdx, 0);
      C.pass(npm, W.gxs, 0);
      named_bar(2, nthr_cu);     // decision made
      snapshot();
      named_bar(1, nthr_all);
      if (S.remaining == 0 || iter == iter_end) break;
    }
    finish_common(iter);
    C.pass(npm, W.gxs, 0);                            // P x at the clipped point (sums taken by the update warps)
    C.wait_progress(C.g);   // the update warps of both CTAs are done with our partials: safe to leave the cluster
#ifdef BQP_PANEL_DEBUG
    if (lane == 0 && blockIdx.x == 0 && warp == 0)
      printf("P1 panels %d: full %.0f compute %.0f flow %.0f store %.0f loop %.0f\n", C.g, (double)C.tacc[0] / C.g, (double)C.tacc[1] / C.g,
             (double)C.tacc[2] / C.g, (double)C.tacc[3] / C.g, (double)C.tacc[5] / C.g);
#endif
    return;
  }
  // --------------------------------------------------------------- pass-2 warp
  P2Warp<CS> C;
  const int pw = warp - np1w;
  C.init(L, lane, 2 * pw, min(2, NWc - 2 * pw));
  double acc[2][4][2];
  // b' = sigma x - q + A'(rho z - y) for this warp's columns, from the pass-2 accumulators (C fragments), into vs (read
  // back as B fragments by the pass-1 warps: the M pass input)
  auto finalize_b = [&]() {
#pragma unroll
    for (int tl = 0; tl < 2; tl++) {
      if (tl < C.ntl) {
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
          const int j = 32 * (C.wg0 + tl) + 8 * mt + gq;
          double b0 = 0.0, b1 = 0.0;
          if (j < n) {
            const double qj = __ldg(I.q + j);
            const double2 xx = reinterpret_cast<const double2 *>(W.gxs)[j * (T / 2) + tq];
            b0 = sigma * xx.x - qj + acc[tl][mt][0];
            b1 = sigma * xx.y - qj + acc[tl][mt][1];
          }
          reinterpret_cast<double2 *>(L.vs)[(j - jcol0) * (T / 2) + tq] = make_double2(b0, b1);
        }
      }
    }
  };
  auto store_cols = [&](double *gvec) {
#pragma unroll
    for (int tl = 0; tl < 2; tl++) {
      if (tl < C.ntl) {
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
          const int j = 32 * (C.wg0 + tl) + 8 * mt + gq;
          reinterpret_cast<double2 *>(gvec)[j * (T / 2) + tq] = make_double2(acc[tl][mt][0], acc[tl][mt][1]);
        }
      }
    }
    __syncwarp();
  };
  C.pass(npa, acc);                                   // A'(rho z - y) of the starting point
  finalize_b();
  named_bar(4, ncthr);
  for (iter = iter_begin + 1; iter <= iter_end; iter++) {
    const bool do_check = (iter % check_every == 0) || iter == max_iter;
    C.skip(npm);                                      // x~ = M b has no second pass
    C.pass(npa, acc);                                 // b' += A' w
    finalize_b();
    named_bar(4, ncthr);
    if (!do_check) continue;

    // ---- termination check: A'y | A'dy | P dx   (b' stays in vs)
    C.pass(npa, acc); store_cols(W.gaty);
    C.pass(npa, acc); store_cols(W.gatd);
    C.pass(npm, acc); store_cols(W.gpdx);
    for (int tl = 0; tl < C.ntl; tl++) {
      // column-space quantities: lane <-> column 32 * (column tile) + lane.  Everything read here was written by this
      // warp (store_cols) or by update warps whose panels it has waited for.
      const int wg = C.wg0 + tl, j = wg * 32 + lane;
      const bool inr = j < n;
      const double di = inr ? __ldg(I.Dinv + j) : 0.0, dj = inr ? __ldg(I.D + j) : 0.0, qj = inr ? __ldg(I.q + j) : 0.0;
      auto put = [&](int q, int t, double v) {
        if (lane == 0) {
          const int ri = (q * NW + wg) * T + t;
          L.red[ri] = v;
          if constexpr (CS == 2) st_async_remote_f64(L.red_r + 8u * ri, v, L.ck_r);
        }
      };
      for (int t = 0; t < T; t++) {
        const size_t e = (size_t)j * T + t;
        const double px = inr ? W.gpx[e] : 0.0, aty = inr ? W.gaty[e] : 0.0, xj = inr ? W.gxs[e] : 0.0,
                     dxj = inr ? W.gdx[e] : 0.0, atd = inr ? W.gatd[e] : 0.0, pdx = inr ? W.gpdx[e] : 0.0;
        put(0, t, reduce_warp<0>(fabs(di * (px + qj + aty))));
        put(1, t, reduce_warp<0>(fabs(di * px)));
        put(2, t, reduce_warp<0>(fabs(di * aty)));
        put(3, t, reduce_warp<1>(xj * px));
        put(4, t, reduce_warp<1>(qj * xj));
        put(5, t, reduce_warp<0>(fabs(di * atd)));
        put(6, t, reduce_warp<0>(fabs(di * pdx)));
        put(7, t, reduce_warp<0>(fabs(dj * dxj)));
        put(8, t, reduce_warp<1>(qj * dxj));
      }
    }
    if (lane == 0) {
      if (CS == 2 && pw == 0) mbar_expect_tx(L.ck, (uint32_t)((NW - NWc) * kColQ * T * 8));
      else mbar_arrive(L.ck);
    }
    named_bar(2, nthr_cu);     // decision made
    snapshot();
    named_bar(1, nthr_all);
    if (S.remaining == 0 || iter == iter_end) break;
  }
  finish_common(iter);
  C.skip(npm);
#ifdef BQP_PANEL_DEBUG
  if (lane == 0 && blockIdx.x == 0 && pw == 0)
    printf("P2 panels %d: ud %.0f full %.0f compute %.0f release %.0f loop %.0f\n", C.g, (double)C.tacc[0] / C.g, (double)C.tacc[1] / C.g,
           (double)C.tacc[2] / C.g, (double)C.tacc[3] / C.g, (double)C.tacc[5] / C.g);
#endif
}

}  // namespace

size_t panel_smem_bytes(int npad, int nslots, int cs) {
  const int nw = npad / 32;
  const int nwh = cs == 2 ? (nw + 1) / 2 : nw;
  const int npt = (nwh + kP1T - 1) / kP1T + (cs == 2 ? (nw - nwh + kP1T - 1) / kP1T : 0);
  size_t off = (sizeof(PanelShared) + 15) & ~size_t(15);
  off += sizeof(uint64_t) * (2 * (size_t)nslots + 3 * (size_t)kHB + 1);
  off = (off + 15) & ~size_t(15);
  off += ((size_t)2 * nwh * 32 * T8 + (size_t)kHB * npt * 64 * kPH + (size_t)kHB * kPR * T8 + (size_t)kColQ * nw * T8) * 8;
  off = (off + 127) & ~size_t(127);
  return off + (size_t)nslots * nwh * (kTileDoubles * 8);
}

template <int CS>
static int launch_p(int nw_max, int nslots, double *d_state, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles,
                    const double *d_in, double *d_out, double *d_work, NodeScalars *d_ns, int *d_tile_iters, size_t smem,
                    cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(admm_panel_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return BQP_E_CUDA;
  int prefetch_panels = 0;   // L2 prefetch distance of the producer, in panels (experiment knob)
  if (const char *pk = getenv("BQP_PANEL_PREFETCH")) prefetch_panels = atoi(pk);
  const int nwc = CS == 2 ? (nw_max + 1) / 2 : nw_max;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ntiles * CS), 1, 1);
  cfg.blockDim = dim3((unsigned)(panel_cta_warps(nwc) * 32), 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, admm_panel_kernel<CS>, d_insts, d_tiles, d_in, d_out, d_work, d_ns, d_tile_iters, nslots, d_state,
                         prefetch_panels);
  return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? BQP_OK : BQP_E_CUDA;
}

// cs = CTAs per tile (1, or 2 = a cluster pair splitting the columns).  
int launch_admm_panel(int cs, int nw_max, int nslots, double *d_state, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles,
                      const double *d_in, double *d_out, double *d_work, NodeScalars *d_ns, int *d_tile_iters,
                      size_t smem_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (nw_max < 1 || nw_max > kPanelMaxWarps || (cs != 1 && cs != 2) || (cs == 1 && nw_max > kPanelCtaWarps)) return BQP_E_ARG;
  if (cs == 1) return launch_p<1>(nw_max, nslots, d_state, d_insts, d_tiles, ntiles, d_in, d_out, d_work, d_ns, d_tile_iters, smem_bytes, st);
  return launch_p<2>(nw_max, nslots, d_state, d_insts, d_tiles, ntiles, d_in, d_out, d_work, d_ns, d_tile_iters, smem_bytes, st);
}

}  // namespace bqp
