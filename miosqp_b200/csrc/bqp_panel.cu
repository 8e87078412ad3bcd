// bqp_panel.cu -- fused single-pass batched ADMM kernel for sm_100a (dense A, npad <= 512).
//
// Same node-tile ownership as the other kernels (a tile = up to T <= 4 B&B leaves of one problem, whole OSQP loop
// in-kernel; /root/reference/miosqp/node.py:96-143), but the iteration is restated so that A is streamed from HBM ONCE
// per ADMM iteration instead of twice (A' then A) and the triangular sweeps disappear:
//
//     x~ = M b                         M = (P + sigma I + A' rho A)^-1, explicit, one dependency-free mat-vec
//     z~ = A x~ ; z,y update ; b' = sigma x - q + A'(rho z - y)     ONE pass over A
//
// Every matrix is cut into row PANELS (kPanelRows rows x npad columns, bqp_internal.h), streamed by TMA bulk copies
// into a ring of shared-memory slots.  While panel k of A sits in shared memory it is used twice:
//   pass 1   z~_I = A_I x~            consumer warp w owns columns 32w..32w+31 (x~ in registers), partial sums are
//                                     reduced over the 8 column lanes by a transposing shuffle tree and handed to the
//   update   z_I, y_I, w_I            UPDATE WARPS (one lane per (row, node)), which add the warp partials in a fixed
//                                     order, apply the projection / dual update and publish w_I = rho z_I - y_I;
//   pass 2   b' += A_I' w_I           same panel, accumulators stay in consumer registers for the whole pass.
// pass 2 runs LAG panels behind pass 1, so the update latency is hidden.
//
// Problems wider than 8 column tiles run as a CLUSTER OF TWO CTAs (one SM each): CTA r streams and multiplies only
// its half of the columns of every panel (half the HBM stream, shared-memory traffic and FP64 work per SM, 168 registers
// per thread), the per-warp partial sums of pass 1 are written into BOTH CTAs' shared memory (st.shared::cluster + remote
// mbarrier arrive over DSMEM), and both CTAs run the (cheap) row-space update redundantly, so nothing else crosses.
// Roles per CTA: warps [0, NWc) consumers, kPanelUpdWarps update warps, one TMA producer warp (one lane).
// Synchronisation inside a pass is mbarrier-only (full/empty per ring slot, "partials full" / "update done" per
// hand-off buffer); named barriers only at termination checks.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "bqp_internal.h"

namespace bqp {

namespace {

constexpr int kPR = kPanelRows;
constexpr int kColQ = 9;          // column-space reductions per termination check
constexpr int kFin = 16;
constexpr int kFinP = 9;          // row-space quantities each update warp accumulates
constexpr int kHB = kPanelUpdWarps;   // hand-off buffers; buffer b = panel % kHB always belongs to update warp b
constexpr int kPanelThreads = (kPanelCtaWarps + kPanelUpdWarps + 1) * 32;

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
// flow-control-only arrive on a peer barrier (no data rides on it): relaxed, so it does not wait for this thread's
// earlier global stores the way a cluster-scope release would
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t rbar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
template <bool CLUSTER>
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  if constexpr (CLUSTER) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
  return ok != 0;
}
// Bounded wait: a broken protocol traps (reported as a CUDA error) instead of hanging the GPU.
template <bool CLUSTER = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait<CLUSTER>(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait<CLUSTER>(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) __trap();   // ~10 s at 2 GHz
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
// asynchronous operation, so the sender needs no cluster-scope fence (a release.cluster arrive costs a MEMBAR.ALL.GPU)
__device__ __forceinline__ void st_async_remote_f64(uint32_t raddr, double v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(raddr), "d"(v), "r"(rbar) : "memory");
}

// Position of element (column j, node t) in the shared-memory column vectors xs / xts / vs: [column tile][b][node pair]
// [column lane][2], the order in which a consumer lane (columns 32w + 4cg + b) reads its 16-byte pieces.
template <int T>
__device__ __forceinline__ int vidx(int j, int t) {
  const int cwb = ((j >> 5) << 2) + (j & 3), cg = (j >> 2) & 7;
  if constexpr (T == 1) return cwb * 8 + cg;
  else return ((cwb * (T / 2) + (t >> 1)) * 8 + cg) * 2 + (t & 1);
}

struct PanelShared {
  DevInstance I;
  DevTile tile;
  double fin[kFin][4];
  double finp[kPanelUpdWarps][kFinP][4];
  int status[4], iters[4], newly[4];
  int remaining;
};

// everything a role needs to find its way around shared memory (u32 = shared-window addresses; *_r = the same object in
// the peer CTA of the pair, mapped with mapa)
struct Lay {
  uint32_t full, empty, pf, ud, udp, ck; // barrier arrays: [nslots], [nslots], [kHB], [kHB], [kHB], [1]
  uint32_t pf_r, udp_r, ck_r, part_r, red_r;
  double *xs, *xts, *vs, *part, *ubuf, *red;   // [np][T] x3, [kHB][NW][8T], [kHB][8T], [kColQ][NW][T]
  unsigned char *ring;
  uint32_t ring_u32;
  int nslots, slot_bytes, nw, nwc, w0, np;    // nw: column tiles of the problem; nwc, w0: this CTA's share
};

// ---- transposing shuffle reduction over the 8 column lanes (lane bits 0..2).  C values per lane go in; after the three
// steps every (row, node) sum lives in exactly one lane of each group of 8 (duplicated when 2T < 8): idx says which.
template <int C, int MK>
__device__ __forceinline__ void tstep(double *a, int lane, int &idx) {
  if constexpr (C >= 2) {
    constexpr int H = C / 2;
    const bool up = (lane & MK) != 0;
#pragma unroll
    for (int i = 0; i < H; i++) {
      const double send = up ? a[i] : a[i + H];
      const double keep = up ? a[i + H] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, MK);
    }
    if (up) idx += H;
  } else {
    a[0] += __shfl_xor_sync(0xffffffffu, a[0], MK);
  }
}

// LAG = how many panels pass 2 runs behind pass 1 (the update latency it hides), LAG < kHB.  CS = CTAs per tile.
template <int T, int LAG, int CS>
struct Consumer {
  static_assert(LAG < kHB, "pass 2 may lag at most kHB - 1 panels");
  Lay L;
  int cw, lane, rg, cg;
  int slot; uint32_t phase;          // ring position of the next pass-1 panel
  int slot2;                         // ring position of the next pass-2 panel
  int g, gb;                         // global panel counter (same sequence in the update warps), g % kHB
  int ud_g, ud_b; uint32_t ud_ph;    // next panel whose "update done" barrier this thread has not observed yet
  int up_g, up_b; uint32_t up_ph;    // the same for the peer CTA's update warps (flow control of the DSMEM partials)
  bool writer;

  __device__ __forceinline__ void wait_ud(int target) {
    while (ud_g <= target) {
      mbar_wait(L.ud + 8u * ud_b, ud_ph);
      ud_g++;
      if (++ud_b == kHB) { ud_b = 0; ud_ph ^= 1u; }
    }
  }
  __device__ __forceinline__ void wait_udp(int target) {   // the peer's update warp has read our partials of that panel
    if constexpr (CS == 2) {
      while (up_g <= target) {
        mbar_wait(L.udp + 8u * up_b, up_ph);
        up_g++;
        if (++up_b == kHB) { up_b = 0; up_ph ^= 1u; }
      }
    }
  }
  __device__ __forceinline__ int col0() const { return 32 * (L.w0 + cw) + 4 * cg; }   // first of this lane's 4 columns
  // one pass over `npanels` panels: pass 1 with the column vector `vsrc` (shared memory, vidx layout; this lane keeps its
  // own 4 columns x T nodes in registers); with PASS2 the per-row values published by the update warps are multiplied
  // back into acc (this lane's 4 columns x T nodes) LAG panels later.
  template <bool PASS2>
  __device__ __forceinline__ void pass(int npanels, const double *vsrc, double (&acc)[4][T]) {
    const int g0 = g;
    const int aoff = ((cw * 4) * 32 + lane) * 16;   // bytes into the slot
    double xv[4][T];
    {
      const int c0 = col0();
#pragma unroll
      for (int b = 0; b < 4; b++)
#pragma unroll
        for (int t = 0; t < T; t++) xv[b][t] = vsrc[vidx<T>(c0 + b, t)];
    }
    int g2b = gb;
    slot2 = slot;
    const int wg = L.w0 + cw;
    const int nsteps = npanels + (PASS2 ? LAG : 0);
    for (int k = 0; k < nsteps; k++) {
      if (k < npanels) {
        mbar_wait(L.full + 8u * slot, phase);
        const double2 *ap = reinterpret_cast<const double2 *>(L.ring + (size_t)slot * L.slot_bytes + aoff);
        double2 a[4];
#pragma unroll
        for (int b = 0; b < 4; b++) a[b] = ap[b * 32];
        double zp[2 * T];
#pragma unroll
        for (int t = 0; t < T; t++) { zp[t] = a[0].x * xv[0][t]; zp[T + t] = a[0].y * xv[0][t]; }
#pragma unroll
        for (int b = 1; b < 4; b++)
#pragma unroll
          for (int t = 0; t < T; t++) { zp[t] = fma(a[b].x, xv[b][t], zp[t]); zp[T + t] = fma(a[b].y, xv[b][t], zp[T + t]); }
        int idx = 0;
        tstep<2 * T, 1>(zp, lane, idx);
        tstep<T, 2>(zp, lane, idx);
        tstep<(T >= 4 ? T / 2 : 1), 4>(zp, lane, idx);
        wait_ud(g - kHB); wait_udp(g - kHB);   // the update warps of both CTAs have consumed this partials buffer
        if (writer) {
          const int pi = (gb * L.nw + wg) * (kPR * T) + 2 * rg * T + idx;
          L.part[pi] = zp[0];
          if constexpr (CS == 2) st_async_remote_f64(L.part_r + 8u * pi, zp[0], L.pf_r + 8u * gb);
        }
        __syncwarp();
        if (lane == 0) {
          // pair: the barrier also counts the bytes the peer's consumer warps store into our buffer (posted by warp 0)
          if (CS == 2 && cw == 0) mbar_expect_tx(L.pf + 8u * gb, (uint32_t)((L.nw - L.nwc) * (kPR * T) * 8));
          else mbar_arrive(L.pf + 8u * gb);
          if (!PASS2) mbar_arrive(L.empty + 8u * slot);
        }
        if (++slot == L.nslots) { slot = 0; phase ^= 1u; }
        g++;
        if (++gb == kHB) gb = 0;
      }
      if (PASS2 && k >= LAG) {
        wait_ud(g0 + k - LAG);
        const double *up = L.ubuf + g2b * (kPR * T) + 2 * rg * T;
        double u[2 * T];
        if constexpr (T == 1) {
          const double2 v = *reinterpret_cast<const double2 *>(up);
          u[0] = v.x; u[1] = v.y;
        } else {
#pragma unroll
          for (int i = 0; i < 2 * T; i += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(up + i);
            u[i] = v.x; u[i + 1] = v.y;
          }
        }
        const double2 *ap = reinterpret_cast<const double2 *>(L.ring + (size_t)slot2 * L.slot_bytes + aoff);
        double2 a[4];
#pragma unroll
        for (int b = 0; b < 4; b++) a[b] = ap[b * 32];
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
          for (int t = 0; t < T; t++) { acc[b][t] = fma(a[b].x, u[t], acc[b][t]); acc[b][t] = fma(a[b].y, u[T + t], acc[b][t]); }
        __syncwarp();
        if (lane == 0) mbar_arrive(L.empty + 8u * slot2);
        if (++slot2 == L.nslots) slot2 = 0;
        if (++g2b == kHB) g2b = 0;
      }
    }
  }
  // column sums of pass 2 live spread over the 4 row groups of the warp: butterfly all-reduce (lane bits 3, 4)
  __device__ __forceinline__ void allreduce_rg(double (&acc)[4][T]) {
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
      for (int t = 0; t < T; t++) {
        double v = acc[b][t];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        acc[b][t] = v;
      }
  }
};

template <int T>
__device__ __forceinline__ void zero4(double (&a)[4][T]) {
#pragma unroll
  for (int b = 0; b < 4; b++)
#pragma unroll
    for (int t = 0; t < T; t++) a[b][t] = 0.0;
}

// reduce v over the lanes that hold the same node (lane % T): xor masks 16 .. T
template <int T, int OP>   // OP 0: max, 1: sum, 2: min
__device__ __forceinline__ double reduce_same_node(double v) {
#pragma unroll
  for (int o = 16; o >= T; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = OP == 0 ? fmax(v, w) : (OP == 1 ? v + w : fmin(v, w));
  }
  return v;
}
template <int OP>
__device__ __forceinline__ double reduce_warp(double v) { return reduce_same_node<1, OP>(v); }

enum { PM_M = 0, PM_A_INIT, PM_A_RESUME, PM_A_ITER, PM_A_CHK1, PM_A_CHK2, PM_P_CHK, PM_P_OBJ };

struct WorkPtrs {
  double *gz, *gy, *gl, *gu, *gdy, *gdx, *gpx, *gaty, *gatd, *gpdx, *gsx;
};

// accumulators of an update warp (one lane per (row-in-panel, node)), reduced over rows at decision time
struct RowAcc {
  double pr, a1, a2, vu, vl, ndy, lhs, quad, lin;
};

// Update warps: warp uw handles the panels whose hand-off buffer is uw (global panel counter % kHB), so it waits on its
// "partials full" barrier strictly phase by phase.  Within one pass that is every kHB-th panel starting at some class
// c = k % kHB: `cls` (set by pass()) names it, so that sums over rows can be combined class by class -- a fixed order
// whatever ran earlier in the launch.  With CS == 2 both CTAs of the pair run this redundantly on identical inputs.
template <int T, int CS>
struct Updater {
  Lay L;
  int lane, uw;
  int g, gb, cls; uint32_t gph;      // global panel counter, g % kHB, class of the last pass, parity (g / kHB) & 1
  bool active;

  template <int MODE>
  __device__ __forceinline__ void pass(const PanelShared &S, const WorkPtrs &W, int npanels, bool do_check, RowAcc &R) {
    const DevInstance &I = S.I;
    const int m = I.m, n = I.n;
    constexpr bool kIsA = (MODE == PM_A_INIT || MODE == PM_A_RESUME || MODE == PM_A_ITER || MODE == PM_A_CHK1 || MODE == PM_A_CHK2);
    constexpr bool kPass2 = kIsA || MODE == PM_P_CHK;
    const int r = lane / T, t = lane % T;
    const double alpha = I.alpha, oma = 1.0 - I.alpha;
    cls = uw - gb; if (cls < 0) cls += kHB;   // this warp's panels of the pass: k = cls, cls + kHB, ...
    for (int k = 0; k < npanels; k++) {
      if (gb == uw) {
        const int row = k * kPR + r;
        const bool live = active && (kIsA ? row < m : row < L.np);
        const int e = k * (kPR * T) + lane;        // row * T + t
        const int ev = vidx<T>(row, t);            // same element in the shared-memory column vectors
        // ---- operands that do not depend on the partial sums: fetched before the wait
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0, rho = 0, rinv = 0, ei = 0;
        if (live) {
          if constexpr (MODE == PM_A_ITER) { s0 = W.gz[e]; s1 = W.gy[e]; s2 = W.gl[e]; s3 = W.gu[e]; rho = __ldg(I.rho + row); rinv = __ldg(I.rho_inv + row); }
          if constexpr (MODE == PM_A_INIT) { s1 = W.gy[e]; rho = __ldg(I.rho + row); }
          if constexpr (MODE == PM_A_RESUME) { s0 = W.gz[e]; s1 = W.gy[e]; rho = __ldg(I.rho + row); }
          if constexpr (MODE == PM_A_CHK1) { s0 = W.gz[e]; s1 = W.gy[e]; ei = __ldg(I.Einv + row); }
          if constexpr (MODE == PM_A_CHK2) { s0 = W.gdy[e]; s2 = W.gl[e]; s3 = W.gu[e]; ei = __ldg(I.Einv + row); rho = __ldg(I.E + row); }
          if constexpr (MODE == PM_M) { s0 = L.xs[ev]; }
          if constexpr (MODE == PM_P_CHK) { s0 = W.gdx[e]; }
          if constexpr (MODE == PM_P_OBJ) { s0 = L.xs[ev]; s1 = row < n ? __ldg(I.q + row) : 0.0; }
        }
        mbar_wait(L.pf + 8u * gb, gph);
        double sum = 0.0;
        if (active) {   // the NW warp partials in a fixed order: four interleaved chains, then a fixed tree
          const double *pp = L.part + gb * L.nw * (kPR * T) + lane;
          double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
          int w = 0;
          for (; w + 4 <= L.nw; w += 4) {
            c0 += pp[(w + 0) * (kPR * T)]; c1 += pp[(w + 1) * (kPR * T)];
            c2 += pp[(w + 2) * (kPR * T)]; c3 += pp[(w + 3) * (kPR * T)];
          }
          for (; w < L.nw; w++) c0 += pp[w * (kPR * T)];
          sum = (c0 + c1) + (c2 + c3);
        }
        double u = 0.0;
        if (live) {
          if constexpr (MODE == PM_M) {
            const double xn = alpha * sum + oma * s0;
            L.xs[ev] = xn; L.xts[ev] = sum;
            if (do_check) W.gdx[e] = xn - s0;
          } else if constexpr (MODE == PM_A_ITER) {
            const double zr = alpha * sum + oma * s0;
            double zn = zr + rinv * s1;
            zn = fmin(fmax(zn, s2), s3);
            const double dy = rho * (zr - zn), yn = s1 + dy;
            W.gz[e] = zn; W.gy[e] = yn;
            if (do_check) W.gdy[e] = dy;
            u = fma(rho, zn, -yn);
          } else if constexpr (MODE == PM_A_INIT) {
            W.gz[e] = sum;
            u = fma(rho, sum, -s1);
          } else if constexpr (MODE == PM_A_RESUME) {
            u = fma(rho, s0, -s1);
          } else if constexpr (MODE == PM_A_CHK1) {
            R.pr = fmax(R.pr, fabs(ei * (sum - s0)));
            R.a1 = fmax(R.a1, fabs(ei * sum));
            R.a2 = fmax(R.a2, fabs(ei * s0));
            u = s1;
          } else if constexpr (MODE == PM_A_CHK2) {
            const double v = ei * sum;
            if (s3 < kInfty * kMinScaling) R.vu = fmax(R.vu, v);
            if (s2 > -kInfty * kMinScaling) R.vl = fmin(R.vl, v);
            double d = s0;
            if (s3 > kInfty * kMinScaling) {
              if (s2 < -kInfty * kMinScaling) d = 0.0; else d = fmin(d, 0.0);
            } else if (s2 < -kInfty * kMinScaling) d = fmax(d, 0.0);
            R.ndy = fmax(R.ndy, fabs(rho * d));       // rho holds E[row] in this mode
            R.lhs += s3 * fmax(d, 0.0) + s2 * fmin(d, 0.0);
            u = d;
          } else if constexpr (MODE == PM_P_CHK) {
            W.gpx[e] = sum;
            u = s0;
          } else if constexpr (MODE == PM_P_OBJ) {
            R.quad += s0 * sum;
            R.lin += s1 * s0;
          }
        }
        if constexpr (kPass2) { if (active) L.ubuf[gb * (kPR * T) + lane] = u; }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(L.ud + 8u * gb);
          if constexpr (CS == 2) mbar_arrive_remote_relaxed(L.udp_r + 8u * gb);
        }
      }
      g++;
      if (++gb == kHB) { gb = 0; gph ^= 1u; }
    }
  }
};

template <int T, int LAG, int CS>
__global__ void __launch_bounds__(kPanelThreads, 1)
admm_panel_kernel(const DevInstance *__restrict__ insts, const DevTile *__restrict__ tiles, const double *__restrict__ in,
                  double *__restrict__ out, double *__restrict__ work, NodeScalars *__restrict__ ns,
                  int *__restrict__ tile_iters, int nslots, double *__restrict__ state, int prefetch_panels) {
  constexpr int KU = kPanelUpdWarps;
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
  if (tid < 4) { S.status[tid] = BQP_UNSOLVED; S.iters[tid] = 0; S.newly[tid] = 0; }
  __syncthreads();
  const DevInstance &I = S.I;
  const int n = I.n, m = I.m, np = I.npad, nn = S.tile.nn, NW = I.p_nw;
  const int npm = I.p_npm, npa = I.p_npa;
  const int iter_begin = S.tile.iter_begin, iter_end = S.tile.iter_end;
  // this CTA's share of the column tiles
  const int nwh = CS == 2 ? (NW + 1) / 2 : NW;
  const int w0 = rank == 0 ? 0 : nwh, NWc = rank == 0 ? nwh : NW - nwh;
  const int slot_bytes = NWc * (kPR * 32 * 8);
  const int nwslots = (int)(blockDim.x >> 5) - KU - 1;     // consumer warp slots of this launch
  const int nthr_cu = (NWc + KU) * 32, nthr_all = (NWc + KU + 1) * 32;
  const bool is_consumer = warp < NWc, is_update = warp >= nwslots && warp < nwslots + KU, is_producer = warp == nwslots + KU;

  Lay L;
  size_t off = (sizeof(PanelShared) + 15) & ~size_t(15);
  L.full = smem_u32(smem_raw + off);
  L.empty = L.full + 8u * nslots; L.pf = L.empty + 8u * nslots; L.ud = L.pf + 8u * kHB; L.udp = L.ud + 8u * kHB;
  L.ck = L.udp + 8u * kHB;
  off += sizeof(uint64_t) * (2 * (size_t)nslots + 3 * kHB + 1);
  off = (off + 15) & ~size_t(15);
  L.xs = reinterpret_cast<double *>(smem_raw + off);
  L.xts = L.xs + (size_t)np * T;
  L.vs = L.xts + (size_t)np * T;
  L.part = L.vs + (size_t)np * T;
  L.ubuf = L.part + (size_t)kHB * NW * kPR * T;
  L.red = L.ubuf + (size_t)kHB * kPR * T;
  off += ((size_t)3 * np * T + (size_t)kHB * NW * kPR * T + (size_t)kHB * kPR * T + (size_t)kColQ * NW * T) * 8;
  off = (off + 127) & ~size_t(127);
  L.ring = smem_raw + off;
  L.ring_u32 = smem_u32(L.ring);
  L.nslots = nslots; L.slot_bytes = slot_bytes; L.nw = NW; L.nwc = NWc; L.w0 = w0; L.np = np;
  if constexpr (CS == 2) {
    L.pf_r = mapa(L.pf, peer); L.udp_r = mapa(L.udp, peer); L.ck_r = mapa(L.ck, peer);
    L.part_r = mapa(smem_u32(L.part), peer); L.red_r = mapa(smem_u32(L.red), peer);
  } else {
    L.pf_r = L.udp_r = L.ck_r = L.part_r = L.red_r = 0;
  }
  if (tid == 0) {
    for (int s = 0; s < nslots; s++) { mbar_init(L.full + 8u * s, 1); mbar_init(L.empty + 8u * s, NWc); }
    // "partials full" / check barriers: one arrival per LOCAL consumer warp; the peer's share arrives as transaction bytes
    for (int s = 0; s < kHB; s++) { mbar_init(L.pf + 8u * s, NWc); mbar_init(L.ud + 8u * s, 1); mbar_init(L.udp + 8u * s, 1); }
    mbar_init(L.ck, NWc);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if constexpr (CS == 2) cluster_sync_all(); else __syncthreads();   // barriers of both CTAs exist before anyone arrives
  if (!is_consumer && !is_update && !is_producer) return;            // CTA sized for the widest problem of the launch

  const int max_iter = I.max_iter, check_every = I.check_every;
  const double *pM = I.pstream + (size_t)w0 * (kPR * 32), *pA = I.pstream + I.p_offA + (size_t)w0 * (kPR * 32),
               *pP = I.pstream + I.p_offP + (size_t)w0 * (kPR * 32);

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
    const size_t m8 = (size_t)((m + 7) & ~7);
    double *p = work + S.tile.work_off + (size_t)rank * panel_work_doubles(np, m, T);   // each CTA of a pair keeps its own copy
    W.gz = p; p += m8 * T; W.gy = p; p += m8 * T; W.gl = p; p += m8 * T; W.gu = p; p += m8 * T; W.gdy = p; p += m8 * T;
    W.gdx = p; p += (size_t)np * T; W.gpx = p; p += (size_t)np * T; W.gaty = p; p += (size_t)np * T;
    W.gatd = p; p += (size_t)np * T; W.gpdx = p; p += (size_t)np * T; W.gsx = p;
  }
  const double sigma = I.sigma;
  const int ctid = is_consumer ? tid : NWc * 32 + (tid - nwslots * 32);   // dense index over consumer + update threads

  // ---- prologue (node.py:102-105): bounds, warm start (or the saved state of a resumed round)
  for (int e = ctid; e < m * T; e += nthr_cu) {
    const int i = e / T, t = e - i * T;
    double lo = -kInfty, up = kInfty, yv = 0.0, zv = 0.0;
    if (t < nn) {
      const double *p = in + S.tile.in_off[t];
      lo = fmax(p[i], -kInfty);
      up = fmin(p[m + i], kInfty);
      if (iter_begin == 0) yv = I.c * __ldg(I.Einv + i) * p[2 * (size_t)m + n + i];
      else { const double *sp = state + S.tile.state_off[t] + n; zv = sp[i]; yv = sp[m + i]; }
    }
    const double ei = __ldg(I.E + i);
    W.gl[e] = ei * lo; W.gu[e] = ei * up; W.gy[e] = yv; W.gz[e] = zv;
  }
  for (int e = ctid; e < np * T; e += nthr_cu) {
    const int j = e / T, t = e - j * T;
    double xv = 0.0;
    if (j < n && t < nn)
      xv = iter_begin == 0 ? __ldg(I.Dinv + j) * in[S.tile.in_off[t] + 2 * (size_t)m + j] : state[S.tile.state_off[t] + j];
    L.xs[vidx<T>(j, t)] = xv; L.xts[e] = 0.0;
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
        const double v = bad ? NAN : __ldg(I.D + j) * L.xs[vidx<T>(j, t)];
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
        for (int j = ctid; j < n; j += nthr_cu) sp[j] = L.xs[vidx<T>(j, t)];
        for (int i = ctid; i < m; i += nthr_cu) { sp[n + i] = W.gz[(size_t)i * T + t]; sp[n + m + i] = W.gy[(size_t)i * T + t]; }
        if (ctid == 0) { NodeScalars r; r.status = BQP_UNSOLVED; r.iters = iter_end; r.obj = r.pri_res = r.dua_res = r.lower = NAN; ns[S.tile.node[t]] = r; }
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
    for (int e = ctid; e < np * T; e += nthr_cu) {
      const int j = e / T, t = e - j * T;
      double v = 0.0;
      if (j < n && t < nn) {
        const int st = S.status[t];
        if (st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED) v = __ldg(I.Dinv + j) * W.gsx[(size_t)t * np + j];
      }
      L.xs[vidx<T>(j, t)] = v;
    }
    named_bar(1, nthr_all);                            // with the producer: epilogue operands ready
  };

  if (is_update) {
    // ============================================================= update warps
    Updater<T, CS> U;
    U.L = L; U.lane = lane; U.uw = warp - nwslots; U.g = 0; U.gb = 0; U.cls = 0; U.gph = 0; U.active = lane < kPR * T;
    uint32_t ck_ph = 0;
    RowAcc R;
    auto reset = [&]() { R.pr = R.a1 = R.a2 = R.ndy = R.lhs = R.quad = R.lin = 0.0; R.vu = -INFINITY; R.vl = INFINITY; };
    // this warp's row-space accumulators -> finp (one value per node)
    auto publish_rows = [&](int cls_sum) {
      const double v[kFinP] = {reduce_same_node<T, 0>(R.pr), reduce_same_node<T, 0>(R.a1), reduce_same_node<T, 0>(R.a2),
                               reduce_same_node<T, 0>(R.ndy), reduce_same_node<T, 1>(R.lhs), reduce_same_node<T, 0>(R.vu),
                               reduce_same_node<T, 2>(R.vl), reduce_same_node<T, 1>(R.quad), reduce_same_node<T, 1>(R.lin)};
      if (lane < T) {   // maxima / minima: any order; sums (q = 4 lhs, 7 quad, 8 lin): by the class of the pass that made them
#pragma unroll
        for (int q = 0; q < kFinP; q++) S.finp[(q == 4 || q == 7 || q == 8) ? cls_sum : U.uw][q][lane] = v[q];
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

  // =============================================================== consumer warps
  Consumer<T, LAG, CS> C;
  C.L = L; C.cw = warp; C.lane = lane; C.rg = lane >> 3; C.cg = lane & 7;
  C.slot = 0; C.slot2 = 0; C.phase = 0; C.g = 0; C.gb = 0; C.ud_g = 0; C.ud_b = 0; C.ud_ph = 0; C.up_g = 0; C.up_b = 0; C.up_ph = 0;
  C.writer = T >= 4 ? true : (T == 2 ? (lane & 4) == 0 : (lane & 6) == 0);
  const int wg = w0 + warp;      // this warp's column tile
  double acc[4][T];
  // b' = sigma x - q + A'(rho z - y) for this warp's columns, from the pass-2 accumulators, into vs (read back by the
  // same warp only: the M pass input)
  auto finalize_b = [&]() {
    C.allreduce_rg(acc);
    if (C.rg == 0) {
      const int c0 = C.col0();
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int j = c0 + b;
        const double qj = j < n ? __ldg(I.q + j) : 0.0;
#pragma unroll
        for (int t = 0; t < T; t++) L.vs[vidx<T>(j, t)] = j < n ? sigma * L.xs[vidx<T>(j, t)] - qj + acc[b][t] : 0.0;
      }
    }
    __syncwarp();
  };
  // this warp's columns of a workspace vector staged into xts (x~ is dead by then)
  auto stage_cols = [&](const double *gvec) {
    if (C.rg == 0) {
      const int c0 = C.col0();
#pragma unroll
      for (int b = 0; b < 4; b++)
#pragma unroll
        for (int t = 0; t < T; t++) L.xts[vidx<T>(c0 + b, t)] = gvec[(size_t)(c0 + b) * T + t];
    }
    __syncwarp();
  };
  auto store_cols = [&](double *gvec) {
    if (C.rg == 0) {
      const size_t o = (size_t)C.col0() * T;
#pragma unroll
      for (int b = 0; b < 4; b++)
#pragma unroll
        for (int t = 0; t < T; t++) gvec[o + b * T + t] = acc[b][t];
    }
    __syncwarp();
  };
  zero4<T>(acc);
  C.template pass<true>(npa, L.xs, acc);            // z = A x0 (first round) ; A'(rho z - y)
  finalize_b();
  int iter;
  for (iter = iter_begin + 1; iter <= iter_end; iter++) {
    const bool do_check = (iter % check_every == 0) || iter == max_iter;
    C.template pass<false>(npm, L.vs, acc);         // x~ = M b  (acc untouched)
    C.wait_ud(C.g - 1);                             // every row of x~ is in xts
    zero4<T>(acc);
    C.template pass<true>(npa, L.xts, acc);         // z~ = A x~ ; b' += A' w
    finalize_b();
    if (!do_check) continue;

    // ---- termination check (update_info + check_termination): A x, A'y | A dx, A'dy | P x, P dx   (b' stays in vs)
    zero4<T>(acc);
    C.template pass<true>(npa, L.xs, acc);
    C.allreduce_rg(acc); store_cols(W.gaty);
    stage_cols(W.gdx); zero4<T>(acc);
    C.template pass<true>(npa, L.xts, acc);
    C.allreduce_rg(acc); store_cols(W.gatd);
    zero4<T>(acc);
    C.template pass<true>(npm, L.xs, acc);
    C.allreduce_rg(acc); store_cols(W.gpdx);
    {
      // column-space quantities: lane <-> column 32 * (column tile) + lane (the same partition for every tile width).
      // Everything read here was written by this warp (store_cols) or by update warps whose panels it has waited for.
      const int j = wg * 32 + lane;
      const bool inr = j < n;
      const double di = inr ? __ldg(I.Dinv + j) : 0.0, dj = inr ? __ldg(I.D + j) : 0.0, qj = inr ? __ldg(I.q + j) : 0.0;
      auto put = [&](int q, int t, double v) {
        if (lane == 0) {
          const int ri = (q * NW + wg) * T + t;
          L.red[ri] = v;
          if constexpr (CS == 2) st_async_remote_f64(L.red_r + 8u * ri, v, L.ck_r);
        }
      };
#pragma unroll
      for (int t = 0; t < T; t++) {
        const size_t e = (size_t)j * T + t;
        const double px = inr ? W.gpx[e] : 0.0, aty = inr ? W.gaty[e] : 0.0, xj = inr ? L.xs[vidx<T>(j, t)] : 0.0,
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
      if (lane == 0) {
        if (CS == 2 && warp == 0) mbar_expect_tx(L.ck, (uint32_t)((NW - NWc) * kColQ * T * 8));
        else mbar_arrive(L.ck);
      }
    }
    named_bar(2, nthr_cu);     // decision made
    snapshot();
    named_bar(1, nthr_all);
    if (S.remaining == 0 || iter == iter_end) break;
  }
  finish_common(iter);
  C.template pass<false>(npm, L.xs, acc);           // P x at the clipped point (sums taken by the update warps)
  C.wait_ud(C.g - 1); C.wait_udp(C.g - 1);   // the update warps of both CTAs are done with our partials: safe to leave the cluster
}

}  // namespace

size_t panel_smem_bytes(int npad, int tt, int nslots, int cs) {
  const int nw = npad / 32;
  const int nwc = cs == 2 ? (nw + 1) / 2 : nw;
  size_t off = (sizeof(PanelShared) + 15) & ~size_t(15);
  off += sizeof(uint64_t) * (2 * (size_t)nslots + 3 * (size_t)kHB + 1);
  off = (off + 15) & ~size_t(15);
  off += ((size_t)3 * npad * tt + (size_t)kHB * nw * kPR * tt + (size_t)kHB * kPR * tt + (size_t)kColQ * nw * tt) * 8;
  off = (off + 127) & ~size_t(127);
  return off + (size_t)nslots * nwc * (kPR * 32 * 8);
}

template <int T, int LAG, int CS>
static int launch_p(int nw_max, int nslots, double *d_state, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles,
                    const double *d_in, double *d_out, double *d_work, NodeScalars *d_ns, int *d_tile_iters, size_t smem,
                    cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(admm_panel_kernel<T, LAG, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return BQP_E_CUDA;
  int prefetch_panels = 0;   // L2 prefetch distance of the producer, in panels (experiment knob)
  if (const char *pk = getenv("BQP_PANEL_PREFETCH")) prefetch_panels = atoi(pk);
  const int nwc = CS == 2 ? (nw_max + 1) / 2 : nw_max;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ntiles * CS), 1, 1);
  cfg.blockDim = dim3((unsigned)((nwc + kPanelUpdWarps + 1) * 32), 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, admm_panel_kernel<T, LAG, CS>, d_insts, d_tiles, d_in, d_out, d_work, d_ns, d_tile_iters, nslots, d_state,
                         prefetch_panels);
  return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? BQP_OK : BQP_E_CUDA;
}

// cs = CTAs per tile (1, or 2 = a cluster pair splitting the columns).  Pass 2 runs two panels behind pass 1.
int launch_admm_panel(int tt, int cs, int nw_max, int nslots, double *d_state, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles,
                      const double *d_in, double *d_out, double *d_work, NodeScalars *d_ns, int *d_tile_iters,
                      size_t smem_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (nw_max < 1 || nw_max > kPanelMaxWarps || (cs != 1 && cs != 2) || (cs == 1 && nw_max > kPanelCtaWarps)) return BQP_E_ARG;
#define BQP_PANEL_CASE(TT, CC)                                                                                              \
  if (tt == TT && cs == CC)                                                                                                  \
    return launch_p<TT, 2, CC>(nw_max, nslots, d_state, d_insts, d_tiles, ntiles, d_in, d_out, d_work, d_ns, d_tile_iters, smem_bytes, st);
  BQP_PANEL_CASE(1, 1) BQP_PANEL_CASE(2, 1) BQP_PANEL_CASE(4, 1)
  BQP_PANEL_CASE(1, 2) BQP_PANEL_CASE(2, 2) BQP_PANEL_CASE(4, 2)
#undef BQP_PANEL_CASE
  return BQP_E_ARG;
}

}  // namespace bqp
