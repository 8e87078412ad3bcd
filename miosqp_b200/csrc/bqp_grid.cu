// bqp_grid.cu -- whole-GPU ADMM kernel for ONE large tile (sm_100a): BASELINE config 4 (random_miqp n = 2000, m = 4200,
// A 5 % dense), i.e. dense-reduced problems wider than the 512 columns the rows kernel (bqp_rows.cu) keeps in registers.
//
// The B&B of the reference offers two unsolved leaves per step (/root/reference/miosqp/workspace.py:282-350), so config 4
// is ONE tile of <= 8 leaves at a time and its cost is the latency of an ADMM iteration (node.py:96-143 per leaf).  The
// stream kernel runs that tile on one SM: 28.8 MB of LDL' factor through one SM's L2 port, 770 us per iteration.  Here
// every SM of the GPU works on the same tile.  Same restated iteration as the dense kernels,
//     x~ = M b            M = (P + sigma I + A' rho A)^-1 explicit (guarded at setup), 8-row fragment-ordered panels, L2 resident
//     z~ = A x~ , z / y update , w = rho z - y          A as CSR
//     b' = sigma x - q + A' w                           A' as CSR
// as three phases separated by grid-wide barriers:
//     M phase   CTA g owns the row panels p = g (mod G): FP64 mma.sync.m8n8k4 (the 8 leaves are the N dimension), the 16 warps
//               of the CTA split the column tiles, partial 8x8 blocks meet in shared memory; b staged once in shared memory
//     A phase   CTA g owns a contiguous block of rows of A_ext; z, y, l, u of those rows LIVE IN SHARED MEMORY for the whole
//               solve; one warp per sparse row, 8 lanes (= 8 leaves) per entry, x~ staged in shared memory; the same phase
//               updates x for the CTA's block of columns
//     A' phase  CTA g owns a contiguous block of columns; one warp per column, w gathered from L2
// Termination checks (every check_termination iterations): row owners form A x, A dx and their norms, panel owners P x, P dx,
// column owners A'y, A' dy; per-CTA partial norms go to global memory and EVERY CTA reduces them in the same canonical
// order -- identical decisions everywhere, nothing to broadcast.  All norms are maxima (order-free); the four sums
// (x'Px, q'x, the two certificate products) are added CTA by CTA in index order.
// ADAPTIVE RHO (osqp adapt_rho, settings adaptive_rho with a fixed interval): the M phase becomes x~ = V (d . (V' b)) with
// d = 1 / (1 + (rho_leaf - rho0) mu) -- the spectral form of K(rho)^-1 built at setup (bqp_setup.cpp build_grid) -- so every
// leaf of the tile carries its OWN rho and a rho update costs nothing but one extra A' pass; the estimate uses the same scaled
// norms as osqp's compute_rho_estimate, evaluated with the termination check the interval coincides with.
// Every cross-CTA vector is read with ld.global.cg (L2); every wait is bounded and traps instead of hanging the GPU.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>

#include "bqp_internal.h"

namespace bqp {

namespace {

constexpr int T8 = 8;
constexpr int kGW = 16;                    // warps per CTA
constexpr int kGT = kGW * 32;              // threads per CTA
constexpr int kFinN = 24;                 // 16 quantities of the termination tests + 7 scaled norms of osqp's compute_rho_estimate
constexpr int kTileD = 256;                // doubles of one column tile of one panel (8 rows x 32 columns)

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct GridShared {
  DevInstance I;
  DevTile tile;
  double fin[kFinN][T8];
  int status[T8], iters[T8], newly[T8];
  int remaining;
  double rho_t[T8];                          // adaptive rho: the leaf's current rho (osqp settings->rho after osqp_update_rho)
  int rho_changed;
};

// grid-wide barrier: a monotone arrival counter in global memory (reset by the host before the launch); CTA-level
// bar.sync + one release/acquire pair by thread 0, as cooperative groups does, with a bounded spin
__device__ __forceinline__ void grid_sync(unsigned *ctr, unsigned &target, unsigned G) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += G;
    // fire-and-forget arrival (no round trip before the polling starts); release orders this CTA's writes (cumulative over bar.sync)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    const long long t0 = clock64();
    while (ld_acquire_u32(ctr) < target) {
      if (clock64() - t0 > 8000000000LL) __trap();     // ~4 s
    }
  }
  __syncthreads();
}

// shared-memory carve-up (doubles unless noted), rpc / cpc = rows / columns a CTA owns
struct GCarve { size_t vec, part, red, z, y, l, u, dy, ax, x, dx, sx, total; };
__host__ __device__ inline GCarve grid_carve(int npad, int rpc, int cpc) {
  GCarve c;
  size_t off = (sizeof(GridShared) + 127) & ~size_t(127);
  c.vec = off; off += (size_t)npad * T8 * 8;
  c.part = off; off += (size_t)kGW * 64 * 8;
  c.red = off; off += (size_t)kFinN * kGW * T8 * 8;
  c.z = off; off += (size_t)rpc * T8 * 8; c.y = off; off += (size_t)rpc * T8 * 8;
  c.l = off; off += (size_t)rpc * T8 * 8; c.u = off; off += (size_t)rpc * T8 * 8;
  c.dy = off; off += (size_t)rpc * T8 * 8; c.ax = off; off += (size_t)rpc * T8 * 8;
  c.x = off; off += (size_t)cpc * T8 * 8; c.dx = off; off += (size_t)cpc * T8 * 8; c.sx = off; off += (size_t)cpc * T8 * 8;
  c.total = off;
  return c;
}

// one warp: sparse row [beg, end) times a [column][8] vector; every lane returns the sum for leaf (lane & 7).
// 8 lanes per entry, 4 entries per step; the (value, column) pairs of 32 entries are loaded coalesced and broadcast by shuffle
template <bool kSmem>
__device__ __forceinline__ double sparse_row(const double *__restrict__ val, const int *__restrict__ col, int beg, int end,
                                             const double *vec, int lane) {
  const int t = lane & 7, sub = lane >> 3;
  double acc = 0.0;
  for (int base = beg; base < end; base += 32) {
    const int e = base + lane;
    double v = 0.0; int c = 0;
    if (e < end) { v = __ldg(val + e); c = __ldg(col + e); }
    double xv[8], av[8];
#pragma unroll
    for (int s = 0; s < 8; s++) {
      const int src = 4 * s + sub;
      av[s] = __shfl_sync(0xffffffffu, v, src);
      const int cc = __shfl_sync(0xffffffffu, c, src);
      xv[s] = kSmem ? vec[(size_t)cc * T8 + t] : __ldcg(vec + (size_t)cc * T8 + t);
    }
#pragma unroll
    for (int s = 0; s < 8; s++) acc = fma(av[s], xv[s], acc);      // padding entries: 0 * vec[0] (finite)
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 8);
  acc += __shfl_xor_sync(0xffffffffu, acc, 16);
  return acc;
}

__global__ void __launch_bounds__(kGT, 1)
admm_grid_kernel(const DevInstance *__restrict__ insts, const DevTile *__restrict__ tiles, int ntiles, const double *__restrict__ in,
                 double *__restrict__ out, double *__restrict__ work, NodeScalars *__restrict__ ns, int *__restrict__ tile_iters,
                 unsigned *__restrict__ barrier) {
  constexpr int T = T8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GridShared &S = *reinterpret_cast<GridShared *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  const int gq = lane >> 2, tq = lane & 3;
  unsigned bar_target = 0;

  for (int ti = 0; ti < ntiles; ti++) {
    __syncthreads();
    if (tid == 0) {
      S.tile = tiles[ti];
      S.I = insts[S.tile.inst];
      S.remaining = S.tile.nn;
    }
    if (tid < T) { S.status[tid] = BQP_UNSOLVED; S.iters[tid] = 0; S.newly[tid] = 0; }
    __syncthreads();
    if (tid < T) S.rho_t[tid] = S.I.rho_base;
    if (tid == 0) S.rho_changed = 0;
    __syncthreads();
    const DevInstance &I = S.I;
    const int n = I.n, m = I.m, np = I.npad, nn = S.tile.nn, NW = np / 32, npm = I.g_npm;
    const int m8 = (m + 7) / 8 * 8;
    const int rpc = (m + G - 1) / G, cpc = (np + G - 1) / G;
    const int r0 = min(m, cta * rpc), nr = min(m, r0 + rpc) - r0;
    const int c0 = min(np, cta * cpc), nc = min(np, c0 + cpc) - c0;
    const GCarve cv = grid_carve(np, rpc, cpc);
    double *vec = reinterpret_cast<double *>(smem_raw + cv.vec), *part = reinterpret_cast<double *>(smem_raw + cv.part);
    double *red = reinterpret_cast<double *>(smem_raw + cv.red);
    double *sz = reinterpret_cast<double *>(smem_raw + cv.z), *sy = reinterpret_cast<double *>(smem_raw + cv.y);
    double *sl = reinterpret_cast<double *>(smem_raw + cv.l), *su = reinterpret_cast<double *>(smem_raw + cv.u);
    double *sdy = reinterpret_cast<double *>(smem_raw + cv.dy), *sax = reinterpret_cast<double *>(smem_raw + cv.ax);
    double *sx = reinterpret_cast<double *>(smem_raw + cv.x), *sdx = reinterpret_cast<double *>(smem_raw + cv.dx);
    double *ssx = reinterpret_cast<double *>(smem_raw + cv.sx);
    // per-tile workspace in global memory (L2 resident), [row][8 leaves]
    double *gb = work + S.tile.work_off, *gxt = gb + (size_t)np * T, *gx = gxt + (size_t)np * T, *gdx = gx + (size_t)np * T;
    double *gpx = gdx + (size_t)np * T, *gpdx = gpx + (size_t)np * T, *gxo = gpdx + (size_t)np * T;
    double *gc = gxo + (size_t)np * T;       // adaptive rho: d . (V' b)
    double *gw = gc + (size_t)np * T, *gy = gw + (size_t)m8 * T, *gdyp = gy + (size_t)m8 * T, *gpart = gdyp + (size_t)m8 * T;
    const bool adaptive = I.adaptive != 0;
    // rho of row i for leaf t: typed once at setup; with adaptive rho the inequality / equality rows follow the leaf's rho
    auto row_rho = [&](int i, int t) -> double {
      if (!adaptive) return __ldg(I.rho + i);
      const int ty = __ldg(I.g_rtype + i);
      return ty == 0 ? S.rho_t[t] : (ty == 1 ? kRhoEqFactor * S.rho_t[t] : kRhoMin);
    };
    const int max_iter = I.max_iter, check_every = I.check_every;
    const double alpha = I.alpha, oma = 1.0 - I.alpha, sigma = I.sigma;

    // vec (shared) <- a [npad][8] vector in global memory
    auto stage = [&](const double *src) {
      const double2 *s2 = reinterpret_cast<const double2 *>(src);
      double2 *d2 = reinterpret_cast<double2 *>(vec);
      for (int e = tid; e < np * T / 2; e += kGT) d2[e] = __ldcg(s2 + e);
      __syncthreads();
    };
    // rows of (panel matrix) x vec for the panels this CTA owns: f(row, leaf pair index tq', value0, value1) by 32 threads
    auto dense_pass = [&](const double *Mat, auto &&f) {
      for (int p = cta; p < npm; p += G) {
        const double *pp = Mat + (size_t)p * 8 * np + lane;
        double cc[4][2];
#pragma unroll
        for (int k = 0; k < 4; k++) cc[k][0] = cc[k][1] = 0.0;
#pragma unroll 2
        for (int ct = warp; ct < NW; ct += kGW) {
          double a[8], b[8];
#pragma unroll
          for (int ks = 0; ks < 8; ks++) a[ks] = __ldcg(pp + (size_t)ct * kTileD + ks * 32);
#pragma unroll
          for (int ks = 0; ks < 8; ks++) b[ks] = vec[(size_t)(32 * ct + 4 * ks + tq) * T + gq];
#pragma unroll
          for (int ks = 0; ks < 8; ks++) dmma(cc[ks & 3], a[ks], b[ks]);
        }
        double2 *pb = reinterpret_cast<double2 *>(part);
        pb[warp * 32 + lane] = make_double2((cc[0][0] + cc[1][0]) + (cc[2][0] + cc[3][0]), (cc[0][1] + cc[1][1]) + (cc[2][1] + cc[3][1]));
        __syncthreads();
        if (warp == 0) {
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int w = 0; w < kGW; w++) { const double2 v = pb[w * 32 + lane]; s0 += v.x; s1 += v.y; }     // warps in order
          f(p * 8 + gq, tq, s0, s1);       // C fragment: row lane >> 2, leaves 2 (lane & 3) + {0, 1}
        }
        __syncthreads();
      }
    };
    auto st2 = [](double *v, int row, int pair, double a, double b) { reinterpret_cast<double2 *>(v)[(size_t)row * (T8 / 2) + pair] = make_double2(a, b); };

    // ---- prologue (node.py:102-105): scaled bounds and duals of the CTA's rows, scaled warm start of its columns
    for (int e = tid; e < nr * T; e += kGT) {
      const int i = r0 + e / T, t = e % T;
      double lo = -kInfty, up = kInfty, yv = 0.0;
      if (t < nn) {
        const double *p = in + S.tile.in_off[t];
        lo = fmax(p[i], -kInfty); up = fmin(p[m + i], kInfty);
        yv = I.c * __ldg(I.Einv + i) * p[2 * (size_t)m + n + i];
      }
      const double ei = __ldg(I.E + i);
      sl[e] = ei * lo; su[e] = ei * up; sy[e] = yv; sz[e] = 0.0; sdy[e] = 0.0;
    }
    for (int e = tid; e < nc * T; e += kGT) {
      const int j = c0 + e / T, t = e % T;
      double xv = 0.0;
      if (j < n && t < nn) xv = __ldg(I.Dinv + j) * in[S.tile.in_off[t] + 2 * (size_t)m + j];
      sx[e] = xv; sdx[e] = 0.0; ssx[e] = 0.0;
      gx[(size_t)j * T + t] = xv;
    }
    grid_sync(barrier, bar_target, (unsigned)G);

    // row phase.  MODE 0: z = A x0 (warm start), 1: ADMM update from x~, 2: check products A x / A dx (no state change)
    // column phase: s = A' w for the CTA's columns, then b' = sigma x - q + s
    auto col_phase = [&](bool publish) {
      for (int jl = warp; jl < nc; jl += kGW) {
        const int j = c0 + jl;
        double s = 0.0;
        if (j < n) s = sparse_row<false>(I.g_tvl, I.g_tci, __ldg(I.g_trp + j), __ldg(I.g_trp + j + 1), gw, lane);
        if (lane < T) {
          const int t = lane;
          double b = 0.0;
          if (j < n) b = sigma * sx[jl * T + t] - __ldg(I.q + j) + s;
          gb[(size_t)j * T + t] = b;
          if (publish) { gx[(size_t)j * T + t] = sx[jl * T + t]; gdx[(size_t)j * T + t] = sdx[jl * T + t]; }
        }
      }
    };

    // ---- start: z = A x0, w = rho z - y; b of the starting point
    stage(gx);
    for (int il = warp; il < nr; il += kGW) {
      const int i = r0 + il;
      const double s = sparse_row<true>(I.g_avl, I.g_aci, __ldg(I.g_arp + i), __ldg(I.g_arp + i + 1), vec, lane);
      if (lane < T) {
        sz[il * T + lane] = s;
        gw[(size_t)i * T + lane] = fma(row_rho(i, lane), s, -sy[il * T + lane]);
      }
    }
    grid_sync(barrier, bar_target, (unsigned)G);
    col_phase(false);
    grid_sync(barrier, bar_target, (unsigned)G);

#ifdef BQP_GRID_DEBUG
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pl = clock64();
#define GSTAMP(i) do { const long long now_ = clock64(); ph[i] += now_ - pl; pl = now_; } while (0)
#else
#define GSTAMP(i) do { } while (0)
#endif
    int iter;
    for (iter = 1; iter <= max_iter; iter++) {
      const bool do_check = (iter % check_every == 0) || iter == max_iter;
      // ---- M phase: x~ = M b
      GSTAMP(7);
      stage(gb);
      GSTAMP(0);
      if (!adaptive) {
        dense_pass(I.g_M, [&](int row, int pair, double a, double b) { st2(gxt, row, pair, a, b); });
      } else {
        // x~ = V (d . (V' b)),  d = 1 / (1 + (rho_leaf - rho0) mu_row)
        dense_pass(I.g_M, [&](int row, int pair, double a, double b) {
          const double mu = __ldg(I.g_mu + row);
          st2(gc, row, pair, a / (1.0 + (S.rho_t[2 * pair] - I.rho_base) * mu), b / (1.0 + (S.rho_t[2 * pair + 1] - I.rho_base) * mu));
        });
        grid_sync(barrier, bar_target, (unsigned)G);
        stage(gc);
        dense_pass(I.g_V, [&](int row, int pair, double a, double b) { st2(gxt, row, pair, a, b); });
      }
      GSTAMP(1);
      grid_sync(barrier, bar_target, (unsigned)G);
      GSTAMP(2);
      // ---- A phase: z~ = A x~, projection, dual update; x update of the CTA's columns
      stage(gxt);
      GSTAMP(0);
      for (int il = warp; il < nr; il += kGW) {
        const int i = r0 + il;
        const double zt = sparse_row<true>(I.g_avl, I.g_aci, __ldg(I.g_arp + i), __ldg(I.g_arp + i + 1), vec, lane);
        if (lane < T) {
          const int e = il * T + lane;
          const double rho = row_rho(i, lane), rinv = adaptive ? 1.0 / rho : __ldg(I.rho_inv + i);
          const double zr = alpha * zt + oma * sz[e], yo = sy[e];
          double z = zr + rinv * yo;
          z = fmin(fmax(z, sl[e]), su[e]);
          const double dy = rho * (zr - z), yn = yo + dy;
          sz[e] = z; sy[e] = yn; sdy[e] = dy;
          gw[(size_t)i * T + lane] = fma(rho, z, -yn);
          if (do_check) {
            gy[(size_t)i * T + lane] = yn;
            double d = dy;                                     // dy projected on the recession directions of [l, u]
            if (su[e] > kInfty * kMinScaling) {
              if (sl[e] < -kInfty * kMinScaling) d = 0.0; else d = fmin(d, 0.0);
            } else if (sl[e] < -kInfty * kMinScaling) d = fmax(d, 0.0);
            gdyp[(size_t)i * T + lane] = d;
          }
        }
      }
      for (int e = tid; e < nc * T; e += kGT) {
        const int j = c0 + e / T;
        const double xp = sx[e], xn = alpha * vec[(size_t)j * T + (e % T)] + oma * xp;
        sx[e] = xn; sdx[e] = xn - xp;
      }
      GSTAMP(3);
      grid_sync(barrier, bar_target, (unsigned)G);
      GSTAMP(4);
      // ---- A' phase: b' = sigma x - q + A' w
      col_phase(do_check);
      GSTAMP(5);
      grid_sync(barrier, bar_target, (unsigned)G);
      GSTAMP(6);
      if (!do_check) continue;

      // ---- termination check (update_info + check_termination)
      // rows: A x, A dx against the CTA's z, dy, l, u (lane < 8 holds leaf `lane` of the warp's rows)
      stage(gx);
      dense_pass(I.g_P, [&](int row, int pair, double a, double b) { st2(gpx, row, pair, a, b); });
      for (int il = warp; il < nr; il += kGW) {
        const double ax = sparse_row<true>(I.g_avl, I.g_aci, __ldg(I.g_arp + r0 + il), __ldg(I.g_arp + r0 + il + 1), vec, lane);
        if (lane < T) sax[il * T + lane] = ax;
      }
      __syncthreads();
      stage(gdx);
      dense_pass(I.g_P, [&](int row, int pair, double a, double b) { st2(gpdx, row, pair, a, b); });
      double v[kFinN];
      v[0] = v[1] = v[2] = v[5] = v[6] = v[7] = v[8] = v[10] = v[12] = v[13] = 0.0;
      v[3] = v[4] = v[9] = v[11] = 0.0; v[14] = -INFINITY; v[15] = INFINITY;
#pragma unroll
      for (int q = 16; q < kFinN; q++) v[q] = 0.0;          // scaled norms of compute_rho_estimate: |Ax - z|, |z|, |Ax|, |Px + q + A'y|, |q|, |A'y|, |Px|
      {
        for (int il = warp; il < nr; il += kGW) {
          const int i = r0 + il;
          const double adx = sparse_row<true>(I.g_avl, I.g_aci, __ldg(I.g_arp + i), __ldg(I.g_arp + i + 1), vec, lane);
          if (lane < T) {
            const int e = il * T + lane;
            const double ei = __ldg(I.Einv + i), Ei = __ldg(I.E + i), z = sz[e], lo = sl[e], up = su[e], ax = sax[e];
            v[5] = fmax(v[5], fabs(ei * (ax - z)));
            v[6] = fmax(v[6], fabs(ei * ax));
            v[7] = fmax(v[7], fabs(ei * z));
            v[16] = fmax(v[16], fabs(ax - z)); v[17] = fmax(v[17], fabs(z)); v[18] = fmax(v[18], fabs(ax));
            const double w = ei * adx;
            if (up < kInfty * kMinScaling) v[14] = fmax(v[14], w);
            if (lo > -kInfty * kMinScaling) v[15] = fmin(v[15], w);
            double d = sdy[e];
            if (up > kInfty * kMinScaling) {
              if (lo < -kInfty * kMinScaling) d = 0.0; else d = fmin(d, 0.0);
            } else if (lo < -kInfty * kMinScaling) d = fmax(d, 0.0);
            v[8] = fmax(v[8], fabs(Ei * d));
            v[9] += up * fmax(d, 0.0) + lo * fmin(d, 0.0);
          }
        }
      }
      grid_sync(barrier, bar_target, (unsigned)G);            // P x, P dx, y, projected dy of every CTA in global memory
      // columns: A'y, A' dy_proj, and the column-space norms
      for (int jl = warp; jl < nc; jl += kGW) {
        const int j = c0 + jl;
        if (j >= n) continue;
        const int beg = __ldg(I.g_trp + j), end = __ldg(I.g_trp + j + 1);
        const double aty = sparse_row<false>(I.g_tvl, I.g_tci, beg, end, gy, lane);
        const double atd = sparse_row<false>(I.g_tvl, I.g_tci, beg, end, gdyp, lane);
        if (lane < T) {
          const size_t e = (size_t)j * T + lane;
          const double px = __ldcg(gpx + e), pdx = __ldcg(gpdx + e), xj = sx[jl * T + lane], dxj = sdx[jl * T + lane];
          const double di = __ldg(I.Dinv + j), dj = __ldg(I.D + j), qj = __ldg(I.q + j);
          v[0] = fmax(v[0], fabs(di * (px + qj + aty)));
          v[1] = fmax(v[1], fabs(di * px));
          v[2] = fmax(v[2], fabs(di * aty));
          v[19] = fmax(v[19], fabs(px + qj + aty)); v[20] = fmax(v[20], fabs(qj)); v[21] = fmax(v[21], fabs(aty)); v[22] = fmax(v[22], fabs(px));
          v[3] += xj * px;
          v[4] += qj * xj;
          v[12] = fmax(v[12], fabs(di * atd));
          v[13] = fmax(v[13], fabs(di * pdx));
          v[10] = fmax(v[10], fabs(dj * dxj));
          v[11] += qj * dxj;
        }
      }
      // CTA partials: lanes < 8 of every warp hold leaf `lane`; warps in order; then to global, [cta][16][8]
      const int op[kFinN] = {0, 0, 0, 1, 1, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0};
      if (lane < T)
#pragma unroll
        for (int q = 0; q < kFinN; q++) red[(q * kGW + warp) * T + lane] = v[q];
      __syncthreads();
      if (tid < kFinN * T) {
        const int q = tid >> 3, t = tid & 7;
        double x = red[(q * kGW) * T + t];
        for (int w = 1; w < kGW; w++) {
          const double y = red[(q * kGW + w) * T + t];
          x = op[q] == 0 ? fmax(x, y) : (op[q] == 1 ? x + y : fmin(x, y));
        }
        gpart[((size_t)cta * kFinN + q) * T + t] = x;
      }
      grid_sync(barrier, bar_target, (unsigned)G);
      if (tid < kFinN * T) {
        const int q = tid >> 3, t = tid & 7;
        double x = __ldcg(gpart + (size_t)q * T + t);
        for (int c = 1; c < G; c++) {
          const double y = __ldcg(gpart + ((size_t)c * kFinN + q) * T + t);
          x = op[q] == 0 ? fmax(x, y) : (op[q] == 1 ? x + y : fmin(x, y));
        }
        S.fin[q][t] = x;
      }
      __syncthreads();
      if (tid < T) {
        // scalar decision (optimality / infeasibility tests of OSQP), every CTA identically
        const int t = tid;
        S.newly[t] = 0;
        if (t < nn && S.status[t] == BQP_UNSOLVED) {
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
            if (cta == 0) {
              NodeScalars r;
              r.status = status; r.iters = iter; r.pri_res = pri; r.dua_res = dua;
              r.obj = (status == BQP_PRIMAL_INFEASIBLE || status == BQP_PRIMAL_INFEASIBLE_INACCURATE) ? kInfty
                      : (status == BQP_DUAL_INFEASIBLE || status == BQP_DUAL_INFEASIBLE_INACCURATE) ? -kInfty
                      : (status == BQP_NON_CVX ? NAN : obj);
              r.lower = NAN;
              ns[S.tile.node[t]] = r;
            }
            atomicSub(&S.remaining, 1);
          }
        }
      }
      __syncthreads();
      // unscaled iterates of the leaves that terminated at this check: the CTA's columns of x and rows of y
      for (int t = 0; t < nn; t++) {
        if (!S.newly[t]) continue;
        const int st = S.status[t];
        const bool bad = !(st == BQP_SOLVED || st == BQP_SOLVED_INACCURATE || st == BQP_MAX_ITER_REACHED);
        double *ox = out + S.tile.out_off[t], *oy = ox + n;
        for (int jl = tid; jl < nc; jl += kGT) {
          const int j = c0 + jl;
          if (j < n) { const double val = bad ? NAN : __ldg(I.D + j) * sx[jl * T + t]; ssx[jl * T + t] = val; ox[j] = val; }
        }
        for (int il = tid; il < nr; il += kGT) oy[r0 + il] = bad ? NAN : I.cinv * __ldg(I.E + r0 + il) * sy[il * T + t];
      }
      __syncthreads();
      if (S.remaining == 0) break;
      if (adaptive && iter % I.adapt_interval == 0) {
        // osqp adapt_rho: rho_new = rho sqrt(pri / dua) on the normalised SCALED residuals; adopted outside [rho / tol, rho tol]
        if (tid < T && tid < nn && S.status[tid] == BQP_UNSOLVED) {
          const int t = tid;
          const double pri = S.fin[16][t] / (fmax(S.fin[17][t], S.fin[18][t]) + 1e-10);
          const double dua = S.fin[19][t] / (fmax(fmax(S.fin[20][t], S.fin[21][t]), S.fin[22][t]) + 1e-10);
          const double rho = S.rho_t[t];
          double rn = rho * sqrt(pri / (dua + 1e-10));
          rn = fmin(fmax(rn, kRhoMin), 1e6);
          if (rn > rho * I.adapt_tol || rn < rho / I.adapt_tol) { S.rho_t[t] = rn; S.rho_changed = 1; }
        }
        __syncthreads();
        if (S.rho_changed) {
          // the right-hand side of the next iteration carries rho: w = rho z - y and b' = sigma x - q + A' w again
          for (int e = tid; e < nr * T; e += kGT) {
            const int il = e / T, t = e % T;
            gw[(size_t)(r0 + il) * T + t] = fma(row_rho(r0 + il, t), sz[e], -sy[e]);
          }
          grid_sync(barrier, bar_target, (unsigned)G);
          col_phase(false);
          grid_sync(barrier, bar_target, (unsigned)G);
          if (tid == 0) S.rho_changed = 0;
          __syncthreads();
        }
      }
    }

#ifdef BQP_GRID_DEBUG
    if (tid == 0 && (cta == 0 || cta == G / 2 || cta == G - 1)) {
      const double k = 1.0 / (iter > max_iter ? max_iter : iter);
      printf("GRID cta %d per iteration (clk): stage %.0f  M %.0f  sync1 %.0f  A+x %.0f  sync2 %.0f  A' %.0f  sync3 %.0f  checks etc %.0f\n", cta,
             ph[0] * k, ph[1] * k, ph[2] * k, ph[3] * k, ph[4] * k, ph[5] * k, ph[6] * k, ph[7] * k);
    }
#endif
    // ---- clip the integer variables to the leaf's bounds (node.py:128-143) and evaluate the objective at the clipped point
    for (int k = tid; k < I.n_int; k += kGT) {
      const int j = __ldg(I.i_idx + k), row = m - I.n_int + k;
      if (j < c0 || j >= c0 + nc) continue;
      for (int t = 0; t < nn; t++) {
        const int st = S.status[t];
        if (!(st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED)) continue;
        const double *p = in + S.tile.in_off[t];
        const double val = fmin(fmax(ssx[(j - c0) * T + t], p[row]), p[m + row]);
        ssx[(j - c0) * T + t] = val;
        out[S.tile.out_off[t] + j] = val;
      }
    }
    __syncthreads();
    for (int e = tid; e < nc * T; e += kGT) {
      const int j = c0 + e / T, t = e % T;
      double val = 0.0;
      if (j < n && t < nn) {
        const int st = S.status[t];
        if (st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED) val = __ldg(I.Dinv + j) * ssx[e];
      }
      sdx[e] = val;                                           // scaled clipped point of the CTA's columns
      gxo[(size_t)j * T + t] = val;
    }
    grid_sync(barrier, bar_target, (unsigned)G);
    stage(gxo);
    dense_pass(I.g_P, [&](int row, int pair, double a, double b) { st2(gpx, row, pair, a, b); });
    grid_sync(barrier, bar_target, (unsigned)G);
    {
      double v0 = 0.0, v1 = 0.0;                              // thread <-> (column tid >> 3 (+ 64 k), leaf tid & 7)
      const int t = tid & 7;
      for (int jl = tid >> 3; jl < nc; jl += kGT / 8) {
        const int j = c0 + jl;
        if (j < n) { const double xo = sdx[jl * T + t]; v0 += xo * __ldcg(gpx + (size_t)j * T + t); v1 += __ldg(I.q + j) * xo; }
      }
      // lanes with the same leaf: xor 8, 16; then the warps in order
      v0 += __shfl_xor_sync(0xffffffffu, v0, 8); v0 += __shfl_xor_sync(0xffffffffu, v0, 16);
      v1 += __shfl_xor_sync(0xffffffffu, v1, 8); v1 += __shfl_xor_sync(0xffffffffu, v1, 16);
      if (lane < T) { red[(0 * kGW + warp) * T + lane] = v0; red[(1 * kGW + warp) * T + lane] = v1; }
      __syncthreads();
      if (tid < 2 * T) {
        const int q = tid >> 3, tt = tid & 7;
        double x = red[(q * kGW) * T + tt];
        for (int w = 1; w < kGW; w++) x += red[(q * kGW + w) * T + tt];
        gpart[((size_t)cta * kFinN + q) * T + tt] = x;
      }
    }
    grid_sync(barrier, bar_target, (unsigned)G);
    if (cta == 0) {
      if (tid < 2 * T) {
        const int q = tid >> 3, tt = tid & 7;
        double x = __ldcg(gpart + (size_t)q * T + tt);
        for (int c = 1; c < G; c++) x += __ldcg(gpart + ((size_t)c * kFinN + q) * T + tt);
        S.fin[q][tt] = x;
      }
      __syncthreads();
      if (tid < nn) {
        const int st = S.status[tid];
        if (st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED) ns[S.tile.node[tid]].lower = (0.5 * S.fin[0][tid] + S.fin[1][tid]) * I.cinv;
      }
      if (tid == 0) tile_iters[ti] = iter > max_iter ? max_iter : iter;
    }
  }
}

}  // namespace

size_t grid_smem_bytes(int npad, int m, int n, int nctas) {
  (void)n;
  const int rpc = (m + nctas - 1) / nctas, cpc = (npad + nctas - 1) / nctas;
  return grid_carve(npad, rpc, cpc).total;
}

int grid_max_ctas(int device, size_t smem_bytes) {
  int sms = 0, per_sm = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
  if (cudaFuncSetAttribute(admm_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, admm_grid_kernel, kGT, smem_bytes) != cudaSuccess) return 0;
  return per_sm > 0 ? sms : 0;      // one CTA per SM
}

int launch_admm_grid(int nctas, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles, const double *d_in, double *d_out,
                     double *d_work, NodeScalars *d_ns, int *d_tile_iters, unsigned *d_barrier, size_t smem_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (nctas < 1 || ntiles < 1) return BQP_E_ARG;
  static std::atomic<int> attr_set{0};
  if (!attr_set.load()) {
    if (cudaFuncSetAttribute(admm_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem) != cudaSuccess) return BQP_E_CUDA;
    attr_set.store(1);
  }
  if (cudaMemsetAsync(d_barrier, 0, sizeof(unsigned), st) != cudaSuccess) return BQP_E_CUDA;
  void *args[] = {(void *)&d_insts, (void *)&d_tiles, (void *)&ntiles, (void *)&d_in, (void *)&d_out, (void *)&d_work, (void *)&d_ns,
                  (void *)&d_tile_iters, (void *)&d_barrier};
  // cooperative launch: every CTA of the grid is resident at once (the grid barrier needs it)
  cudaError_t e = cudaLaunchCooperativeKernel((const void *)admm_grid_kernel, dim3((unsigned)nctas), dim3((unsigned)kGT), args, smem_bytes, st);
  return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? BQP_OK : BQP_E_CUDA;
}

}  // namespace bqp
