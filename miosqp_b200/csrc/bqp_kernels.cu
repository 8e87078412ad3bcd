// bqp_kernels.cu -- the batched ADMM kernel for sm_100a.
//
// One CTA owns one TILE = up to T nodes (B&B leaves) of one set-up problem and runs their whole OSQP
// ADMM loop (what osqp.solve() does for /root/reference/miosqp/node.py:108, plus update/warm_start of
// node.py:102-105 and the clip+objective of node.py:128-143) without returning to the host:
//
//   prologue   l,u -> E*clamp(l,u); x = Dinv x0; y = c Einv y0; z = A x            (osqp update_bounds/warm_start)
//   iteration  b  = sigma x - q + A'(rho z - y)                                     (SpMV A', fused rhs)
//              b <- L22^-T D2^-1 L22^-1 b   (blocked dense triangular sweeps, node axis innermost)
//              x  = alpha b + (1-alpha) x ; zt = A b ; z = clip(alpha zt+(1-alpha) z + y/rho) ; y += rho(..)
//                                                                                   (SpMV A, fused projection + dual update)
//   every check_termination iterations: A x, P x, A' y, A' dy, P dx, A dx, inf-norms by warp shuffles,
//              OSQP's optimality / primal / dual infeasibility tests per node.
//   epilogue   x = D x, y = E y / c, clip integer entries, lower = 1/2 x'Px + q'x.
//
// The KKT solve is the LDL^T of [[P+sigma I, A'],[A, -1/rho]] in constraints-first order: its sparse
// columns are the rows of A (L21 = -A' diag(rho)) and its trailing supernode L22 is dense, so one solve is
// SpMV A' -> dense forward/backward substitution -> SpMV A.  All matrices are streamed from HBM/L2 in
// 32-row slices (lane = row, 256 B contiguous per warp load); every vector lives as [row][T] with the
// node index fastest, so one matrix entry is loaded once and used for all T nodes of the tile.
// FP64 throughout; no tensor cores (no dense contraction wider than T <= 8 right-hand sides).
#include <cuda_runtime.h>
#include <math.h>

#include "bqp_internal.h"

namespace bqp {

namespace {

constexpr int kRedSlots = 18;   // quantities reduced across the CTA in one termination check

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {   // xor butterfly: same value, same order on every lane
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// acc[t] += a * in[t], t < T, `in` in shared memory (16-byte aligned when T >= 2)
template <int T>
__device__ __forceinline__ void fma_row(double a, const double *__restrict__ in, double (&acc)[T]) {
  if constexpr (T == 1) {
    acc[0] = fma(a, in[0], acc[0]);
  } else {
#pragma unroll
    for (int t = 0; t < T; t += 2) {
      const double2 v = *reinterpret_cast<const double2 *>(in + t);
      acc[t] = fma(a, v.x, acc[t]);
      acc[t + 1] = fma(a, v.y, acc[t + 1]);
    }
  }
}

// acc += (slice s of M) * in, for the row owned by this lane
template <int T>
__device__ __forceinline__ void slice_dot(const DevMat &M, int s, const double *__restrict__ in, int lane,
                                          double (&acc)[T]) {
  const int p0 = __ldg(M.sptr + s), w = __ldg(M.sptr + s + 1) - p0;
  const double *__restrict__ v = M.vals + (size_t)p0 * 32 + lane;
  const int ip = __ldg(M.iptr + s);
  constexpr int U = 16;
  int j = 0;
  if (ip < 0) {
    for (; j + U <= w; j += U) {
      double a[U];
#pragma unroll
      for (int u = 0; u < U; u++) a[u] = __ldg(v + (size_t)(j + u) * 32);
#pragma unroll
      for (int u = 0; u < U; u++) fma_row<T>(a[u], in + (size_t)(j + u) * T, acc);
    }
    for (; j < w; j++) fma_row<T>(__ldg(v + (size_t)j * 32), in + (size_t)j * T, acc);
  } else {
    const int *__restrict__ ix = M.idx + (size_t)ip * 32 + lane;
    for (; j + U <= w; j += U) {
      double a[U];
      int c[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        a[u] = __ldg(v + (size_t)(j + u) * 32);
        c[u] = __ldg(ix + (size_t)(j + u) * 32);
      }
#pragma unroll
      for (int u = 0; u < U; u++) fma_row<T>(a[u], in + (size_t)c[u] * T, acc);
    }
    for (; j < w; j++) fma_row<T>(__ldg(v + (size_t)j * 32), in + (size_t)__ldg(ix + (size_t)j * 32) * T, acc);
  }
}

// acc[t] (+)= sum_{c<32} Lblk[c*ld] * vec[c][t]   (one 32x32 block of the dense tail, lane = row or column)
template <int T>
__device__ __forceinline__ void block_dot(const double *__restrict__ Lblk, size_t ld, const double *__restrict__ vec,
                                          double (&acc)[T]) {
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; c++) a[c] = __ldg(Lblk + (size_t)c * ld);
#pragma unroll
  for (int c = 0; c < 32; c++) fma_row<T>(a[c], vec + c * T, acc);
}

// In-place solve of L22 D2 L22' v = b for the T right-hand sides in shared memory b[npad][T].
template <int T>
__device__ void dense_tail_solve(const DevInstance &I, double *__restrict__ b, int warp, int nwarps, int lane) {
  const int np = I.npad, nb = np / kNB;
  // ---- forward: block columns, right-looking.  Warp 0 owns the next diagonal block.
  const double *Lc = I.Lcol;
  if (warp == 0) {
    double acc[T];
#pragma unroll
    for (int t = 0; t < T; t++) acc[t] = b[lane * T + t];
    block_dot<T>(Lc + lane, np, b, acc);
    __syncwarp();
#pragma unroll
    for (int t = 0; t < T; t++) b[lane * T + t] = acc[t];
  }
  __syncthreads();
  for (int J = 0; J + 1 < nb; J++) {
    const int r0 = J * kNB, ld = np - r0;
    const double *yJ = b + (size_t)r0 * T;
    for (int K = J + 1 + warp; K < nb; K += nwarps) {
      double acc[T];
#pragma unroll
      for (int t = 0; t < T; t++) acc[t] = 0.0;
      block_dot<T>(Lc + (size_t)(K * kNB - r0) + lane, ld, yJ, acc);
      double *bk = b + (size_t)(K * kNB + lane) * T;
#pragma unroll
      for (int t = 0; t < T; t++) acc[t] = bk[t] - acc[t];
      if (K == J + 1) {   // warp 0: finish block J+1 with its inverted diagonal block right away
        const double *Ln = Lc + (size_t)kNB * ld;   // block column J+1, ld-32 rows
#pragma unroll
        for (int t = 0; t < T; t++) bk[t] = acc[t];
        __syncwarp();
        block_dot<T>(Ln + lane, ld - kNB, b + (size_t)(r0 + kNB) * T, acc);
        __syncwarp();
      }
#pragma unroll
      for (int t = 0; t < T; t++) bk[t] = acc[t];
    }
    Lc += (size_t)kNB * ld;
    __syncthreads();
  }
  // ---- diagonal
  for (int e = threadIdx.x; e < np * T; e += blockDim.x) b[e] *= __ldg(I.D2inv + e / T);
  __syncthreads();
  // ---- backward: block rows from the bottom, lane = column.
  const double *Lr = I.Lrow + (size_t)kNB * kNB * ((size_t)nb * (nb - 1) / 2);   // block row nb-1
  if (warp == 0) {
    const int r0 = (nb - 1) * kNB;
    double acc[T];
#pragma unroll
    for (int t = 0; t < T; t++) acc[t] = b[(size_t)(r0 + lane) * T + t];
    block_dot<T>(Lr + r0 + lane, (size_t)nb * kNB, b + (size_t)r0 * T, acc);
    __syncwarp();
#pragma unroll
    for (int t = 0; t < T; t++) b[(size_t)(r0 + lane) * T + t] = acc[t];
  }
  __syncthreads();
  for (int J = nb - 1; J >= 1; J--) {
    const int r0 = J * kNB;
    const size_t ld = (size_t)(J + 1) * kNB;
    const double *xJ = b + (size_t)r0 * T;
    for (int K = J - 1 - warp; K >= 0; K -= nwarps) {
      double acc[T];
#pragma unroll
      for (int t = 0; t < T; t++) acc[t] = 0.0;
      block_dot<T>(Lr + K * kNB + lane, ld, xJ, acc);
      double *bk = b + (size_t)(K * kNB + lane) * T;
#pragma unroll
      for (int t = 0; t < T; t++) acc[t] = bk[t] - acc[t];
      if (K == J - 1) {   // warp 0: transposed inverted diagonal block of block row J-1
        const double *Lp = Lr - (size_t)kNB * (ld - kNB);   // block row J-1, ld-32 columns
#pragma unroll
        for (int t = 0; t < T; t++) bk[t] = acc[t];
        __syncwarp();
        block_dot<T>(Lp + K * kNB + lane, ld - kNB, b + (size_t)(K * kNB) * T, acc);
        __syncwarp();
      }
#pragma unroll
      for (int t = 0; t < T; t++) bk[t] = acc[t];
    }
    Lr -= (size_t)kNB * (ld - kNB);
    __syncthreads();
  }
}

struct TileShared {
  DevInstance I;
  DevTile tile;
  double fin[kRedSlots][kMaxTT];
  int status[kMaxTT], iters[kMaxTT], newly[kMaxTT];
  int remaining;
};

// reduce v over the CTA for node t: warp shuffle, then one slot per warp; combined later in warp order
template <int T, int OP>   // OP 0: max, 1: sum, 2: min
__device__ __forceinline__ void red_put(double (&v)[T], double *red, int slot, int warp, int nwarps, int lane) {
#pragma unroll
  for (int t = 0; t < T; t++) {
    double r = OP == 0 ? warp_max(v[t]) : (OP == 1 ? warp_sum(v[t]) : warp_min(v[t]));
    if (lane == 0) red[((size_t)slot * nwarps + warp) * T + t] = r;
  }
}

template <int T>
__global__ void __launch_bounds__(kMaxThreads, 1)
admm_tile_kernel(const DevInstance *__restrict__ insts, const DevTile *__restrict__ tiles, const double *__restrict__ in,
                 double *__restrict__ out, double *__restrict__ work, NodeScalars *__restrict__ ns,
                 int *__restrict__ tile_iters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  TileShared &S = *reinterpret_cast<TileShared *>(smem_raw);
  if (tid == 0) {
    S.tile = tiles[blockIdx.x];
    S.I = insts[S.tile.inst];
    S.remaining = S.tile.nn;
  }
  if (tid < kMaxTT) { S.status[tid] = BQP_UNSOLVED; S.iters[tid] = 0; S.newly[tid] = 0; }
  __syncthreads();
  const DevInstance &I = S.I;
  const int n = I.n, m = I.m, np = I.npad, nn = S.tile.nn;
  const int mvec = ((m > np ? m : np) + 1) & ~1;
  double *vin = reinterpret_cast<double *>(smem_raw + ((sizeof(TileShared) + 15) & ~size_t(15)));   // [mvec][T]
  double *bb = vin + (size_t)mvec * T;                                                                  // [npad][T]
  double *red = bb + (size_t)np * T;                                                                    // [slots][nwarps][T]
  double *W = work + S.tile.work_off;
  double *gx = W, *gdx = gx + (size_t)n * T, *gz = gdx + (size_t)n * T, *gy = gz + (size_t)m * T,
         *gl = gy + (size_t)m * T, *gu = gl + (size_t)m * T, *gdy = gu + (size_t)m * T;
  const double alpha = I.alpha, sigma = I.sigma;

  // ------------------------------------------------------------------ prologue (node.py:102-105)
  for (int e = tid; e < m * T; e += blockDim.x) {
    const int i = e / T, t = e - i * T;
    double lo = -kInfty, up = kInfty, yv = 0.0;
    if (t < nn) {
      const double *p = in + S.tile.in_off[t];
      lo = fmax(p[i], -kInfty);
      up = fmin(p[m + i], kInfty);
      yv = I.c * __ldg(I.Einv + i) * p[2 * (size_t)m + n + i];
    }
    const double ei = __ldg(I.E + i);
    gl[e] = ei * lo; gu[e] = ei * up; gy[e] = yv;
  }
  for (int e = tid; e < np * T; e += blockDim.x) {
    const int j = e / T, t = e - j * T;
    double xv = 0.0;
    if (j < n && t < nn) xv = __ldg(I.Dinv + j) * in[S.tile.in_off[t] + 2 * (size_t)m + j];
    bb[e] = xv;
    if (j < n) gx[e] = xv;
  }
  for (int e = tid + m * T; e < mvec * T; e += blockDim.x) vin[e] = 0.0;
  __syncthreads();
  for (int s = warp; s < I.Ab.nslices; s += nwarps) {   // z = A x ; vin = rho z - y
    double acc[T];
#pragma unroll
    for (int t = 0; t < T; t++) acc[t] = 0.0;
    slice_dot<T>(I.Ab, s, bb, lane, acc);
    const int i = s * 32 + lane;
    if (i < m) {
      const double rho = __ldg(I.rho + i);
#pragma unroll
      for (int t = 0; t < T; t++) {
        gz[(size_t)i * T + t] = acc[t];
        vin[(size_t)i * T + t] = rho * acc[t] - gy[(size_t)i * T + t];
      }
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ ADMM loop (osqp_solve)
  const int max_iter = I.max_iter, check_every = I.check_every;
  int iter = 0;
  for (iter = 1; iter <= max_iter; iter++) {
    const bool do_check = (iter % check_every == 0) || iter == max_iter;
    // rhs of the reduced system: b = sigma x - q + A'(rho z - y)
    for (int s = warp; s < I.At.nslices; s += nwarps) {
      double acc[T];
#pragma unroll
      for (int t = 0; t < T; t++) acc[t] = 0.0;
      slice_dot<T>(I.At, s, vin, lane, acc);
      const int j = s * 32 + lane;
      if (j < n) {
        const double qj = __ldg(I.q + j);
#pragma unroll
        for (int t = 0; t < T; t++) bb[(size_t)j * T + t] = sigma * gx[(size_t)j * T + t] - qj + acc[t];
      }
    }
    __syncthreads();
    dense_tail_solve<T>(I, bb, warp, nwarps, lane);   // ends with __syncthreads
    // x update (and dx at check iterations)
    for (int e = tid; e < n * T; e += blockDim.x) {
      const double xp = gx[e], xn = alpha * bb[e] + (1.0 - alpha) * xp;
      gx[e] = xn;
      if (do_check) gdx[e] = xn - xp;
    }
    // zt = A xt, projection, dual update, next rhs_z
    for (int s = warp; s < I.Ab.nslices; s += nwarps) {
      double acc[T];
#pragma unroll
      for (int t = 0; t < T; t++) acc[t] = 0.0;
      slice_dot<T>(I.Ab, s, bb, lane, acc);
      const int i = s * 32 + lane;
      if (i < m) {
        const double rho = __ldg(I.rho + i), rinv = __ldg(I.rho_inv + i);
#pragma unroll
        for (int t = 0; t < T; t++) {
          const size_t e = (size_t)i * T + t;
          const double zp = gz[e], yv = gy[e];
          const double zr = alpha * acc[t] + (1.0 - alpha) * zp;
          double zn = zr + rinv * yv;
          zn = fmin(fmax(zn, gl[e]), gu[e]);
          const double dy = rho * (zr - zn), yn = yv + dy;
          gz[e] = zn; gy[e] = yn;
          if (do_check) gdy[e] = dy;
          vin[e] = rho * zn - yn;
        }
      }
    }
    __syncthreads();
    if (!do_check) continue;

    // -------------------------------------------------------------- termination check (update_info + check_termination)
    for (int e = tid; e < n * T; e += blockDim.x) bb[e] = gx[e];
    for (int e = tid; e < m * T; e += blockDim.x) vin[e] = gy[e];
    __syncthreads();
    {
      double dr[T], b1[T], b2[T], quad[T], lin[T];
#pragma unroll
      for (int t = 0; t < T; t++) { dr[t] = 0; b1[t] = 0; b2[t] = 0; quad[t] = 0; lin[t] = 0; }
      for (int s = warp; s < I.Pm.nslices; s += nwarps) {
        double px[T], aty[T];
#pragma unroll
        for (int t = 0; t < T; t++) { px[t] = 0; aty[t] = 0; }
        slice_dot<T>(I.Pm, s, bb, lane, px);
        slice_dot<T>(I.At, s, vin, lane, aty);
        const int j = s * 32 + lane;
        if (j < n) {
          const double qj = __ldg(I.q + j), di = __ldg(I.Dinv + j);
#pragma unroll
          for (int t = 0; t < T; t++) {
            const double xj = bb[(size_t)j * T + t];
            dr[t] = fmax(dr[t], fabs(di * (px[t] + qj + aty[t])));
            b1[t] = fmax(b1[t], fabs(di * px[t]));
            b2[t] = fmax(b2[t], fabs(di * aty[t]));
            quad[t] += xj * px[t];
            lin[t] += qj * xj;
          }
        }
      }
      red_put<T, 0>(dr, red, 0, warp, nwarps, lane);
      red_put<T, 0>(b1, red, 1, warp, nwarps, lane);
      red_put<T, 0>(b2, red, 2, warp, nwarps, lane);
      red_put<T, 1>(quad, red, 3, warp, nwarps, lane);
      red_put<T, 1>(lin, red, 4, warp, nwarps, lane);
      double pr[T], a1[T], a2[T];
#pragma unroll
      for (int t = 0; t < T; t++) { pr[t] = 0; a1[t] = 0; a2[t] = 0; }
      for (int s = warp; s < I.Ab.nslices; s += nwarps) {
        double ax[T];
#pragma unroll
        for (int t = 0; t < T; t++) ax[t] = 0;
        slice_dot<T>(I.Ab, s, bb, lane, ax);
        const int i = s * 32 + lane;
        if (i < m) {
          const double ei = __ldg(I.Einv + i);
#pragma unroll
          for (int t = 0; t < T; t++) {
            const double zv = gz[(size_t)i * T + t];
            pr[t] = fmax(pr[t], fabs(ei * (ax[t] - zv)));
            a1[t] = fmax(a1[t], fabs(ei * ax[t]));
            a2[t] = fmax(a2[t], fabs(ei * zv));
          }
        }
      }
      red_put<T, 0>(pr, red, 5, warp, nwarps, lane);
      red_put<T, 0>(a1, red, 6, warp, nwarps, lane);
      red_put<T, 0>(a2, red, 7, warp, nwarps, lane);
    }
    __syncthreads();
    // certificates: vin <- projected dy, bb <- dx
    {
      double ndy[T], lhs[T], ndx[T], qdx[T];
#pragma unroll
      for (int t = 0; t < T; t++) { ndy[t] = 0; lhs[t] = 0; ndx[t] = 0; qdx[t] = 0; }
      // rows are assigned to fixed (warp, lane) positions so the sums have a fixed order
      for (int i0 = warp * 32; i0 < m; i0 += nwarps * 32) {
        const int i = i0 + lane;
        if (i < m) {
          const double ei = __ldg(I.E + i);
#pragma unroll
          for (int t = 0; t < T; t++) {
            const size_t e = (size_t)i * T + t;
            const double lo = gl[e], up = gu[e];
            double d = gdy[e];
            if (up > kInfty * kMinScaling) {
              if (lo < -kInfty * kMinScaling) d = 0.0; else d = fmin(d, 0.0);
            } else if (lo < -kInfty * kMinScaling) d = fmax(d, 0.0);
            vin[e] = d;
            ndy[t] = fmax(ndy[t], fabs(ei * d));
            lhs[t] += up * fmax(d, 0.0) + lo * fmin(d, 0.0);
          }
        }
      }
      for (int j0 = warp * 32; j0 < n; j0 += nwarps * 32) {
        const int j = j0 + lane;
        if (j < n) {
          const double dj = __ldg(I.D + j), qj = __ldg(I.q + j);
#pragma unroll
          for (int t = 0; t < T; t++) {
            const double d = gdx[(size_t)j * T + t];
            bb[(size_t)j * T + t] = d;
            ndx[t] = fmax(ndx[t], fabs(dj * d));
            qdx[t] += qj * d;
          }
        }
      }
      red_put<T, 0>(ndy, red, 8, warp, nwarps, lane);
      red_put<T, 1>(lhs, red, 9, warp, nwarps, lane);
      red_put<T, 0>(ndx, red, 10, warp, nwarps, lane);
      red_put<T, 1>(qdx, red, 11, warp, nwarps, lane);
    }
    __syncthreads();
    {
      double t1[T], t2[T];
#pragma unroll
      for (int t = 0; t < T; t++) { t1[t] = 0; t2[t] = 0; }
      for (int s = warp; s < I.Pm.nslices; s += nwarps) {
        double atd[T], pdx[T];
#pragma unroll
        for (int t = 0; t < T; t++) { atd[t] = 0; pdx[t] = 0; }
        slice_dot<T>(I.At, s, vin, lane, atd);
        slice_dot<T>(I.Pm, s, bb, lane, pdx);
        const int j = s * 32 + lane;
        if (j < n) {
          const double di = __ldg(I.Dinv + j);
#pragma unroll
          for (int t = 0; t < T; t++) {
            t1[t] = fmax(t1[t], fabs(di * atd[t]));
            t2[t] = fmax(t2[t], fabs(di * pdx[t]));
          }
        }
      }
      red_put<T, 0>(t1, red, 12, warp, nwarps, lane);
      red_put<T, 0>(t2, red, 13, warp, nwarps, lane);
      double vu[T], vl[T];
#pragma unroll
      for (int t = 0; t < T; t++) { vu[t] = -INFINITY; vl[t] = INFINITY; }
      for (int s = warp; s < I.Ab.nslices; s += nwarps) {
        double adx[T];
#pragma unroll
        for (int t = 0; t < T; t++) adx[t] = 0;
        slice_dot<T>(I.Ab, s, bb, lane, adx);
        const int i = s * 32 + lane;
        if (i < m) {
          const double ei = __ldg(I.Einv + i);
#pragma unroll
          for (int t = 0; t < T; t++) {
            const size_t e = (size_t)i * T + t;
            const double v = ei * adx[t];
            if (gu[e] < kInfty * kMinScaling) vu[t] = fmax(vu[t], v);
            if (gl[e] > -kInfty * kMinScaling) vl[t] = fmin(vl[t], v);
          }
        }
      }
      red_put<T, 0>(vu, red, 14, warp, nwarps, lane);
      red_put<T, 2>(vl, red, 15, warp, nwarps, lane);
    }
    __syncthreads();
    for (int idx = tid; idx < 16 * T; idx += blockDim.x) {   // combine the per-warp partials in warp order
      const int slot = idx / T, t = idx - slot * T;
      const bool is_sum = (slot == 3 || slot == 4 || slot == 9 || slot == 11), is_min = (slot == 15);
      double r = red[((size_t)slot * nwarps) * T + t];
      for (int w = 1; w < nwarps; w++) {
        const double v = red[((size_t)slot * nwarps + w) * T + t];
        r = is_sum ? r + v : (is_min ? fmin(r, v) : fmax(r, v));
      }
      S.fin[slot][t] = r;
    }
    __syncthreads();
    if (tid < T) {
      const int t = tid;
      S.newly[t] = 0;
      if (t < nn && S.status[t] == BQP_UNSOLVED) {
        const double cinv = I.cinv, c = I.c;
        const double pri = S.fin[5][t], dua = cinv * S.fin[0][t];
        const double nAx = S.fin[6][t], nz = S.fin[7][t], nPx = cinv * S.fin[1][t], nAty = cinv * S.fin[2][t], nq = cinv * I.nq;
        const double obj = (0.5 * S.fin[3][t] + S.fin[4][t]) * cinv;
        int status = BQP_UNSOLVED;
        const int passes = (iter == max_iter) ? 2 : 1;   // second pass = OSQP's "approximate" test at max_iter
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
          ns[S.tile.node[t]] = r;
          atomicSub(&S.remaining, 1);
        }
      }
    }
    __syncthreads();
    // snapshot the iterates of nodes that just terminated (unscaled; NaN for certificates, as osqp returns)
    for (int t = 0; t < nn; t++) {
      if (!S.newly[t]) continue;
      const int st = S.status[t];
      const bool bad = !(st == BQP_SOLVED || st == BQP_SOLVED_INACCURATE || st == BQP_MAX_ITER_REACHED);
      double *ox = out + S.tile.out_off[t], *oy = ox + n;
      for (int j = tid; j < n; j += blockDim.x) ox[j] = bad ? NAN : __ldg(I.D + j) * gx[(size_t)j * T + t];
      for (int i = tid; i < m; i += blockDim.x) oy[i] = bad ? NAN : I.cinv * __ldg(I.E + i) * gy[(size_t)i * T + t];
    }
    if (S.remaining == 0) break;
    for (int e = tid; e < m * T; e += blockDim.x) vin[e] = __ldg(I.rho + e / T) * gz[e] - gy[e];
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) tile_iters[blockIdx.x] = iter > max_iter ? max_iter : iter;

  // ------------------------------------------------------------------ epilogue (node.py:128-143)
  // clip integer entries into the node's own bounds, then lower = 1/2 x'Px + q'x at the clipped point
  for (int t = 0; t < nn; t++) {
    const int st = S.status[t];
    if (!(st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED)) continue;
    double *ox = out + S.tile.out_off[t];
    const double *p = in + S.tile.in_off[t];
    for (int k = tid; k < I.n_int; k += blockDim.x) {
      const int j = __ldg(I.i_idx + k), row = m - I.n_int + k;
      ox[j] = fmin(fmax(ox[j], p[row]), p[m + row]);
    }
  }
  __syncthreads();
  for (int e = tid; e < np * T; e += blockDim.x) {
    const int j = e / T, t = e - j * T;
    double v = 0.0;
    if (j < n && t < nn) {
      const int st = S.status[t];
      if (st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED) v = __ldg(I.Dinv + j) * out[S.tile.out_off[t] + j];
    }
    bb[e] = v;
  }
  __syncthreads();
  {
    double quad[T], lin[T];
#pragma unroll
    for (int t = 0; t < T; t++) { quad[t] = 0; lin[t] = 0; }
    for (int s = warp; s < I.Pm.nslices; s += nwarps) {
      double px[T];
#pragma unroll
      for (int t = 0; t < T; t++) px[t] = 0;
      slice_dot<T>(I.Pm, s, bb, lane, px);
      const int j = s * 32 + lane;
      if (j < n) {
        const double qj = __ldg(I.q + j);
#pragma unroll
        for (int t = 0; t < T; t++) {
          const double xj = bb[(size_t)j * T + t];
          quad[t] += xj * px[t];
          lin[t] += qj * xj;
        }
      }
    }
    red_put<T, 1>(quad, red, 0, warp, nwarps, lane);
    red_put<T, 1>(lin, red, 1, warp, nwarps, lane);
  }
  __syncthreads();
  if (tid < nn) {
    const int t = tid, st = S.status[t];
    if (st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED) {
      double qd = 0, ln = 0;
      for (int w = 0; w < nwarps; w++) { qd += red[((size_t)0 * nwarps + w) * T + t]; ln += red[((size_t)1 * nwarps + w) * T + t]; }
      ns[S.tile.node[t]].lower = (0.5 * qd + ln) * I.cinv;
    }
  }
}

}  // namespace

size_t tile_smem_bytes(int n, int m, int tt, int threads) {
  const int np = ((n + kNB - 1) / kNB) * kNB;
  const int mvec = ((m > np ? m : np) + 1) & ~1;
  size_t b = (sizeof(TileShared) + 15) & ~size_t(15);
  b += ((size_t)mvec + np) * tt * 8;
  b += (size_t)kRedSlots * (threads / 32) * tt * 8;
  return b;
}

template <int T>
static int launch_t(int threads, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles, const double *d_in,
                    double *d_out, double *d_work, NodeScalars *d_ns, int *d_tile_iters, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(admm_tile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return BQP_E_CUDA;
  admm_tile_kernel<T><<<ntiles, threads, smem, st>>>(d_insts, d_tiles, d_in, d_out, d_work, d_ns, d_tile_iters);
  return cudaGetLastError() == cudaSuccess ? BQP_OK : BQP_E_CUDA;
}

int launch_admm(int tt, int threads, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles, const double *d_in,
                double *d_out, double *d_work, NodeScalars *d_ns, int *d_tile_iters, size_t smem_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  switch (tt) {
    case 1: return launch_t<1>(threads, d_insts, d_tiles, ntiles, d_in, d_out, d_work, d_ns, d_tile_iters, smem_bytes, st);
    case 2: return launch_t<2>(threads, d_insts, d_tiles, ntiles, d_in, d_out, d_work, d_ns, d_tile_iters, smem_bytes, st);
    case 4: return launch_t<4>(threads, d_insts, d_tiles, ntiles, d_in, d_out, d_work, d_ns, d_tile_iters, smem_bytes, st);
    case 8: return launch_t<8>(threads, d_insts, d_tiles, ntiles, d_in, d_out, d_work, d_ns, d_tile_iters, smem_bytes, st);
  }
  return BQP_E_ARG;
}

}  // namespace bqp
