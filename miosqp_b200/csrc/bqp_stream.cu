// bqp_stream.cu -- TMA-fed batched ADMM kernel for sm_100a (problems with >= 4 slices of variables).
//
// Same algorithm and node-tile ownership as bqp_kernels.cu (one CTA = up to T B&B leaves of one problem,
// whole OSQP loop in-kernel; /root/reference/miosqp/node.py:96-143), but every matrix the loop touches --
// the A' and A panels (the sparse columns of the KKT factor), the blocked dense tail L22 with explicitly
// inverted diagonal super-blocks, and P for the termination checks -- is laid out once on the host in
// CONSUMPTION ORDER (HostStream, bqp_internal.h) and streamed through a ring of shared-memory stages by
// cp.async.bulk (TMA, 1-D) + mbarrier:
//
//   warp 12         : producer.  Lane q feeds quad q: it walks the same control flow as the consumers and issues
//                     one bulk copy per stage (8 KiB of values, + 4 KiB of column indices for sparse groups)
//                     into quad q's private ring of slots (canonical full/empty mbarrier pipeline per quad).
//   warps 0..11     : consumers.  Warp w owns slice w (32 rows, lane = row) of the current group; the four
//                     warps of a quad share a stage.  One matrix entry is read from shared memory once and
//                     used for all T nodes (vectors are [row][T], node index fastest).
//
// The triangular solve is a sequence of such groups too: per super-block J of L22 a diagonal group
// (y_J = inv(L_JJ) b_J as a triangular mat-vec) and an update group (b_below -= L_below,J y_J), so the only
// serial dependency left is one consumer barrier per super-block, not one per column.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "bqp_internal.h"

namespace bqp {

namespace {

constexpr int kConsumerWarps = kStreamWarps;
constexpr int kConsumers = kConsumerWarps * 32;
constexpr int kStreamThreads = kConsumers + 32;
constexpr int kRed = 16;          // reduced quantities per termination check
constexpr int kMaxGroups = 96;    // groups cached in shared memory

// ------------------------------------------------------------------ mbarrier / TMA wrappers (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a broken producer/consumer protocol traps (reported as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) __trap();   // ~4 s at 2 GHz
  }
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }

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
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

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

struct StreamShared {
  DevInstance I;
  DevTile tile;
  double fin[kRed][kMaxTT];
  int status[kMaxTT], iters[kMaxTT], newly[kMaxTT];
  int remaining;
  StreamGroup groups[kMaxGroups];
};

constexpr int kQuads = kConsumerWarps / 4;

// Per-quad ring state.  A consumer warp tracks the ring of its own quad; producer lane q tracks quad q's.
struct Ring {
  unsigned char *base;      // first slot of this quad
  uint64_t *full, *empty;   // [nslots] each, this quad's barriers
  int nslots, slot_bytes;
  int slot;                 // next slot to fill / consume
  uint32_t phase;           // parity of the current pass over the ring
  __device__ __forceinline__ void advance() {
    if (++slot == nslots) { slot = 0; phase ^= 1; }
  }
};

__device__ __forceinline__ void l2_prefetch(const void *src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

// producer lane q (< kQuads): one bulk copy per stage of quad q of group G.  When `wsrc` is given (A' groups in
// w-in-stage mode) a second, small bulk copy appends the kKC rows of the input vector w = rho z - y that the stage
// multiplies, fetched from the tile's global workspace (written by the consumers, published through `wbar`).
// Optional L2 prefetch `prefetch_ahead` bytes further down the (cyclic) per-iteration stream.
__device__ __forceinline__ void produce_group(const StreamGroup &G, const unsigned char *__restrict__ stream, long long iter_bytes,
                                              long long prefetch_ahead, const double *wsrc, int T, Ring &R, int lane) {
  if (lane < kQuads) {
    const uint32_t sb = kStageValBytes + (G.sparse ? kStageIdxBytes : 0);
    const uint32_t wb = wsrc ? (uint32_t)(kKC * T * 8) : 0u;
    int before = 0;
    for (int q = 0; q < lane; q++) before += G.qch[q];
    const long long off0 = G.data_off + (long long)before * sb;
    const unsigned char *src = stream + off0;
    const double *w = wsrc ? wsrc + (size_t)(G.in_off + G.qcol0[lane]) * T : nullptr;
    const int nch = G.qch[lane];
    for (int c = 0; c < nch; c++) {
      mbar_wait(R.empty + R.slot, R.phase ^ 1);     // passes at once on the first lap
      mbar_expect_tx(R.full + R.slot, sb + wb);
      unsigned char *dst = R.base + (size_t)R.slot * R.slot_bytes;
      tma_load_1d(dst, src, sb, R.full + R.slot);
      if (wb) tma_load_1d(dst + kStageValBytes, w + (size_t)c * kKC * T, wb, R.full + R.slot);
      if (prefetch_ahead > 0 && G.data_off < iter_bytes) {
        long long o = off0 + (long long)c * sb + prefetch_ahead;
        if (o >= iter_bytes) o -= iter_bytes;
        if (o + sb <= iter_bytes) l2_prefetch(stream + o, sb);
      }
      src += sb;
      R.advance();
    }
  }
  __syncwarp();   // keep the producer warp converged: its lanes take the CTA-wide barriers together
}

// consumer: acc += (this warp's slice of group G) * in
template <int T>
__device__ __forceinline__ void consume_group(const StreamGroup &G, const double *__restrict__ in, Ring &R, int warp, int lane,
                                              double (&acc)[T], bool wtail = false) {
  const int q = warp >> 2, wq = warp & 3;
  const int myc = G.qch[q];
  const int kp = lane >> 3;
  double acc4[4][T];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int t = 0; t < T; t++) acc4[i][t] = 0.0;
  if (myc > 0) {
    const double *inq = in + (size_t)(G.in_off + (G.sparse ? 0 : G.qcol0[q])) * T;
    for (int c = 0; c < myc; c++) {
      const int slot = R.slot;
      mbar_wait(R.full + slot, R.phase);
      const unsigned char *stage = R.base + (size_t)slot * R.slot_bytes;
      if (!G.sparse) {
        // register-blocked dense stage: this lane owns rows r8+8i (i<4) and the columns 4j+kp of the chunk
        const double *v = reinterpret_cast<const double *>(stage) + (size_t)wq * (kKC * 32) + lane * 2;
        // input rows of this chunk: from the vector in shared memory, or from the stage's own tail (w-in-stage A')
        const double *x = wtail ? reinterpret_cast<const double *>(stage + kStageValBytes) + kp * T
                                : inq + ((size_t)c * kKC + kp) * T;
#pragma unroll
        for (int j = 0; j < kKC / 4; j++) {
          const double2 a01 = *reinterpret_cast<const double2 *>(v + j * 128);
          const double2 a23 = *reinterpret_cast<const double2 *>(v + j * 128 + 64);
          double xv[T];
          if constexpr (T == 1) {
            xv[0] = x[(size_t)j * 4 * T];
          } else {
#pragma unroll
            for (int t = 0; t < T; t += 2) {
              const double2 w = *reinterpret_cast<const double2 *>(x + (size_t)j * 4 * T + t);
              xv[t] = w.x; xv[t + 1] = w.y;
            }
          }
#pragma unroll
          for (int t = 0; t < T; t++) {
            acc4[0][t] = fma(a01.x, xv[t], acc4[0][t]);
            acc4[1][t] = fma(a01.y, xv[t], acc4[1][t]);
            acc4[2][t] = fma(a23.x, xv[t], acc4[2][t]);
            acc4[3][t] = fma(a23.y, xv[t], acc4[3][t]);
          }
        }
      } else {
        const double *v = reinterpret_cast<const double *>(stage) + (wq * kKC) * 32 + lane;
        const int *ix = reinterpret_cast<const int *>(stage + kStageValBytes) + (wq * kKC) * 32 + lane;
        double a[kKC];
        int col[kKC];
#pragma unroll
        for (int j = 0; j < kKC; j++) { a[j] = v[j * 32]; col[j] = ix[j * 32]; }
#pragma unroll
        for (int j = 0; j < kKC; j++) fma_row<T>(a[j], inq + (size_t)col[j] * T, acc);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(R.empty + slot);
      R.advance();
    }
  }
  if (myc > 0 && !G.sparse) {
    // combine the four column parts (lanes kp = 0..3 of every r8) in a fixed order; lane (kp, r8) keeps row r8+8kp
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int t = 0; t < T; t++) {
        double v = acc4[i][t];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        acc4[i][t] = v;
      }
#pragma unroll
    for (int t = 0; t < T; t++) {
      const double lo = kp & 1 ? acc4[1][t] : acc4[0][t], hi = kp & 1 ? acc4[3][t] : acc4[2][t];
      acc[t] += kp & 2 ? hi : lo;
    }
  }
}

template <int T, int OP>   // OP 0: max, 1: sum, 2: min
__device__ __forceinline__ void red_put(double (&v)[T], double *red, int slot, int warp, int lane) {
#pragma unroll
  for (int t = 0; t < T; t++) {
    double r = OP == 0 ? warp_max(v[t]) : (OP == 1 ? warp_sum(v[t]) : warp_min(v[t]));
    if (lane == 0) red[((size_t)slot * kConsumerWarps + warp) * T + t] = r;
  }
}

template <int T>
__device__ __forceinline__ void zero(double (&a)[T]) {
#pragma unroll
  for (int t = 0; t < T; t++) a[t] = 0.0;
}

template <int T>
__global__ void __launch_bounds__(kStreamThreads, 1)
admm_stream_kernel(const DevInstance *__restrict__ insts, const DevTile *__restrict__ tiles, const double *__restrict__ in,
                   double *__restrict__ out, double *__restrict__ work, NodeScalars *__restrict__ ns,
                   int *__restrict__ tile_iters, int nslots, int slot_bytes, int w_in_stage, long long prefetch_ahead,
                   double *__restrict__ state) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = warp == kConsumerWarps;
  StreamShared &S = *reinterpret_cast<StreamShared *>(smem_raw);
  if (tid == 0) {
    S.tile = tiles[blockIdx.x];
    S.I = insts[S.tile.inst];
    S.remaining = S.tile.nn;
  }
  if (tid < kMaxTT) { S.status[tid] = BQP_UNSOLVED; S.iters[tid] = 0; S.newly[tid] = 0; }
  __syncthreads();
  const DevInstance &I = S.I;
  const int n = I.n, m = I.m, np = I.npad, nn = S.tile.nn;
  const int iter_begin = S.tile.iter_begin, iter_end = S.tile.iter_end;   // this round's slice of the ADMM loop
  const int ngroups = I.g_pm[1];
  for (int g = tid; g < ngroups; g += blockDim.x) S.groups[g] = I.groups[g];
  const int mvec = w_in_stage ? 0 : ((m > np ? m : np) + kKC + 15) & ~15;   // A' input vector in smem only when needed
  size_t off = (sizeof(StreamShared) + 15) & ~size_t(15);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + off);
  uint64_t *wbar = bars + 2 * (size_t)kQuads * nslots;       // "w published" (consumers -> producer), w-in-stage mode
  off += sizeof(uint64_t) * (2 * (size_t)kQuads * nslots + 2);
  off = (off + 15) & ~size_t(15);
  double *vin = reinterpret_cast<double *>(smem_raw + off);   // [mvec][T]
  double *bb = vin + (size_t)mvec * T;                        // [np + kKC][T]
  double *red = bb + (size_t)(np + kKC) * T;                  // [kRed][16][T]
  off += ((size_t)mvec + np + kKC + (size_t)kRed * kConsumerWarps) * T * 8;
  off = (off + 127) & ~size_t(127);
  // rings: quad q owns slots [q*nslots, (q+1)*nslots); consumer warps look at their quad, producer lane q at quad q
  const int rq = producer ? (lane < kQuads ? lane : 0) : (warp >> 2);
  Ring R;
  R.base = smem_raw + off + (size_t)rq * nslots * slot_bytes;
  R.full = bars + (size_t)rq * nslots; R.empty = bars + (size_t)(kQuads + rq) * nslots;
  R.nslots = nslots; R.slot_bytes = slot_bytes; R.slot = 0; R.phase = 0;
  if (tid == 0) {
    for (int s = 0; s < kQuads * nslots; s++) { mbar_init(bars + s, 1); mbar_init(bars + kQuads * nslots + s, 4); }
    mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int max_iter = I.max_iter, check_every = I.check_every;
  const unsigned char *stream = I.stream;
  const long long iter_bytes = S.groups[I.g_pm[0]].data_off;   // A', L fwd, L bwd, A: what one iteration streams

  double *W = work + S.tile.work_off;
  double *gw = W + tile_w_offset(n, m, T);             // A' input vector in global memory (w-in-stage mode)

  // =============================================================== producer warp: mirror of the consumer control flow
  if (producer) {
    auto produce = [&](const int (&range)[2]) {
      for (int g = range[0]; g < range[1]; g++)
        produce_group(S.groups[g], stream, iter_bytes, prefetch_ahead, nullptr, T, R, lane);
    };
    uint32_t wphase = 0;
    auto produce_at = [&]() {        // every pass over A' waits for the consumers to have published its input vector
      if (w_in_stage) { mbar_wait(wbar, wphase); wphase ^= 1; }
      for (int g = I.g_at[0]; g < I.g_at[1]; g++)
        produce_group(S.groups[g], stream, iter_bytes, prefetch_ahead, w_in_stage ? gw : nullptr, T, R, lane);
    };
    if (iter_begin == 0) produce(I.g_ab);              // prologue: z = A x (a resumed round restores z instead)
    for (int iter = iter_begin + 1; iter <= iter_end; iter++) {
      const bool do_check = (iter % check_every == 0) || iter == max_iter;
      produce_at(); produce(I.g_fw); produce(I.g_bw); produce(I.g_ab);
      if (!do_check) continue;
      produce(I.g_pm); produce_at(); produce(I.g_ab);        // P x, A' y, A x
      produce_at(); produce(I.g_pm); produce(I.g_ab);        // A' dy, P dx, A dx
      __syncthreads();                                       // decision published by the consumers
      if (S.remaining == 0 || iter == iter_end) break;
    }
    produce(I.g_pm);                                   // epilogue objective
    return;
  }

  // =============================================================== consumer warps
  // wv: where the A' input vector is written.  publish_w(): make it visible to the producer's TMA reads (generic ->
  // async proxy fence by every writer, consumer barrier, one arrival on wbar); plain consumer barrier otherwise.
  double *const wv = w_in_stage ? gw : vin;
  auto publish_w = [&]() {
    if (w_in_stage) asm volatile("fence.proxy.async;" ::: "memory");
    consumer_bar();
    if (w_in_stage && tid == 0) mbar_arrive(wbar);
  };
  double *gx = W, *gdx = gx + (size_t)n * T, *gz = gdx + (size_t)n * T, *gy = gz + (size_t)m * T,
         *gl = gy + (size_t)m * T, *gu = gl + (size_t)m * T, *gdy = gu + (size_t)m * T, *gpx = gdy + (size_t)m * T;
  const double alpha = I.alpha, sigma = I.sigma;

  // ---- prologue (node.py:102-105)
  for (int e = tid; e < m * T; e += kConsumers) {
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
  for (int e = tid; e < (np + kKC) * T; e += kConsumers) {
    const int j = e / T, t = e - j * T;
    double xv = 0.0;
    if (j < n && t < nn)
      xv = iter_begin == 0 ? __ldg(I.Dinv + j) * in[S.tile.in_off[t] + 2 * (size_t)m + j] : state[S.tile.state_off[t] + j];
    bb[e] = xv;
    if (j < n) gx[e] = xv;
  }
  if (iter_begin > 0) {   // resumed round: z and y continue from the saved ADMM state, w = rho z - y
    for (int e = tid; e < m * T; e += kConsumers) {
      const int i = e / T, t = e - i * T;
      double zv = 0.0, yv = 0.0;
      if (t < nn) { const double *sp = state + S.tile.state_off[t] + n; zv = sp[i]; yv = sp[m + i]; }
      gz[e] = zv; gy[e] = yv;
    }
  }
  for (int e = tid + m * T; e < (w_in_stage ? m + 32 : mvec) * T; e += kConsumers) wv[e] = 0.0;
  consumer_bar();
  if (iter_begin > 0) {
    for (int e = tid; e < m * T; e += kConsumers) wv[e] = __ldg(I.rho + e / T) * gz[e] - gy[e];
  }
  for (int g = I.g_ab[0]; g < (iter_begin == 0 ? I.g_ab[1] : I.g_ab[0]); g++) {   // z = A x ; w = rho z - y
    const StreamGroup &G = S.groups[g];
    double acc[T]; zero<T>(acc);
    consume_group<T>(G, bb, R, warp, lane, acc);
    const int i = G.row0 + warp * 32 + lane;
    if (warp < G.nsl && i < m) {
      const double rho = __ldg(I.rho + i);
#pragma unroll
      for (int t = 0; t < T; t++) {
        gz[(size_t)i * T + t] = acc[t];
        wv[(size_t)i * T + t] = rho * acc[t] - gy[(size_t)i * T + t];
      }
    }
  }
  publish_w();

  int iter = 0;
  for (iter = iter_begin + 1; iter <= iter_end; iter++) {
    const bool do_check = (iter % check_every == 0) || iter == max_iter;
    // ---- b = sigma x - q + A'(rho z - y)
    for (int g = I.g_at[0]; g < I.g_at[1]; g++) {
      const StreamGroup &G = S.groups[g];
      double acc[T]; zero<T>(acc);
      consume_group<T>(G, vin, R, warp, lane, acc, w_in_stage != 0);
      const int j = G.row0 + warp * 32 + lane;
      if (warp < G.nsl && j < n) {
        const double qj = __ldg(I.q + j);
#pragma unroll
        for (int t = 0; t < T; t++) bb[(size_t)j * T + t] = sigma * gx[(size_t)j * T + t] - qj + acc[t];
      }
    }
    consumer_bar();
    // ---- b <- L22^-T D2^-1 L22^-1 b : super-block sweeps
    for (int pass = 0; pass < 2; pass++) {
      const int g0 = pass ? I.g_bw[0] : I.g_fw[0], g1 = pass ? I.g_bw[1] : I.g_fw[1];
      if (pass) {
        for (int e = tid; e < np * T; e += kConsumers) bb[e] *= __ldg(I.D2inv + e / T);
        consumer_bar();
      }
      for (int g = g0; g < g1; g++) {
        const StreamGroup &G = S.groups[g];
        const bool diag = (G.kind == GK_FWD_D || G.kind == GK_BWD_D);
        double acc[T]; zero<T>(acc);
        consume_group<T>(G, bb, R, warp, lane, acc);
        const int r = G.row0 + warp * 32 + lane;
        const bool mine = warp < G.nsl && r < np;
        if (diag) {
          consumer_bar();                       // every warp has read b_J before it is overwritten
          if (mine) {
#pragma unroll
            for (int t = 0; t < T; t++) bb[(size_t)r * T + t] = acc[t];
          }
          consumer_bar();
        } else {
          if (mine) {
#pragma unroll
            for (int t = 0; t < T; t++) bb[(size_t)r * T + t] -= acc[t];
          }
          // the next diagonal group reads rows written here: barrier when the update phase ends
          if (g + 1 == g1 || S.groups[g + 1].kind != G.kind) consumer_bar();
        }
      }
    }
    // ---- x update (and dx at check iterations)
    for (int e = tid; e < n * T; e += kConsumers) {
      const double xp = gx[e], xn = alpha * bb[e] + (1.0 - alpha) * xp;
      gx[e] = xn;
      if (do_check) gdx[e] = xn - xp;
    }
    // ---- zt = A xt, projection, dual update, next rhs_z
    for (int g = I.g_ab[0]; g < I.g_ab[1]; g++) {
      const StreamGroup &G = S.groups[g];
      double acc[T]; zero<T>(acc);
      consume_group<T>(G, bb, R, warp, lane, acc);
      const int i = G.row0 + warp * 32 + lane;
      if (warp < G.nsl && i < m) {
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
          wv[e] = rho * zn - yn;
        }
      }
    }
    if (!do_check) { publish_w(); continue; }
    consumer_bar();

    // ---- termination check (update_info + check_termination)
    for (int e = tid; e < n * T; e += kConsumers) bb[e] = gx[e];
    for (int e = tid; e < m * T; e += kConsumers) wv[e] = gy[e];
    publish_w();
    {
      // P x and A' y share the row ownership (both have n rows): group k of P pairs with group k of A'
      double dr[T], b1[T], b2[T], quad[T], lin[T];
      zero<T>(dr); zero<T>(b1); zero<T>(b2); zero<T>(quad); zero<T>(lin);
      // P x goes through a per-tile scratch row (same thread writes and reads it back) to keep registers free
      const int npm = I.g_pm[1] - I.g_pm[0];
      for (int k = 0; k < npm; k++) {
        const StreamGroup &G = S.groups[I.g_pm[0] + k];
        double px[T]; zero<T>(px);
        consume_group<T>(G, bb, R, warp, lane, px);
        const int j = G.row0 + warp * 32 + lane;
        if (warp < G.nsl && j < n) {
#pragma unroll
          for (int t = 0; t < T; t++) gpx[(size_t)j * T + t] = px[t];
        }
      }
      for (int k = 0; k < npm; k++) {
        const StreamGroup &G = S.groups[I.g_at[0] + k];
        double aty[T]; zero<T>(aty);
        consume_group<T>(G, vin, R, warp, lane, aty, w_in_stage != 0);
        const int j = G.row0 + warp * 32 + lane;
        if (warp < G.nsl && j < n) {
          const double qj = __ldg(I.q + j), di = __ldg(I.Dinv + j);
#pragma unroll
          for (int t = 0; t < T; t++) {
            const double xj = bb[(size_t)j * T + t], pxj = gpx[(size_t)j * T + t];
            dr[t] = fmax(dr[t], fabs(di * (pxj + qj + aty[t])));
            b1[t] = fmax(b1[t], fabs(di * pxj));
            b2[t] = fmax(b2[t], fabs(di * aty[t]));
            quad[t] += xj * pxj;
            lin[t] += qj * xj;
          }
        }
      }
      red_put<T, 0>(dr, red, 0, warp, lane);
      red_put<T, 0>(b1, red, 1, warp, lane);
      red_put<T, 0>(b2, red, 2, warp, lane);
      red_put<T, 1>(quad, red, 3, warp, lane);
      red_put<T, 1>(lin, red, 4, warp, lane);
      double pr[T], a1[T], a2[T];
      zero<T>(pr); zero<T>(a1); zero<T>(a2);
      for (int g = I.g_ab[0]; g < I.g_ab[1]; g++) {
        const StreamGroup &G = S.groups[g];
        double ax[T]; zero<T>(ax);
        consume_group<T>(G, bb, R, warp, lane, ax);
        const int i = G.row0 + warp * 32 + lane;
        if (warp < G.nsl && i < m) {
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
      red_put<T, 0>(pr, red, 5, warp, lane);
      red_put<T, 0>(a1, red, 6, warp, lane);
      red_put<T, 0>(a2, red, 7, warp, lane);
    }
    consumer_bar();
    {
      double ndy[T], lhs[T], ndx[T], qdx[T];
      zero<T>(ndy); zero<T>(lhs); zero<T>(ndx); zero<T>(qdx);
      for (int i0 = warp * 32; i0 < m; i0 += kConsumers) {
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
            wv[e] = d;
            ndy[t] = fmax(ndy[t], fabs(ei * d));
            lhs[t] += up * fmax(d, 0.0) + lo * fmin(d, 0.0);
          }
        }
      }
      for (int j0 = warp * 32; j0 < n; j0 += kConsumers) {
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
      red_put<T, 0>(ndy, red, 8, warp, lane);
      red_put<T, 1>(lhs, red, 9, warp, lane);
      red_put<T, 0>(ndx, red, 10, warp, lane);
      red_put<T, 1>(qdx, red, 11, warp, lane);
    }
    publish_w();   // projected dy is the next A' input
    {
      double t1[T], t2[T];
      zero<T>(t1); zero<T>(t2);
      for (int g = I.g_at[0]; g < I.g_at[1]; g++) {
        const StreamGroup &G = S.groups[g];
        double atd[T]; zero<T>(atd);
        consume_group<T>(G, vin, R, warp, lane, atd, w_in_stage != 0);
        const int j = G.row0 + warp * 32 + lane;
        if (warp < G.nsl && j < n) {
          const double di = __ldg(I.Dinv + j);
#pragma unroll
          for (int t = 0; t < T; t++) t1[t] = fmax(t1[t], fabs(di * atd[t]));
        }
      }
      for (int g = I.g_pm[0]; g < I.g_pm[1]; g++) {
        const StreamGroup &G = S.groups[g];
        double pdx[T]; zero<T>(pdx);
        consume_group<T>(G, bb, R, warp, lane, pdx);
        const int j = G.row0 + warp * 32 + lane;
        if (warp < G.nsl && j < n) {
          const double di = __ldg(I.Dinv + j);
#pragma unroll
          for (int t = 0; t < T; t++) t2[t] = fmax(t2[t], fabs(di * pdx[t]));
        }
      }
      red_put<T, 0>(t1, red, 12, warp, lane);
      red_put<T, 0>(t2, red, 13, warp, lane);
      double vu[T], vl[T];
#pragma unroll
      for (int t = 0; t < T; t++) { vu[t] = -INFINITY; vl[t] = INFINITY; }
      for (int g = I.g_ab[0]; g < I.g_ab[1]; g++) {
        const StreamGroup &G = S.groups[g];
        double adx[T]; zero<T>(adx);
        consume_group<T>(G, bb, R, warp, lane, adx);
        const int i = G.row0 + warp * 32 + lane;
        if (warp < G.nsl && i < m) {
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
      red_put<T, 0>(vu, red, 14, warp, lane);
      red_put<T, 2>(vl, red, 15, warp, lane);
    }
    consumer_bar();
    for (int idx = tid; idx < kRed * T; idx += kConsumers) {   // combine the per-warp partials in warp order
      const int slot = idx / T, t = idx - slot * T;
      const bool is_sum = (slot == 3 || slot == 4 || slot == 9 || slot == 11), is_min = (slot == 15);
      double r = red[((size_t)slot * kConsumerWarps) * T + t];
      for (int w = 1; w < kConsumerWarps; w++) {
        const double v = red[((size_t)slot * kConsumerWarps + w) * T + t];
        r = is_sum ? r + v : (is_min ? fmin(r, v) : fmax(r, v));
      }
      S.fin[slot][t] = r;
    }
    consumer_bar();
    if (tid < T) {
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
    consumer_bar();
    for (int t = 0; t < nn; t++) {
      if (!S.newly[t]) continue;
      const int st = S.status[t];
      const bool bad = !(st == BQP_SOLVED || st == BQP_SOLVED_INACCURATE || st == BQP_MAX_ITER_REACHED);
      double *ox = out + S.tile.out_off[t], *oy = ox + n;
      for (int j = tid; j < n; j += kConsumers) ox[j] = bad ? NAN : __ldg(I.D + j) * gx[(size_t)j * T + t];
      for (int i = tid; i < m; i += kConsumers) oy[i] = bad ? NAN : I.cinv * __ldg(I.E + i) * gy[(size_t)i * T + t];
    }
    __syncthreads();   // with the producer warp: it reads S.remaining after this barrier
    if (S.remaining == 0 || iter == iter_end) break;
    for (int e = tid; e < m * T; e += kConsumers) wv[e] = __ldg(I.rho + e / T) * gz[e] - gy[e];
    publish_w();
  }
  consumer_bar();
  if (tid == 0) tile_iters[blockIdx.x] = (iter > iter_end ? iter_end : iter) - iter_begin;
  // nodes still running when this round's iteration budget ends: save the scaled ADMM state for the next round
  for (int t = 0; t < nn; t++) {
    if (S.status[t] != BQP_UNSOLVED) continue;
    double *sp = state + S.tile.state_off[t];
    for (int j = tid; j < n; j += kConsumers) sp[j] = gx[(size_t)j * T + t];
    for (int i = tid; i < m; i += kConsumers) { sp[n + i] = gz[(size_t)i * T + t]; sp[n + m + i] = gy[(size_t)i * T + t]; }
    if (tid == 0) { NodeScalars r; r.status = BQP_UNSOLVED; r.iters = iter_end; r.obj = r.pri_res = r.dua_res = r.lower = NAN; ns[S.tile.node[t]] = r; }
  }

  // ---- epilogue (node.py:128-143): clip integer entries, lower = 1/2 x'Px + q'x at the clipped point
  for (int t = 0; t < nn; t++) {
    const int st = S.status[t];
    if (!(st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED)) continue;
    double *ox = out + S.tile.out_off[t];
    const double *p = in + S.tile.in_off[t];
    for (int k = tid; k < I.n_int; k += kConsumers) {
      const int j = __ldg(I.i_idx + k), row = m - I.n_int + k;
      ox[j] = fmin(fmax(ox[j], p[row]), p[m + row]);
    }
  }
  consumer_bar();
  for (int e = tid; e < np * T; e += kConsumers) {
    const int j = e / T, t = e - j * T;
    double v = 0.0;
    if (j < n && t < nn) {
      const int st = S.status[t];
      if (st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED) v = __ldg(I.Dinv + j) * out[S.tile.out_off[t] + j];
    }
    bb[e] = v;
  }
  consumer_bar();
  {
    double quad[T], lin[T];
    zero<T>(quad); zero<T>(lin);
    for (int g = I.g_pm[0]; g < I.g_pm[1]; g++) {
      const StreamGroup &G = S.groups[g];
      double px[T]; zero<T>(px);
      consume_group<T>(G, bb, R, warp, lane, px);
      const int j = G.row0 + warp * 32 + lane;
      if (warp < G.nsl && j < n) {
        const double qj = __ldg(I.q + j);
#pragma unroll
        for (int t = 0; t < T; t++) {
          const double xj = bb[(size_t)j * T + t];
          quad[t] += xj * px[t];
          lin[t] += qj * xj;
        }
      }
    }
    red_put<T, 1>(quad, red, 0, warp, lane);
    red_put<T, 1>(lin, red, 1, warp, lane);
  }
  consumer_bar();
  if (tid < nn) {
    const int t = tid, st = S.status[t];
    if (st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED) {
      double qd = 0, ln = 0;
      for (int w = 0; w < kConsumerWarps; w++) { qd += red[((size_t)0 * kConsumerWarps + w) * T + t]; ln += red[((size_t)1 * kConsumerWarps + w) * T + t]; }
      ns[S.tile.node[t]].lower = (0.5 * qd + ln) * I.cinv;
    }
  }
}

}  // namespace

size_t stream_smem_bytes(int n, int m, int tt, int slot_bytes, int nslots, int w_in_stage) {
  const int np = ((n + kNB - 1) / kNB) * kNB;
  const int mvec = w_in_stage ? 0 : ((m > np ? m : np) + kKC + 15) & ~15;
  size_t off = (sizeof(StreamShared) + 15) & ~size_t(15);
  off += sizeof(uint64_t) * (2 * (size_t)kQuads * nslots + 2);
  off = (off + 15) & ~size_t(15);
  off += ((size_t)mvec + np + kKC + (size_t)kRed * kConsumerWarps) * tt * 8;
  off = (off + 127) & ~size_t(127);
  return off + (size_t)kQuads * nslots * slot_bytes;
}

template <int T>
static int launch_t(int slot_bytes, int nslots, int w_in_stage, double *d_state, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles, const double *d_in,
                    double *d_out, double *d_work, NodeScalars *d_ns, int *d_tile_iters, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(admm_stream_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return BQP_E_CUDA;
  long long prefetch_ahead = 0;   // experiment knob: L2 prefetch distance of the producer warp (KiB), off by default
  if (const char *pk = getenv("BQP_PREFETCH_KB")) prefetch_ahead = 1024LL * atoll(pk);
  admm_stream_kernel<T><<<ntiles, kStreamThreads, smem, st>>>(d_insts, d_tiles, d_in, d_out, d_work, d_ns, d_tile_iters, nslots, slot_bytes, w_in_stage, prefetch_ahead, d_state);
  return cudaGetLastError() == cudaSuccess ? BQP_OK : BQP_E_CUDA;
}

int launch_admm_stream(int tt, int slot_bytes, int nslots, int w_in_stage, double *d_state, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles,
                       const double *d_in, double *d_out, double *d_work, NodeScalars *d_ns, int *d_tile_iters,
                       size_t smem_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  switch (tt) {
    case 1: return launch_t<1>(slot_bytes, nslots, w_in_stage, d_state, d_insts, d_tiles, ntiles, d_in, d_out, d_work, d_ns, d_tile_iters, smem_bytes, st);
    case 2: return launch_t<2>(slot_bytes, nslots, w_in_stage, d_state, d_insts, d_tiles, ntiles, d_in, d_out, d_work, d_ns, d_tile_iters, smem_bytes, st);
    case 4: return launch_t<4>(slot_bytes, nslots, w_in_stage, d_state, d_insts, d_tiles, ntiles, d_in, d_out, d_work, d_ns, d_tile_iters, smem_bytes, st);
    case 8: return launch_t<8>(slot_bytes, nslots, w_in_stage, d_state, d_insts, d_tiles, ntiles, d_in, d_out, d_work, d_ns, d_tile_iters, smem_bytes, st);
  }
  return BQP_E_ARG;
}

}  // namespace bqp
