// bqp_small.cu -- shared-memory-resident ADMM kernel for small sparse problems: npad <= 64, at most 4 entries per row of A and
// 8 per column, m <= 192 -- BASELINE config 3 (power-converter MPC: n = 60, m = 150, 3 entries per row of A, 5 per column).
//
// Same tile ownership as the other kernels (a tile = up to 8 B&B leaves of one problem, whole OSQP loop in-kernel:
// /root/reference/miosqp/node.py:96-143 -- update(l,u), warm_start(x,y), solve(), clip + objective) and the restated
// iteration of the dense kernels,
//     b = sigma x - q + A'(rho z - y),   x~ = M b  (M = (P + sigma I + A' rho A)^-1, formed and guarded at setup),
//     z~ = A x~,   z = clip(alpha z~ + (1 - alpha) z + y / rho),   y += rho (alpha z~ + (1 - alpha) z - z_new),
// but nothing is streamed: the problem's matrices arrive ONCE per launch as one blob (M and P as FP64 mma fragments, A and
// A' as entry-major ELL with u16 indices, rho, 1 / rho, E, 1 / E per row: bqp_internal.h HostSmall) through TMA bulk copies
// into shared memory, and every iterate (x, z, y, l, u, w of the 8 leaves, [row][8] with the leaf index fastest) lives in
// shared memory or registers for the whole solve.  The direct-load kernel this replaces for these shapes (bqp_kernels.cu)
// re-reads the LDL' factor from L2 every iteration behind block-by-block triangular sweeps and keeps the iterates in global
// memory (3.9 us per iteration on the config-3 problem with ONE leaf per CTA; here 1.13 us with 8).
//
// Thread map (256 threads): thread tid owns the leaf PAIR p = tid & 3 (leaves 2p, 2p + 1) of
//   column-space row j = tid >> 2                (x, b, x~: 64 rows),  and of
//   row-space rows  r = (tid >> 2) + 64 i, i < 3 (z, y, l, u, w),
// which is also where the C fragment of mma.m8n8k4 puts the product rows of warp w's panel (rows 8 w + (lane >> 2),
// columns 2 (lane & 3) + {0, 1}): x~ = M b comes out of the FP64 tensor pipe in the registers of the thread that owns x.
// The thread's own ELL entries -- value and byte offset of 3 rows of A x 4 entries and one row of A' x 8 -- stay in registers
// for the whole launch, and the hot loop addresses shared memory by 32-bit address (ld.shared / st.shared): the first version
// read the entries from shared memory through a generic predicated loop over 64-bit pointers and spent 800 instructions per
// warp and iteration, now 377.  M's fragments are read from shared memory in the M b phase (in registers as well the kernel
// spilled and ran at 2.1 us); P's at the termination checks only.  What bounds the kernel now is the shared-memory data pipe
// (profiles/r02_small_kernel_ncu.json); the bounds l, u and the iterates z, y of the thread's own rows therefore live in
// registers too -- shared memory sees z and y at the termination checks only.
// Sums over rows use a fixed order (entries even / odd, per-thread rows ascending, xor butterflies over the 8 rows of a
// warp, warps ascending) and every leaf's arithmetic is its own: a node's result does not depend on the other nodes of the
// tile or on the launch.  Threads of leaf pairs without a leaf skip the vector phases.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "bqp_internal.h"

namespace bqp {

namespace {

constexpr int kT = 8;            // leaves per tile = N of the mma
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSlots = 16;       // quantities of one termination check
constexpr int kNPmax = 64;
constexpr int kKS = kNPmax / 4;  // k-steps of a full-width row panel = columns of M per thread
constexpr int kRegWA = 4, kRegWT = 8, kRegRows = 3;   // ELL entries a thread can keep in registers: 3 rows of A x 4, one row of A' x 8

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();   // ~2 s: a broken copy traps instead of hanging the GPU
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// shared-memory accesses of the hot loop by 32-bit address: one instruction each (through generic 64-bit pointers every access
// cost two or three more for the address)
__device__ __forceinline__ double2 lds2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double lds1(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts2(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}

struct SmallShared {
  DevInstance I;
  DevTile tile;
  double fin[kSlots][kT];
  int status[kT], iters[kT], newly[kT];
  int remaining;
  unsigned long long bar;
};

__host__ __device__ constexpr size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

template <int OP>   // 0 max, 1 sum, 2 min
__device__ __forceinline__ double rop(double a, double b) { return OP == 0 ? fmax(a, b) : (OP == 1 ? a + b : fmin(a, b)); }

// reduce the pair (v0, v1) over the 8 rows of the warp (lanes with the same lane & 3) and file it under (slot, warp)
template <int OP>
__device__ __forceinline__ void red_put(double v0, double v1, double *red, int slot, int warp, int lane) {
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    v0 = rop<OP>(v0, __shfl_xor_sync(0xffffffffu, v0, o));
    v1 = rop<OP>(v1, __shfl_xor_sync(0xffffffffu, v1, o));
  }
  if (lane < 4) *reinterpret_cast<double2 *>(red + ((size_t)slot * kWarps + warp) * kT + 2 * lane) = make_double2(v0, v1);
}

// One ELL row times a [row][8] vector for the thread's leaf pair, the row's entries in registers: value and element offset
// (row index * 8 + 2 * pair) -- valid for every vector of that space; padding entries are 0.0 times the vector's zero row.
// Every load is issued before the first use (index -> element -> fma one at a time is pure latency), even and odd entries
// accumulate separately: the order of addition is the same wherever the row is used.
template <int W>
__device__ __forceinline__ double2 reg_dot(const double (&val)[W], const uint32_t (&off)[W], const double *__restrict__ v) {
  double2 x[W], e = make_double2(0.0, 0.0), o = e;
#pragma unroll
  for (int k = 0; k < W; k++) {
    x[k] = make_double2(0.0, 0.0);
    if (val[k] != 0.0) x[k] = *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(v) + off[k]);
  }
#pragma unroll
  for (int k = 0; k < W; k += 2) {
    e.x = fma(val[k], x[k].x, e.x); e.y = fma(val[k], x[k].y, e.y);
    o.x = fma(val[k + 1], x[k + 1].x, o.x); o.y = fma(val[k + 1], x[k + 1].y, o.y);
  }
  return make_double2(e.x + o.x, e.y + o.y);
}
// the same arithmetic with the vector given by its 32-bit shared address (hot loop)
template <int W>
__device__ __forceinline__ double2 reg_dot_s(const double (&val)[W], const uint32_t (&off)[W], uint32_t v) {
  double2 x[W], e = make_double2(0.0, 0.0), o = e;
#pragma unroll
  for (int k = 0; k < W; k++) {      // padding entries (and explicit zeros) are not loaded: 1.25 -> 1.13 us per iteration at config 3, where
    x[k] = make_double2(0.0, 0.0);   // 56 % of the 4-wide slots of A are padding
    if (val[k] != 0.0) x[k] = lds2(v + off[k]);
  }
#pragma unroll
  for (int k = 0; k < W; k += 2) {
    e.x = fma(val[k], x[k].x, e.x); e.y = fma(val[k], x[k].y, e.y);
    o.x = fma(val[k + 1], x[k + 1].x, o.x); o.y = fma(val[k + 1], x[k + 1].y, o.y);
  }
  return make_double2(e.x + o.x, e.y + o.y);
}

// rows 8 warp .. 8 warp + 7 of (fragment-ordered matrix) * v for the 8 leaves; the calling thread gets row 8 warp + (lane >> 2),
// leaves 2 (lane & 3) + {0, 1}.  Two accumulator chains hide the mma latency.
__device__ __forceinline__ double2 frag_rows_smem(const double *__restrict__ F, int ks, int warp, int lane, const double *__restrict__ v) {
  double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
  const double *f = F + (size_t)warp * ks * 32 + lane;
  const double *b = v + (lane & 3) * kT + (lane >> 2);
  for (int s = 0; s < ks; s += 2) {
    dmma(c0, f[s * 32], b[s * 4 * kT]);
    dmma(c1, f[(s + 1) * 32], b[(s + 1) * 4 * kT]);
  }
  return make_double2(c0[0] + c1[0], c0[1] + c1[1]);
}

template <int NP>      // padded number of variables: 32 or 64 (a CTA whose problem has the other width leaves at once)
__global__ void __launch_bounds__(kThreads, 1)
admm_small_kernel(const DevInstance *__restrict__ insts, const DevTile *__restrict__ tiles, const double *__restrict__ in,
                  double *__restrict__ out, NodeScalars *__restrict__ ns, int *__restrict__ tile_iters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  SmallShared &S = *reinterpret_cast<SmallShared *>(smem_raw);
  if (tid == 0) {
    S.tile = tiles[blockIdx.x];
    S.I = insts[S.tile.inst];
    S.remaining = S.tile.nn;
  }
  if (tid < kT) { S.status[tid] = BQP_UNSOLVED; S.iters[tid] = 0; S.newly[tid] = 0; }
  __syncthreads();
  const DevInstance &I = S.I;
  if (I.npad != NP) return;
  constexpr int np = NP, ks = NP / 4;
  const int n = I.n, m = I.m, nn = S.tile.nn, mp = I.s_mp, wa = I.s_wa, wt = I.s_wt;
  unsigned char *blob = smem_raw + align16(sizeof(SmallShared));
  const double *Mf = reinterpret_cast<const double *>(blob), *Pf = reinterpret_cast<const double *>(blob + I.s_offP);
  const double *Av = reinterpret_cast<const double *>(blob + I.s_offAv), *Tv = reinterpret_cast<const double *>(blob + I.s_offTv);
  const double *rho_s = reinterpret_cast<const double *>(blob + I.s_offRho), *rinv_s = reinterpret_cast<const double *>(blob + I.s_offRinv);
  const uint16_t *Ac = reinterpret_cast<const uint16_t *>(blob + I.s_offAc), *Tc = reinterpret_cast<const uint16_t *>(blob + I.s_offTc);
  // vectors, [row][8] with the leaf fastest; every one carries a ZERO ROW at its end (row np / row mp): the padding entries of the
  // ELL rows point there
  const size_t nv = (size_t)(np + 1) * kT, mv = (size_t)(mp + 1) * kT;
  double *sx = reinterpret_cast<double *>(blob + I.s_bytes);
  double *sdx = sx + nv, *sb = sdx + nv, *sxt = sb + nv;
  double *sz = sxt + nv, *sy = sz + mv, *sl = sy + mv, *su = sl + mv, *sdy = su + mv, *sdp = sdy + mv, *sw = sdp + mv;
  double *red = sw + mv;                                               // [kSlots][kWarps][8]
  const double *E_s = reinterpret_cast<const double *>(blob + I.s_offE), *Einv_s = reinterpret_cast<const double *>(blob + I.s_offEinv);
  const double alpha = I.alpha, sigma = I.sigma;

  // ------------------------------------------------------------------ the problem's matrices: TMA bulk copies, one barrier
  const uint32_t bar = smem_u32(&S.bar);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, (uint32_t)I.s_bytes);
    constexpr int kChunk = 32768;
    for (int o = 0; o < I.s_bytes; o += kChunk) {
      const int len = I.s_bytes - o < kChunk ? I.s_bytes - o : kChunk;
      tma_load_1d(smem_u32(blob + o), I.s_blob + o, (uint32_t)len, bar);
    }
  }

  // ------------------------------------------------------------------ prologue (node.py:102-105), overlapping the copy
  const int p = tid & 3, j = tid >> 2;                // node pair; column-space row
  // pairs that hold a leaf: the others' threads skip the vector phases (a tile of 1 or 2 leaves moves a quarter of the
  // shared-memory traffic of a full one -- its warps still issue every instruction, so the host does not narrow the tiles)
  const int npairs = (nn + 1) >> 1;
  const bool act = p < npairs;
  const bool col = j < np, colr = j < n;
  const double qj = colr ? __ldg(I.q + j) : 0.0, dj = colr ? __ldg(I.D + j) : 0.0, dinvj = colr ? __ldg(I.Dinv + j) : 0.0;
  for (int e = tid; e < mp * kT; e += kThreads) {
    const int i = e >> 3, t = e & 7;
    double lo = -kInfty, up = kInfty, yv = 0.0;
    if (i < m && t < nn) {
      const double *q0 = in + S.tile.in_off[t];
      lo = fmax(q0[i], -kInfty);
      up = fmin(q0[m + i], kInfty);
      yv = I.c * __ldg(I.Einv + i) * q0[2 * (size_t)m + n + i];
    }
    const double ei = i < m ? __ldg(I.E + i) : 1.0;
    sl[e] = ei * lo; su[e] = ei * up; sy[e] = yv;
  }
  double2 xr = make_double2(0.0, 0.0);                // this thread's x entries (row j, leaves 2p, 2p + 1)
  if (col) {
    if (colr) {
      if (2 * p < nn) xr.x = dinvj * in[S.tile.in_off[2 * p] + 2 * (size_t)m + j];
      if (2 * p + 1 < nn) xr.y = dinvj * in[S.tile.in_off[2 * p + 1] + 2 * (size_t)m + j];
    }
    *reinterpret_cast<double2 *>(sx + (size_t)j * kT + 2 * p) = xr;
    *reinterpret_cast<double2 *>(sdx + (size_t)j * kT + 2 * p) = make_double2(0.0, 0.0);     // columns of unused pairs stay finite
    *reinterpret_cast<double2 *>(sxt + (size_t)j * kT + 2 * p) = make_double2(0.0, 0.0);
  }
  if (tid < kT) {
    for (double *v : {sx, sdx, sb, sxt}) v[(size_t)np * kT + tid] = 0.0;
    for (double *v : {sz, sy, sl, su, sdy, sdp, sw}) v[(size_t)mp * kT + tid] = 0.0;
  }
  __syncthreads();                                    // barrier initialised, sx / sl / su / sy complete
  mbar_wait(bar, 0);
  const bool mwarp = warp < np / 8;                   // warps that own a row panel of P (and rows of M)
  // The thread's ELL entries stay in registers for the whole launch (config 3: 3 entries per row of A, 5 per row of A'):
  // value + offset of the vector element, padding entries 0.0 * the zero row (reg_dot)
  // (host_setup builds the layout only for problems that fit: wa <= 4, wt <= 8, m <= 192; wider ones stay on bqp_kernels.cu)
  double tv[kRegWT], av[kRegRows][kRegWA], rho3[kRegRows], rinv3[kRegRows];
  uint32_t to[kRegWT], ao[kRegRows][kRegWA];          // byte offsets of the vector elements
  bool rv[kRegRows];
#pragma unroll
  for (int k = 0; k < kRegWT; k++) {
    const bool ok = col && act && k < wt;
    tv[k] = ok ? Tv[k * np + j] : 0.0;
    to[k] = (uint32_t)(((ok ? (int)Tc[k * np + j] : mp) * kT + 2 * p) * 8);
  }
#pragma unroll
  for (int i = 0; i < kRegRows; i++) {
    const int r = j + (kThreads / 4) * i;
    rv[i] = act && r < m;
    rho3[i] = rv[i] ? rho_s[r] : 0.0; rinv3[i] = rv[i] ? rinv_s[r] : 0.0;
#pragma unroll
    for (int k = 0; k < kRegWA; k++) {
      const bool ok = rv[i] && k < wa;
      av[i][k] = ok ? Av[k * mp + r] : 0.0;
      ao[i][k] = (uint32_t)(((ok ? (int)Ac[k * mp + r] : np) * kT + 2 * p) * 8);
    }
  }
#pragma unroll
  for (int i = 0; i < kRegRows; i++) {                 // z = A x ; w = rho z - y
    if (rv[i]) {
      const double2 ax = reg_dot<kRegWA>(av[i], ao[i], sx);
      const size_t e = (size_t)(j + (kThreads / 4) * i) * kT + 2 * p;
      const double2 yv = *reinterpret_cast<const double2 *>(sy + e);
      *reinterpret_cast<double2 *>(sz + e) = ax;
      *reinterpret_cast<double2 *>(sw + e) = make_double2(rho3[i] * ax.x - yv.x, rho3[i] * ax.y - yv.y);
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ ADMM loop (osqp_solve)
  const int max_iter = I.max_iter, check_every = I.check_every;
  int iter = 0;
#ifdef BQP_SMALL_DEBUG
  long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pl = clock64();
#define PSTAMP(i) do { const long long now_ = clock64(); ph[i] += now_ - pl; pl = now_; } while (0)
  long long ch[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; int nchk = 0;
#define CSTAMP(i) do { const long long now_ = clock64(); ch[i] += now_ - pl; pl = now_; } while (0)
#else
#define PSTAMP(i) do { } while (0)
#define CSTAMP(i) do { } while (0)
#endif
  // 32-bit shared addresses of the hot loop
  const uint32_t a_sw = smem_u32(sw), a_sb = smem_u32(sb), a_sxt = smem_u32(sxt), a_sx = smem_u32(sx), a_sdx = smem_u32(sdx),
                 a_sz = smem_u32(sz), a_sy = smem_u32(sy), a_sl = smem_u32(sl), a_su = smem_u32(su), a_sdy = smem_u32(sdy);
  const uint32_t e_col = (uint32_t)((j * kT + 2 * p) * 8);                                  // this thread's element of a column-space vector
  uint32_t e_row[kRegRows];
#pragma unroll
  for (int i = 0; i < kRegRows; i++) e_row[i] = (uint32_t)(((j + (kThreads / 4) * i) * kT + 2 * p) * 8);
  const uint32_t a_bfrag = a_sb + (uint32_t)(((lane & 3) * kT + (lane >> 2)) * 8);            // B fragments of b: + 256 bytes per k-step
  const uint32_t a_mfrag = smem_u32(Mf) + (uint32_t)(((warp * ks) * 32 + lane) * 8);          // A fragments of this warp's panel of M
  // the leaf pair's bounds of the thread's rows are constant over the solve: registers (1.40 -> 1.32 us per iteration)
  double2 lo3[kRegRows], up3[kRegRows];
#pragma unroll
  for (int i = 0; i < kRegRows; i++) {
    lo3[i] = rv[i] ? lds2(a_sl + e_row[i]) : make_double2(0.0, 0.0);
    up3[i] = rv[i] ? lds2(a_su + e_row[i]) : make_double2(0.0, 0.0);
  }
  // ... and z, y of the thread's rows live in registers between the termination checks (1.32 -> 1.25 us): only their owner
  // touches them in an iteration; shared memory sees them at the checks (A'y, the snapshot of a finished leaf)
  double2 z3[kRegRows], y3[kRegRows];
#pragma unroll
  for (int i = 0; i < kRegRows; i++) {
    z3[i] = rv[i] ? lds2(a_sz + e_row[i]) : make_double2(0.0, 0.0);
    y3[i] = rv[i] ? lds2(a_sy + e_row[i]) : make_double2(0.0, 0.0);
  }
  int to_check = check_every;                         // iterations until the next termination check (no division in the loop)
  for (iter = 1; iter <= max_iter; iter++) {
    const bool do_check = (--to_check == 0) || iter == max_iter;
    if (to_check == 0) to_check = check_every;
    PSTAMP(7);
    // b = sigma x - q + A'(rho z - y)
    if (col && act) {
      const double2 aw = reg_dot_s<kRegWT>(tv, to, a_sw);
      sts2(a_sb + e_col, sigma * xr.x - qj + aw.x, sigma * xr.y - qj + aw.y);
    }
    PSTAMP(0);
    __syncthreads();
    PSTAMP(1);
    // x~ = M b on the FP64 tensor pipe: 16 fma per clock and sub-partition, the rate of the plain fma pipe, but for all 8 leaves
    // at once.  M's fragments come from shared memory every iteration (conflict-free 8-byte loads): holding them in registers
    // was measured slower (2.1 instead of 1.5 us per iteration) -- the register file is what this kernel runs out of.  Every
    // operand is loaded before the first mma: a load per mma serialises the chain.  x update in the registers of the owner.
    if (mwarp) {
      double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0}, bv[ks], mf[ks];
#pragma unroll
      for (int s = 0; s < ks; s++) { bv[s] = lds1(a_bfrag + (uint32_t)(s * 4 * kT * 8)); mf[s] = lds1(a_mfrag + (uint32_t)(s * 32 * 8)); }
#pragma unroll
      for (int s = 0; s < ks; s += 2) {
        dmma(c0, mf[s], bv[s]);
        dmma(c1, mf[s + 1], bv[s + 1]);
      }
      if (act) {
        const double2 xt = make_double2(c0[0] + c1[0], c0[1] + c1[1]);
        const double2 xn = make_double2(alpha * xt.x + (1.0 - alpha) * xr.x, alpha * xt.y + (1.0 - alpha) * xr.y);
        sts2(a_sxt + e_col, xt.x, xt.y);
        if (do_check) {
          sts2(a_sdx + e_col, xn.x - xr.x, xn.y - xr.y);
          sts2(a_sx + e_col, xn.x, xn.y);
        }
        xr = xn;
      }
    }
    PSTAMP(2);
    __syncthreads();
    PSTAMP(3);
    // z~ = A x~, projection, dual update, next w: the thread's rows j, j + 64, j + 128 together (loads first, then the updates)
    if (act) {
      double2 zt[kRegRows], zp[kRegRows], yv[kRegRows];
#pragma unroll
      for (int i = 0; i < kRegRows; i++) {
        zt[i] = reg_dot_s<kRegWA>(av[i], ao[i], a_sxt);       // padding rows: 0.0 * the zero row
        zp[i] = z3[i]; yv[i] = y3[i];
      }
#pragma unroll
      for (int i = 0; i < kRegRows; i++) {
        if (rv[i]) {
          const double zr0 = alpha * zt[i].x + (1.0 - alpha) * zp[i].x, zr1 = alpha * zt[i].y + (1.0 - alpha) * zp[i].y;
          // osqp's c_max / c_min are compare-and-select macros; FP64 fmax / fmin compile to a DSETP + SEL + FSEL + LOP3 sequence
          // each (measured: 1.49 -> 1.40 us per iteration).  Same values for every comparable pair of operands
          const double t0 = zr0 + rinv3[i] * yv[i].x, t1 = zr1 + rinv3[i] * yv[i].y;
          const double m0 = t0 > lo3[i].x ? t0 : lo3[i].x, m1 = t1 > lo3[i].y ? t1 : lo3[i].y;
          const double zn0 = m0 < up3[i].x ? m0 : up3[i].x, zn1 = m1 < up3[i].y ? m1 : up3[i].y;
          const double dy0 = rho3[i] * (zr0 - zn0), dy1 = rho3[i] * (zr1 - zn1);
          const double yn0 = yv[i].x + dy0, yn1 = yv[i].y + dy1;
          z3[i] = make_double2(zn0, zn1); y3[i] = make_double2(yn0, yn1);
          if (do_check) { sts2(a_sz + e_row[i], zn0, zn1); sts2(a_sy + e_row[i], yn0, yn1); }      // other threads read y at the checks only
          sts2(a_sw + e_row[i], rho3[i] * zn0 - yn0, rho3[i] * zn1 - yn1);
          if (do_check) sts2(a_sdy + e_row[i], dy0, dy1);
        }
      }
    }
    PSTAMP(4);
    __syncthreads();
    PSTAMP(5);
    if (!do_check) continue;

    // -------------------------------------------------------------- termination check (update_info + check_termination)
    {
      // column space: P x, A' y -> dual residual and its norms, objective; dx norms
      double dr0 = 0, dr1 = 0, b10 = 0, b11 = 0, b20 = 0, b21 = 0, qd0 = 0, qd1 = 0, ln0 = 0, ln1 = 0, ndx0 = 0, ndx1 = 0, qdx0 = 0, qdx1 = 0;
      if (mwarp) {
        const double2 px = frag_rows_smem(Pf, ks, warp, lane, sx);
        const double2 aty = reg_dot<kRegWT>(tv, to, sy);
        if (colr && act) {
          dr0 = fabs(dinvj * (px.x + qj + aty.x)); dr1 = fabs(dinvj * (px.y + qj + aty.y));
          b10 = fabs(dinvj * px.x); b11 = fabs(dinvj * px.y);
          b20 = fabs(dinvj * aty.x); b21 = fabs(dinvj * aty.y);
          qd0 = xr.x * px.x; qd1 = xr.y * px.y;
          ln0 = qj * xr.x; ln1 = qj * xr.y;
          const double2 d = *reinterpret_cast<const double2 *>(sdx + (size_t)j * kT + 2 * p);
          ndx0 = fabs(dj * d.x); ndx1 = fabs(dj * d.y);
          qdx0 = qj * d.x; qdx1 = qj * d.y;
        }
      }
      CSTAMP(0);
      red_put<0>(dr0, dr1, red, 0, warp, lane);
      red_put<0>(b10, b11, red, 1, warp, lane);
      red_put<0>(b20, b21, red, 2, warp, lane);
      red_put<1>(qd0, qd1, red, 3, warp, lane);
      red_put<1>(ln0, ln1, red, 4, warp, lane);
      red_put<0>(ndx0, ndx1, red, 10, warp, lane);
      red_put<1>(qdx0, qdx1, red, 11, warp, lane);
      CSTAMP(1);
      // row space: A x -> primal residual and its norms; projected dy (certificate of primal infeasibility)
      double pr0 = 0, pr1 = 0, a10 = 0, a11 = 0, a20 = 0, a21 = 0, ndy0 = 0, ndy1 = 0, lh0 = 0, lh1 = 0;
#pragma unroll
      for (int i = 0; i < kRegRows; i++) if (rv[i]) {
        const int r = j + (kThreads / 4) * i;
        const double2 ax = reg_dot<kRegWA>(av[i], ao[i], sx);
        const size_t e = (size_t)r * kT + 2 * p;
        const double einv = Einv_s[r], ei = E_s[r];
        const double2 zv = *reinterpret_cast<const double2 *>(sz + e);
        const double2 lo = *reinterpret_cast<const double2 *>(sl + e), up = *reinterpret_cast<const double2 *>(su + e);
        double2 d = *reinterpret_cast<const double2 *>(sdy + e);
        pr0 = fmax(pr0, fabs(einv * (ax.x - zv.x))); pr1 = fmax(pr1, fabs(einv * (ax.y - zv.y)));
        a10 = fmax(a10, fabs(einv * ax.x)); a11 = fmax(a11, fabs(einv * ax.y));
        a20 = fmax(a20, fabs(einv * zv.x)); a21 = fmax(a21, fabs(einv * zv.y));
        if (up.x > kInfty * kMinScaling) { if (lo.x < -kInfty * kMinScaling) d.x = 0.0; else d.x = fmin(d.x, 0.0); }
        else if (lo.x < -kInfty * kMinScaling) d.x = fmax(d.x, 0.0);
        if (up.y > kInfty * kMinScaling) { if (lo.y < -kInfty * kMinScaling) d.y = 0.0; else d.y = fmin(d.y, 0.0); }
        else if (lo.y < -kInfty * kMinScaling) d.y = fmax(d.y, 0.0);
        *reinterpret_cast<double2 *>(sdp + e) = d;
        ndy0 = fmax(ndy0, fabs(ei * d.x)); ndy1 = fmax(ndy1, fabs(ei * d.y));
        lh0 += up.x * fmax(d.x, 0.0) + lo.x * fmin(d.x, 0.0); lh1 += up.y * fmax(d.y, 0.0) + lo.y * fmin(d.y, 0.0);
      }
      CSTAMP(2);
      red_put<0>(pr0, pr1, red, 5, warp, lane);
      red_put<0>(a10, a11, red, 6, warp, lane);
      red_put<0>(a20, a21, red, 7, warp, lane);
      red_put<0>(ndy0, ndy1, red, 8, warp, lane);
      red_put<1>(lh0, lh1, red, 9, warp, lane);
    }
    CSTAMP(3);
    __syncthreads();                                  // projected dy complete
    {
      double t10 = 0, t11 = 0, t20 = 0, t21 = 0;
      if (mwarp) {
        const double2 atd = reg_dot<kRegWT>(tv, to, sdp);
        const double2 pdx = frag_rows_smem(Pf, ks, warp, lane, sdx);
        if (colr && act) {
          t10 = fabs(dinvj * atd.x); t11 = fabs(dinvj * atd.y);
          t20 = fabs(dinvj * pdx.x); t21 = fabs(dinvj * pdx.y);
        }
      }
      CSTAMP(4);
      red_put<0>(t10, t11, red, 12, warp, lane);
      red_put<0>(t20, t21, red, 13, warp, lane);
      double vu0 = -INFINITY, vu1 = -INFINITY, vl0 = INFINITY, vl1 = INFINITY;
#pragma unroll
      for (int i = 0; i < kRegRows; i++) if (rv[i]) {
        const int r = j + (kThreads / 4) * i;
        const double2 adx = reg_dot<kRegWA>(av[i], ao[i], sdx);
        const size_t e = (size_t)r * kT + 2 * p;
        const double einv = Einv_s[r];
        const double2 lo = *reinterpret_cast<const double2 *>(sl + e), up = *reinterpret_cast<const double2 *>(su + e);
        const double v0 = einv * adx.x, v1 = einv * adx.y;
        if (up.x < kInfty * kMinScaling) vu0 = fmax(vu0, v0);
        if (up.y < kInfty * kMinScaling) vu1 = fmax(vu1, v1);
        if (lo.x > -kInfty * kMinScaling) vl0 = fmin(vl0, v0);
        if (lo.y > -kInfty * kMinScaling) vl1 = fmin(vl1, v1);
      }
      CSTAMP(5);
      red_put<0>(vu0, vu1, red, 14, warp, lane);
      red_put<2>(vl0, vl1, red, 15, warp, lane);
    }
    __syncthreads();
    CSTAMP(6);
    if (tid < kSlots * kT) {                          // combine the per-warp partials in warp order
      const int slot = tid >> 3, t = tid & 7;
      const bool is_sum = (slot == 3 || slot == 4 || slot == 9 || slot == 11), is_min = (slot == 15);
      double r = red[((size_t)slot * kWarps) * kT + t];
      for (int w = 1; w < kWarps; w++) {
        const double v = red[((size_t)slot * kWarps + w) * kT + t];
        r = is_sum ? r + v : (is_min ? fmin(r, v) : fmax(r, v));
      }
      S.fin[slot][t] = r;
    }
    __syncthreads();
    if (tid < kT) {
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
    CSTAMP(7);
    // snapshot the iterates of nodes that just terminated (unscaled; NaN for certificates, as osqp returns)
    for (int t = 0; t < nn; t++) {
      if (!S.newly[t]) continue;
      const int st = S.status[t];
      const bool bad = !(st == BQP_SOLVED || st == BQP_SOLVED_INACCURATE || st == BQP_MAX_ITER_REACHED);
      double *ox = out + S.tile.out_off[t], *oy = ox + n;
      for (int c = tid; c < n; c += kThreads) ox[c] = bad ? NAN : __ldg(I.D + c) * sx[(size_t)c * kT + t];
      for (int i = tid; i < m; i += kThreads) oy[i] = bad ? NAN : I.cinv * __ldg(I.E + i) * sy[(size_t)i * kT + t];
    }
    CSTAMP(8);
#ifdef BQP_SMALL_DEBUG
    nchk++;
#endif
    if (S.remaining == 0) break;
  }
#ifdef BQP_SMALL_DEBUG
  if (blockIdx.x == 0 && (tid == 0 || tid == 255)) {
    const double it = (double)(iter > max_iter ? max_iter : iter);
    printf("small kernel thread %d, %d iterations, clk per iteration: A' %.0f | wait %.0f | M b %.0f | wait %.0f | A + update %.0f | wait %.0f | checks (total) %lld | loop overhead %.0f\n",
           tid, (int)it, ph[0] / it, ph[1] / it, ph[2] / it, ph[3] / it, ph[4] / it, ph[5] / it, ph[6], ph[7] / it);
    printf("  %d checks, clk per check: P x + A'y + column stats %lld | their reductions %lld | A x + row stats %lld | their reductions + wait %lld | "
           "A'dy + P dx %lld | A dx + reductions %lld | wait %lld | combine + decision %lld | snapshot %lld\n", nchk, ch[0] / nchk, ch[1] / nchk, ch[2] / nchk,
           ch[3] / nchk, ch[4] / nchk, ch[5] / nchk, ch[6] / nchk, ch[7] / nchk, ch[8] / nchk);
  }
#endif
  __syncthreads();
  if (tid == 0) tile_iters[blockIdx.x] = iter > max_iter ? max_iter : iter;

  // ------------------------------------------------------------------ epilogue (node.py:128-143)
  // clip integer entries into the node's own bounds, then lower = 1/2 x'Px + q'x at the clipped point
  for (int t = 0; t < nn; t++) {
    const int st = S.status[t];
    if (!(st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED)) continue;
    double *ox = out + S.tile.out_off[t];
    const double *q0 = in + S.tile.in_off[t];
    for (int k = tid; k < I.n_int; k += kThreads) {
      const int c = __ldg(I.i_idx + k), row = m - I.n_int + k;
      ox[c] = fmin(fmax(ox[c], q0[row]), q0[m + row]);
    }
  }
  __syncthreads();
  double2 xo = make_double2(0.0, 0.0);
  if (col) {
    if (colr) {
      const int s0 = 2 * p < nn ? S.status[2 * p] : BQP_UNSOLVED, s1 = 2 * p + 1 < nn ? S.status[2 * p + 1] : BQP_UNSOLVED;
      if (s0 == BQP_SOLVED || s0 == BQP_MAX_ITER_REACHED) xo.x = dinvj * out[S.tile.out_off[2 * p] + j];
      if (s1 == BQP_SOLVED || s1 == BQP_MAX_ITER_REACHED) xo.y = dinvj * out[S.tile.out_off[2 * p + 1] + j];
    }
    *reinterpret_cast<double2 *>(sxt + (size_t)j * kT + 2 * p) = xo;
  }
  __syncthreads();
  {
    double qd0 = 0, qd1 = 0, ln0 = 0, ln1 = 0;
    if (mwarp) {
      const double2 px = frag_rows_smem(Pf, ks, warp, lane, sxt);
      if (colr) { qd0 = xo.x * px.x; qd1 = xo.y * px.y; ln0 = qj * xo.x; ln1 = qj * xo.y; }
    }
    red_put<1>(qd0, qd1, red, 0, warp, lane);
    red_put<1>(ln0, ln1, red, 1, warp, lane);
  }
  __syncthreads();
  if (tid < nn) {
    const int t = tid, st = S.status[t];
    if (st == BQP_SOLVED || st == BQP_MAX_ITER_REACHED) {
      double qd = 0, ln = 0;
      for (int w = 0; w < kWarps; w++) { qd += red[((size_t)0 * kWarps + w) * kT + t]; ln += red[((size_t)1 * kWarps + w) * kT + t]; }
      ns[S.tile.node[t]].lower = (0.5 * qd + ln) * I.cinv;
    }
  }
}

}  // namespace

size_t small_smem_bytes(int npad, int m, int blob_bytes) {
  const int mp = m + 7 > 8 ? (m + 7) / 8 * 8 : 8;
  return align16(sizeof(SmallShared)) + (size_t)blob_bytes + 8 * ((size_t)4 * (npad + 1) * kT + (size_t)7 * (mp + 1) * kT + (size_t)kSlots * kWarps * kT);
}

int launch_admm_small(int npad_mask, const DevInstance *d_insts, const DevTile *d_tiles, int ntiles, const double *d_in, double *d_out,
                      NodeScalars *d_ns, int *d_tile_iters, size_t smem_bytes, void *stream) {
  // npad_mask: bit 0 -- some problem of the launch has npad = 32, bit 1 -- npad = 64.  One launch per width over ALL tiles: a CTA
  // whose problem has the other width leaves at once
  if (!(npad_mask & 3) || smem_bytes > (size_t)kMaxSmem) return BQP_E_ARG;
  // many host threads launch concurrently (one context each): raise the attribute once
  static std::atomic<int> attr_set{0};
  if (!attr_set.load()) {
    if (cudaFuncSetAttribute(admm_small_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem) != cudaSuccess) return BQP_E_CUDA;
    if (cudaFuncSetAttribute(admm_small_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem) != cudaSuccess) return BQP_E_CUDA;
    attr_set.store(1);
  }
  if (npad_mask & 1) admm_small_kernel<32><<<ntiles, kThreads, smem_bytes, (cudaStream_t)stream>>>(d_insts, d_tiles, d_in, d_out, d_ns, d_tile_iters);
  if (npad_mask & 2) admm_small_kernel<64><<<ntiles, kThreads, smem_bytes, (cudaStream_t)stream>>>(d_insts, d_tiles, d_in, d_out, d_ns, d_tile_iters);
  return cudaGetLastError() == cudaSuccess ? BQP_OK : BQP_E_CUDA;
}

}  // namespace bqp
