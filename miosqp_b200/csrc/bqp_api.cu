// bqp_api.cu -- the C ABI of include/bqp.h: device residency of set-up problems, frontier batching
// (tile construction, pinned staging, H2D / launch / D2H on one stream, CUDA-event timing).
// There is no CPU solve path in this library: without a usable CUDA device setup and solve fail.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <atomic>
#include <thread>
#include <vector>

#include "bqp_internal.h"

using namespace bqp;

struct bqp_instance {
  HostInstance h;
  bool on_device = false;
  int device = 0;
  DevInstance d{};
  std::vector<void *> allocs;
  double *d_q = nullptr;
};

namespace {

#define CK(call)                                   \
  do {                                             \
    cudaError_t e_ = (call);                       \
    if (e_ != cudaSuccess) {                       \
      g_last_cuda = e_;                            \
      return e_ == cudaErrorMemoryAllocation ? BQP_E_ALLOC : BQP_E_CUDA; \
    }                                              \
  } while (0)

thread_local cudaError_t g_last_cuda = cudaSuccess;
int g_tune_tt = 0, g_tune_threads = 0;

// Device image of one set-up problem: every array goes into ONE allocation through ONE host-to-device copy (round 1 made
// a cudaMalloc + cudaMemcpy per array, ~25 per instance: 2 500 synchronous driver calls for the 100 instances of config 2).
struct Arena {
  struct Item { const void *src; size_t bytes, off; const void **out; };
  std::vector<Item> items;
  size_t total = 0;
  template <class Tp>
  void add(const std::vector<Tp> &v, const Tp **out) {
    const size_t bytes = v.size() * sizeof(Tp);
    items.push_back({v.empty() ? nullptr : (const void *)v.data(), bytes, total, (const void **)out});
    total += (std::max<size_t>(bytes, 1) + 255) & ~size_t(255);      // 256-byte aligned: TMA sources need 16
  }
  void add_mat(const HostMat &M, DevMat *D) {
    D->rows = M.rows; D->cols = M.cols; D->nslices = M.nslices;
    add(M.sptr, &D->sptr); add(M.iptr, &D->iptr); add(M.vals, &D->vals); add(M.idx, &D->idx);
  }
  int commit(bqp_instance *inst) {
    void *base = nullptr;
    CK(cudaMalloc(&base, std::max<size_t>(total, 256)));
    inst->allocs.push_back(base);
    std::vector<unsigned char> stage(total);
    for (auto &it : items) if (it.bytes) std::memcpy(stage.data() + it.off, it.src, it.bytes);
    if (total) CK(cudaMemcpy(base, stage.data(), total, cudaMemcpyHostToDevice));
    for (auto &it : items) *it.out = (const unsigned char *)base + it.off;
    return BQP_OK;
  }
};

// grow-only buffers of a batch context.  Device buffers come from the stream-ordered allocator (cudaMallocAsync on the
// context's stream): cudaMalloc / cudaFree synchronise the whole device, which with many contexts in flight made every
// growing buffer wait for every other context's running kernel (measured: 134 s instead of 5 s for 100 asynchronous trees).
// Pinned host buffers start at 256 KiB and double, so they are reallocated a handful of times per context at most.
struct Buf {
  void *p = nullptr; size_t cap = 0; bool pinned = false;
  int reserve(size_t bytes, cudaStream_t st) {
    if (bytes <= cap) return BQP_OK;
    const size_t want = std::max(std::max(bytes, cap * 2), (size_t)256 * 1024);
    release(st);
    if (pinned) CK(cudaHostAlloc(&p, want, cudaHostAllocDefault)); else CK(cudaMallocAsync(&p, want, st));
    cap = want;
    return BQP_OK;
  }
  // grow keeping the first `used` bytes (rolling sessions append nodes to resident buffers)
  int reserve_keep(size_t bytes, size_t used, cudaStream_t st) {
    if (bytes <= cap) return BQP_OK;
    const size_t want = std::max(std::max(bytes, cap * 2), (size_t)256 * 1024);
    void *q = nullptr;
    if (pinned) CK(cudaHostAlloc(&q, want, cudaHostAllocDefault)); else CK(cudaMallocAsync(&q, want, st));
    if (p && used) {
      if (pinned) { CK(cudaStreamSynchronize(st)); std::memcpy(q, p, used); }
      else CK(cudaMemcpyAsync(q, p, used, cudaMemcpyDeviceToDevice, st));
    }
    release(st);
    p = q; cap = want;
    return BQP_OK;
  }
  void release(cudaStream_t st) {
    if (p) { if (pinned) cudaFreeHost(p); else cudaFreeAsync(p, st); }
    p = nullptr; cap = 0;
  }
};

struct BatchCtx {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  // bqp_solve_multi on a kernel that finishes in ONE launch (no rounds: direct-load, shared-memory-resident and whole-GPU
  // kernels -- the latency-bound B&B drivers of configs 3 and 4): upload, tile table, launch and every read-back are queued
  // back to back and the host waits once, instead of after each of the five steps (BQP_FAST_PATH=0: the stepwise path)
  bool want_fast = false, fast = false, out_fetched = false;
  std::vector<DevInstance> dinst_host;
  Buf h_in{nullptr, 0, true}, h_out{nullptr, 0, true}, h_ns{nullptr, 0, true}, h_ti{nullptr, 0, true};
  Buf d_in, d_out, d_ns, d_ti, d_work, d_tiles, d_insts, d_gbar;
  // description of the resident batch
  int B = 0, ntiles = 0, tt = 0, threads = 0;
  bool use_grid = false; int grid_ctas = 0;
  int small_mask = 0;             // widths (npad 32 / 64) among the problems of the planned launch: one kernel instantiation each
  bool use_small = false;         // every problem of the batch has the shared-memory-resident layout (npad <= 64) and no dense kernel serves it
  bool rows_ext = false;          // some problem of the batch uses eq_rho == 2 or adaptive rho: the rows kernel's extended instantiation
  bool use_stream = false, use_panel = false, use_rows = false, w_in_stage = false, round_ok = true; int nslots = 0, slot_bytes = 0, stage_bytes = 0, nw_max = 0, cs = 1;
  int round_iters = 0, max_iter_all = 0;
  long long round_h2d_bytes = 0, round_h2d_total = 0;
  std::vector<long long> in_off, state_off, corr_off;
  Buf d_state, d_corr;
  size_t corr_d = 0;             // doubles used in d_corr
  size_t smem = 0, in_doubles = 0, out_doubles = 0;
  std::vector<bqp_instance *> node_inst;
  std::vector<long long> out_off;
  std::vector<DevTile> tiles;
  std::vector<long long> tile_bytes_iter, tile_bytes_check, tile_bytes_launch;
  std::vector<int> tile_check_every;
  bqp_timing timing{};
  bool resident = false, ran = false;
  // rolling session (bqp_session_*): nodes are appended while others are still iterating
  bool eq2_unsupported = false;   // a problem set up with eq_rho == 2 would run on a kernel without the Woodbury correction
  bool session = false;
  std::vector<int> s_alive, s_progress; std::vector<double> s_dist, s_remaining;
  size_t s_in_d = 0, s_out_d = 0, s_st_d = 0;
  long long s_tile_iters = 0, s_bytes = 0; int s_launches = 0; double s_kernel_ms = 0;
  int sm_parts = 1;                    // this context plans for 1 / sm_parts of the SMs (other contexts are busy beside it)
  bool auto_cluster_default = false;   // what closed batches (bqp_solve_multi on this context) use; bqp_ctx_set_auto_cluster
  bool auto_cluster = false;    // rows kernel: clusters of 4 when the round holds few tiles (results then depend on the schedule in the last bits)
  int round_override = -1;      // >= 0: rounds of that many ADMM iterations (0 = run every tile to completion) whatever BQP_ROUND_ITERS says
  bool blocking_sync = false;   // wait on a blocking event instead of spinning (many contexts driven by many host threads)
  cudaEvent_t ev_block = nullptr;
};
BatchCtx g0;                    // the context behind the handle-less entry points (bqp_solve_multi, bqp_batch_*)
std::mutex g_ctx_mu;
std::vector<BatchCtx *> g_ctxs{&g0};   // every live context: bqp_free / bqp_update_q must not pull data from under a running one

// wait for everything queued on the context's stream
cudaError_t ctx_sync(BatchCtx &g) {
  if (!g.blocking_sync) return cudaStreamSynchronize(g.stream);
  cudaError_t e = cudaEventRecord(g.ev_block, g.stream);
  return e != cudaSuccess ? e : cudaEventSynchronize(g.ev_block);
}

void ctx_release(BatchCtx &g) {
  if (!g.stream) return;
  cudaSetDevice(g.device);
  cudaStreamSynchronize(g.stream);
  for (Buf *b : {&g.h_in, &g.h_out, &g.h_ns, &g.h_ti, &g.d_in, &g.d_out, &g.d_ns, &g.d_ti, &g.d_work, &g.d_tiles, &g.d_insts, &g.d_gbar, &g.d_state, &g.d_corr}) b->release(g.stream);
  cudaStreamSynchronize(g.stream);
  for (auto &e : g.ev) { cudaEventDestroy(e); e = nullptr; }
  if (g.ev_block) { cudaEventDestroy(g.ev_block); g.ev_block = nullptr; }
  cudaStreamDestroy(g.stream); g.stream = nullptr;
  g.resident = g.ran = false;
}

int ctx_init(BatchCtx &g, int device) {
  if (g.device == device && g.stream) return BQP_OK;
  ctx_release(g);   // switching devices: drop everything
  CK(cudaSetDevice(device));
  CK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  for (auto &e : g.ev) CK(cudaEventCreate(&e));
  CK(cudaEventCreateWithFlags(&g.ev_block, cudaEventBlockingSync | cudaEventDisableTiming));
  g.device = device;
  return BQP_OK;
}

int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

int to_device(bqp_instance *inst) {
  const HostInstance &h = inst->h;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || h.s.device < 0 || h.s.device >= ndev) return BQP_E_CUDA;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h.s.device));
  if (prop.major < 10) return BQP_E_CUDA;   // sm_100a code only
  inst->device = h.s.device;
  CK(cudaSetDevice(h.s.device));
  DevInstance &d = inst->d;
  d.n = h.n; d.m = h.m; d.npad = h.npad; d.n_int = h.n_int;
  Arena ar;
  ar.add_mat(h.At, &d.At); ar.add_mat(h.Ab, &d.Ab); ar.add_mat(h.Pm, &d.Pm);
  ar.add(h.Lcol, &d.Lcol); ar.add(h.Lrow, &d.Lrow); ar.add(h.D2inv, &d.D2inv);
  ar.add(h.rho, &d.rho); ar.add(h.rho_inv, &d.rho_inv); ar.add(h.q, &d.q);
  ar.add(h.D, &d.D); ar.add(h.Dinv, &d.Dinv); ar.add(h.E, &d.E); ar.add(h.Einv, &d.Einv); ar.add(h.i_idx, &d.i_idx);
  d.stream = nullptr; d.groups = nullptr; d.w_in_stage = 0;
  if (h.st.built) {
    ar.add(h.st.data, &d.stream); ar.add(h.st.groups, &d.groups);
    const HostStream &st = h.st;
    d.g_at[0] = st.range[GK_AT][0]; d.g_at[1] = st.range[GK_AT][1];
    d.g_fw[0] = st.fw[0]; d.g_fw[1] = st.fw[1];
    d.g_bw[0] = st.bw[0]; d.g_bw[1] = st.bw[1];
    d.g_ab[0] = st.range[GK_AB][0]; d.g_ab[1] = st.range[GK_AB][1];
    d.g_pm[0] = st.range[GK_PM][0]; d.g_pm[1] = st.range[GK_PM][1];
    d.w_in_stage = 1;
    for (int g = d.g_at[0]; g < d.g_at[1]; g++) if (st.groups[g].sparse) d.w_in_stage = 0;
    if (const char *e = std::getenv("BQP_W_IN_STAGE")) if (std::atoi(e) == 0) d.w_in_stage = 0;
  }
  d.pstream = nullptr; d.p_nw = d.p_npm = d.p_npa = 0; d.p_panel_doubles = d.p_offA = d.p_offP = 0;
  if (h.pn.built) {
    ar.add(h.pn.data, &d.pstream);
    d.p_nw = h.pn.nw; d.p_npm = h.pn.npm; d.p_npa = h.pn.npa;
    d.p_panel_doubles = h.pn.panel_doubles; d.p_offA = h.pn.offA; d.p_offP = h.pn.offP; d.p_offV = h.pn.offV;
  }
  d.g_M = d.g_P = nullptr; d.g_arp = d.g_aci = d.g_trp = d.g_tci = nullptr; d.g_avl = d.g_tvl = nullptr; d.g_npm = 0;
  d.g_V = d.g_mu = nullptr; d.g_rtype = nullptr; d.adaptive = 0; d.adapt_interval = 0; d.adapt_tol = 5.0;
  const double *g_mp = nullptr;
  if (h.gd.built) {
    ar.add(h.gd.MP, &g_mp);
    ar.add(h.gd.arp, &d.g_arp); ar.add(h.gd.aci, &d.g_aci); ar.add(h.gd.avl, &d.g_avl);
    ar.add(h.gd.trp, &d.g_trp); ar.add(h.gd.tci, &d.g_tci); ar.add(h.gd.tvl, &d.g_tvl);
    d.g_npm = h.gd.npm;
    if (h.gd.spectral) {
      ar.add(h.gd.mu, &d.g_mu); ar.add(h.rtype, &d.g_rtype);
      d.adaptive = 1; d.adapt_interval = h.s.adaptive_rho_interval; d.adapt_tol = h.s.adaptive_rho_tolerance;
    }
  }
  if (h.pn.built && h.pn.spectral) {      // rows kernel, adaptive rho
    ar.add(h.pn.mu, &d.g_mu); ar.add(h.rtype, &d.g_rtype);
    d.adaptive = 1; d.adapt_interval = h.s.adaptive_rho_interval; d.adapt_tol = h.s.adaptive_rho_tolerance;
  }
  d.s_blob = nullptr; d.s_bytes = d.s_wa = d.s_wt = d.s_mp = d.s_offP = d.s_offAv = d.s_offTv = d.s_offRho = d.s_offRinv = d.s_offE = d.s_offEinv = d.s_offAc = d.s_offTc = 0;
  if (h.sm.built) {
    ar.add(h.sm.blob, &d.s_blob);
    d.s_bytes = h.sm.bytes; d.s_wa = h.sm.wa; d.s_wt = h.sm.wt; d.s_mp = h.sm.mp;
    d.s_offP = h.sm.offP; d.s_offAv = h.sm.offAv; d.s_offTv = h.sm.offTv; d.s_offRho = h.sm.offRho; d.s_offRinv = h.sm.offRinv; d.s_offE = h.sm.offE; d.s_offEinv = h.sm.offEinv;
    d.s_offAc = h.sm.offAc; d.s_offTc = h.sm.offTc;
  }
  d.p_mint = nullptr; d.eq2 = h.s.eq_rho == 2 ? 1 : 0; d.rho_base = h.s.rho;
  if (d.eq2) ar.add(h.mint, &d.p_mint);
  int rc = ar.commit(inst);
  if (rc) return rc;
  inst->d_q = const_cast<double *>(d.q);
  if (h.gd.built) { d.g_M = g_mp; d.g_P = g_mp + h.gd.offP; d.g_V = h.gd.spectral ? g_mp + h.gd.offV : nullptr; }
  d.c = h.c; d.cinv = h.cinv; d.nq = h.nq;
  d.sigma = h.s.sigma; d.alpha = h.s.alpha; d.eps_abs = h.s.eps_abs; d.eps_rel = h.s.eps_rel;
  d.eps_pinf = h.s.eps_prim_inf; d.eps_dinf = h.s.eps_dual_inf;
  d.max_iter = h.s.max_iter; d.check_every = h.s.check_termination;
  inst->on_device = true;
  return BQP_OK;
}

int max_tile_nodes(const HostInstance &h, int threads) {
  int tt = kMaxTT;
  while (tt > 1 && tile_smem_bytes(h.n, h.m, tt, threads) > (size_t)kMaxSmem) tt >>= 1;
  return tile_smem_bytes(h.n, h.m, tt, threads) <= (size_t)kMaxSmem ? tt : 0;
}

}  // namespace

extern "C" {

void bqp_default_settings(bqp_settings *s) {
  if (!s) return;
  s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6; s->eps_abs = 1e-3; s->eps_rel = 1e-3;
  s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4; s->max_iter = 4000; s->scaling = 10;
  s->check_termination = 25; s->eq_rho = 1; s->device = 0;
  s->adaptive_rho = 0; s->adaptive_rho_interval = 0; s->adaptive_rho_tolerance = 5.0;
}

int bqp_debug_host_setup(const bqp_problem *p, const bqp_settings *s, bqp_handle *out) {
  if (!out) return BQP_E_ARG;
  *out = nullptr;
  bqp_instance *inst = new (std::nothrow) bqp_instance();
  if (!inst) return BQP_E_ALLOC;
  int rc;
  try { rc = host_setup(p, s, &inst->h); } catch (const std::bad_alloc &) { rc = BQP_E_ALLOC; }
  if (rc) { delete inst; return rc; }
  *out = inst;
  return BQP_OK;
}

int bqp_setup(const bqp_problem *p, const bqp_settings *s, bqp_handle *out) {
  if (!out) return BQP_E_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { *out = nullptr; return BQP_E_CUDA; }
  int rc = bqp_debug_host_setup(p, s, out);
  if (rc) return rc;
  if (max_tile_nodes((*out)->h, 64) == 0) { bqp_free(*out); *out = nullptr; return BQP_E_UNSUPPORTED; }
  rc = to_device(*out);
  if (rc) { bqp_free(*out); *out = nullptr; }
  return rc;
}

// Many problems at once (100 instances of BASELINE config 2): the host halves -- scaling, LDL', explicit reduced inverse,
// streamed layouts; ~0.2 s each at n = 500 and independent of one another -- run on `threads` host threads (0 = hardware
// concurrency); the device uploads then go one by one through the same to_device() as bqp_setup.  All or nothing.
int bqp_setup_many(int count, const bqp_problem *const *p, const bqp_settings *s, bqp_handle *out, int threads, int host_only) {
  if (count <= 0 || !p || !s || !out) return BQP_E_ARG;
  for (int k = 0; k < count; k++) out[k] = nullptr;
  if (!host_only) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return BQP_E_CUDA;
  }
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, count));
  std::vector<int> rcs((size_t)count, BQP_OK);
  std::atomic<int> next(0);
  auto worker = [&]() {
    for (;;) {
      const int k = next.fetch_add(1);
      if (k >= count) break;
      rcs[(size_t)k] = bqp_debug_host_setup(p[k], s, &out[k]);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; t++) pool.emplace_back(worker);
  worker();
  for (auto &th : pool) th.join();
  int rc = BQP_OK;
  for (int k = 0; k < count && !rc; k++) rc = rcs[(size_t)k];
  for (int k = 0; k < count && !rc && !host_only; k++) {
    if (max_tile_nodes(out[k]->h, 64) == 0) rc = BQP_E_UNSUPPORTED;
    else rc = to_device(out[k]);
  }
  if (rc) for (int k = 0; k < count; k++) { bqp_free(out[k]); out[k] = nullptr; }
  return rc;
}

int bqp_free(bqp_handle h) {
  if (!h) return BQP_OK;
  if (!h->allocs.empty()) {      // also after a partial upload that failed half way (on_device is set last)
    cudaSetDevice(h->device);
    {
      std::lock_guard<std::mutex> lk(g_ctx_mu);
      for (BatchCtx *c : g_ctxs) {
        if (c->stream && c->device == h->device) cudaStreamSynchronize(c->stream);
        for (auto &ni : c->node_inst) if (ni == h) { c->resident = false; break; }
      }
    }
    for (void *p : h->allocs) cudaFree(p);
  }
  delete h;
  return BQP_OK;
}

int bqp_update_q(bqp_handle h, const double *q) {
  if (!h || !q) return BQP_E_ARG;
  host_rescale_q(&h->h, q);
  h->d.nq = h->h.nq;
  if (h->on_device) {
    CK(cudaSetDevice(h->device));
    {
      std::lock_guard<std::mutex> lk(g_ctx_mu);
      for (BatchCtx *c : g_ctxs) if (c->stream && c->device == h->device) CK(cudaStreamSynchronize(c->stream));
    }
    CK(cudaMemcpy(h->d_q, h->h.q.data(), sizeof(double) * h->h.n, cudaMemcpyHostToDevice));
  }
  return BQP_OK;
}

int bqp_set_tuning(int tile_nodes, int threads) {
  if (tile_nodes != 0 && tile_nodes != 1 && tile_nodes != 2 && tile_nodes != 4 && tile_nodes != 8) return BQP_E_ARG;
  if (threads != 0 && (threads % 32 != 0 || threads < 32 || threads > kMaxThreads)) return BQP_E_ARG;
  g_tune_tt = tile_nodes; g_tune_threads = threads;
  return BQP_OK;
}

// ---- per-launch planning: tile (a capacity-bounded subset of) the still-running nodes, choose T / ring depth, upload
// the tile table.  Nodes are grouped by (problem, iterations done so far); groups with the least progress go first, and
// when more tiles exist than the GPU holds at once (`capacity`) the rest wait for the next launch -- so every launch is
// one full wave, and stragglers are re-tiled ever narrower as the frontier drains.  `scheduled` returns the chosen nodes.
static int plan_round(BatchCtx &g, const std::vector<int> &alive, const std::vector<int> &progress, const std::vector<double> &remaining,
                      std::vector<int> *scheduled) {
  int ndev_sms = 148;
  cudaDeviceGetAttribute(&ndev_sms, cudaDevAttrMultiProcessorCount, g.device);
  const int all_sms = ndev_sms;
  ndev_sms = std::max(8, ndev_sms / g.sm_parts);
  (void)all_sms;
  const bool use_stream = g.use_stream, use_panel = g.use_panel, w_in_stage = g.w_in_stage;
  const bool rounds = (use_stream || use_panel) && g.round_iters > 0;
  // panel kernel: problems wider than one CTA's consumer warps run as cluster pairs (2 SMs per tile)
  int cs = 1;
  const bool use_rows = use_panel && g.use_rows;
  if (use_rows) {
    // rows kernel: CTA r of a cluster owns the row panels k = r (mod cs); the column chunks of the per-iteration exchange need
    // cs | nw.  The cluster size is a function of the problems only (never of the batch size): a node's result must not depend
    // on what else is in the launch
    bool wide = false, even = true, quad = true;
    for (int b : alive) { const int nw = g.node_inst[b]->h.pn.nw; wide = wide || nw > kPanelCtaWarps; even = even && nw % 2 == 0; quad = quad && nw % 4 == 0; }
    cs = (wide && even) ? 2 : 1;
    bool oct8 = quad;
    for (int b : alive) oct8 = oct8 && g.node_inst[b]->h.pn.nw % 8 == 0;
    if (const char *e = std::getenv("BQP_ROWS_CLUSTER")) { const int v = std::atoi(e); if (v == 1 || (v == 2 && even) || (v == 4 && quad) || (v == 8 && oct8)) cs = v; }
  } else if (use_panel) {
    for (int b : alive) if (g.node_inst[b]->h.pn.nw > kPanelCtaWarps) cs = 2;
    if (const char *e = std::getenv("BQP_PANEL_CLUSTER")) { const int v = std::atoi(e); if (v == 2 || (v == 1 && cs == 1)) cs = v; }
  }
  g.cs = cs;
  int capacity = rounds ? ndev_sms / cs : (1 << 30);   // streamed / panel kernels: one CTA per SM (registers, smem)
  auto panel_slots = [&](const HostInstance &h) {   // ring slots (this CTA's share of one panel each) beside the vectors
    if (use_rows) {   // full-width panels
      const long long fixed = (long long)rows_smem_bytes(h.npad, 0, cs) + 256, sb = (long long)h.pn.nw * kPanelRows * 32 * 8;
      long long cap = 8;
      if (const char *e = std::getenv("BQP_ROWS_SLOTS")) cap = std::max(2, std::atoi(e));
      // a multiple of the groups per CTA: every ring slot then belongs to ONE group for the whole launch.  (With 3 slots and
      // 2 groups a group waits on the parity of a slot whose previous phase belongs to the other group: the wait can alias.)
      const long long ns = std::min<long long>(cap, ((long long)kMaxSmem - fixed) / sb);
      return (int)(ns - ns % kRowsGroups);
    }
    const long long fixed = (long long)panel_smem_bytes(h.npad, 0, cs) + 256;
    const long long sb = (long long)(cs == 2 ? (h.pn.nw + 1) / 2 : h.pn.nw) * kPanelRows * 32 * 8;
    long long cap = 16;
    if (const char *e = std::getenv("BQP_PANEL_SLOTS")) cap = std::max(4, std::atoi(e));   // experiment knob (ring depth)
    return (int)std::min<long long>(cap, ((long long)kMaxSmem - fixed) / sb);
  };
  // groups in (progress, first appearance) order
  std::vector<bqp_instance *> uniq;
  std::map<std::pair<int, bqp_instance *>, int> gid;
  std::vector<std::vector<int>> members;
  std::vector<int> gprog;
  {
    std::vector<int> order(alive);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return progress[a] < progress[b]; });
    if (rounds) {
      // critical path first: a group's priority is the largest predicted number of remaining iterations of its nodes
      // (extrapolated from the decay of their residuals, `remaining`; unknown = infinite), least progress breaking ties
      std::map<std::pair<int, bqp_instance *>, double> prio;
      for (int b : order) {
        auto key = std::make_pair(progress[b], g.node_inst[b]);
        auto it = prio.find(key);
        if (it == prio.end()) prio[key] = remaining[b]; else it->second = std::max(it->second, remaining[b]);
      }
      // one look-up per node, not two per comparison: with 800 running leaves the sort used to cost more than the launch overhead
      std::vector<double> pr(g.node_inst.size(), 0.0);
      for (int b : order) pr[(size_t)b] = prio[std::make_pair(progress[b], g.node_inst[b])];
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pr[(size_t)a] > pr[(size_t)b]; });
    }
    for (int b : order) {
      bqp_instance *h = g.node_inst[b];
      auto key = std::make_pair(progress[b], h);
      auto it = gid.find(key);
      if (it == gid.end()) { gid[key] = (int)members.size(); members.emplace_back(); uniq.push_back(h); gprog.push_back(progress[b]); it = gid.find(key); }
      members[it->second].push_back(b);
    }
  }
  if (g.use_grid) {
    // one tile = up to 8 leaves of one problem on EVERY SM; the tiles of a launch run one after the other inside the kernel
    g.tiles.clear();
    g.tile_bytes_iter.clear(); g.tile_bytes_check.clear(); g.tile_check_every.clear(); g.tile_bytes_launch.clear();
    scheduled->clear();
    std::vector<DevInstance> &dinst = g.dinst_host; dinst.clear();
    std::map<bqp_instance *, int> inst_slot;
    size_t work_d = 0, smem = 0;
    for (size_t k = 0; k < members.size(); k++) {
      const HostInstance &h = uniq[k]->h;
      const int cnt = (int)members[k].size(), nt = (cnt + kMaxTT - 1) / kMaxTT;
      auto is = inst_slot.find(uniq[k]);
      if (is == inst_slot.end()) { inst_slot[uniq[k]] = (int)dinst.size(); dinst.push_back(uniq[k]->d); is = inst_slot.find(uniq[k]); }
      smem = std::max(smem, grid_smem_bytes(h.npad, h.m, h.n, g.grid_ctas));
      for (int ti = 0; ti < nt; ti++) {
        const int lo = (int)((long long)cnt * ti / nt), hi = (int)((long long)cnt * (ti + 1) / nt);
        DevTile t{};
        t.inst = is->second; t.nn = hi - lo; t.iter_begin = 0; t.iter_end = h.s.max_iter;
        for (int q = lo; q < hi; q++) {
          const int b = members[k][q];
          t.node[q - lo] = b; t.in_off[q - lo] = g.in_off[b]; t.out_off[q - lo] = g.out_off[b]; t.state_off[q - lo] = g.state_off[b];
          t.corr_off[q - lo] = -1;
          scheduled->push_back(b);
        }
        t.work_off = (long long)work_d;
        work_d += grid_work_doubles(h.npad, h.m, g.grid_ctas);
        g.tiles.push_back(t);
        g.tile_bytes_iter.push_back(h.gd.iter_bytes(h.npad));
        g.tile_bytes_check.push_back(2LL * h.npad * h.npad * 8 + 2 * 12LL * (long long)(h.gd.avl.size() + h.gd.tvl.size()));
        g.tile_bytes_launch.push_back((long long)h.npad * h.npad * 8);
        g.tile_check_every.push_back(h.s.check_termination);
      }
    }
    g.ntiles = (int)g.tiles.size(); g.tt = kMaxTT; g.smem = smem; g.nslots = 0; g.slot_bytes = 0; g.cs = 1;
    int rc;
    if ((rc = g.h_ti.reserve(sizeof(int) * (size_t)g.ntiles, g.stream))) return rc;
    if ((rc = g.d_ti.reserve(sizeof(int) * (size_t)g.ntiles, g.stream))) return rc;
    if ((rc = g.d_work.reserve(work_d * 8, g.stream))) return rc;
    if ((rc = g.d_tiles.reserve(sizeof(DevTile) * g.tiles.size(), g.stream))) return rc;
    if ((rc = g.d_insts.reserve(sizeof(DevInstance) * dinst.size(), g.stream))) return rc;
    if ((rc = g.d_gbar.reserve(sizeof(unsigned), g.stream))) return rc;
    CK(cudaMemcpyAsync(g.d_tiles.p, g.tiles.data(), sizeof(DevTile) * g.tiles.size(), cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemcpyAsync(g.d_insts.p, dinst.data(), sizeof(DevInstance) * dinst.size(), cudaMemcpyHostToDevice, g.stream));
    // no wait here: g.tiles / g.dinst_host are members, overwritten by the next plan only -- after the wait that ends this round
    g.round_h2d_bytes = (long long)(sizeof(DevTile) * g.tiles.size() + sizeof(DevInstance) * dinst.size());
    return BQP_OK;
  }
  if (use_rows && g.auto_cluster && cs == 2 && rounds) {
    // few tiles left (rolling B&B sessions: most trees are finished or between steps): a cluster of 4 per tile halves the
    // per-iteration latency (45 us instead of 80 at n = 500) and the idle SMs cost nothing.  Opt-in: the summation order of
    // the column sums depends on the cluster size, so a node's last bits then depend on the schedule
    bool quad = true;
    for (auto *inst : uniq) quad = quad && inst->h.pn.nw % 4 == 0;
    long long nt = 0;
    for (auto &mb : members) nt += ((long long)mb.size() + kPanelT - 1) / kPanelT;
    const int cap4 = (all_sms - 16) / 4 / g.sm_parts;     // clusters of 4 cannot use every SM of every GPC (measured: 33 at a time on 148 SMs)
    bool oct = quad;
    for (auto *inst : uniq) oct = oct && inst->h.pn.nw % 8 == 0;
    const int cap8 = 8 / g.sm_parts;          // clusters of 8: 16 of them took two waves (measured), 8 run at once
    if (oct && nt <= cap8) { cs = 8; g.cs = 8; capacity = cap8; }
    else if (quad && nt <= cap4) { cs = 4; g.cs = 4; capacity = cap4; }
  }
  const int slot_bytes = g.stage_bytes;
  if (use_panel && !use_rows && cs == 1)
    for (int b : alive) if (g.node_inst[b]->h.pn.nw > kPanelCtaWarps) return BQP_E_UNSUPPORTED;
  auto slot_size = [&](int t) { return slot_bytes + (w_in_stage ? kKC * t * 8 : 0); };
  auto stream_slots = [&](const HostInstance &h, int t) {   // slots PER QUAD (one ring per quad of consumer warps)
    const long long fixed = (long long)stream_smem_bytes(h.n, h.m, t, slot_size(t), 0, w_in_stage) + 1024;
    return (int)std::min<long long>(8, ((long long)kMaxSmem - fixed) / ((kStreamWarps / 4) * (long long)slot_size(t)));
  };
  // nodes per tile: as wide as shared memory allows ...
  int tt_cap = kMaxTT, nslots = 8;
  for (auto *inst : uniq) {
    int t = 0;
    if (use_panel) {
      t = panel_slots(inst->h) >= (use_rows ? 2 : 4) ? kPanelT : 0;   // panel kernel: pass 2 holds 2 panels; the rest are in flight
    } else if (use_stream) {
      t = kMaxTT;
      while (t > 1 && stream_slots(inst->h, t) < 3) t >>= 1;   // >= 3 stages in flight per quad: measured knee
      if (t == kMaxTT) t = kMaxTT / 2;                         // T=8 is FP64/smem-issue bound per SM
      if (stream_slots(inst->h, t) < 2) t = 0;
    } else if (g.use_small) {
      t = kMaxTT;                                                // the 8 leaves of a tile are the N dimension of the mma
    } else {
      t = max_tile_nodes(inst->h, g.threads);
    }
    if (t == 0) return BQP_E_UNSUPPORTED;
    tt_cap = std::min(tt_cap, t);
  }
  int widest = 1;
  for (auto &mb : members) widest = std::max<int>(widest, (int)mb.size());
  int tt = g_tune_tt ? std::min(g_tune_tt, use_panel ? tt_cap : (use_stream ? kMaxTT : tt_cap)) : std::min(tt_cap, pow2ceil(widest));
  if (g_tune_tt && use_stream)
    for (auto *inst : uniq) while (tt > 1 && stream_slots(inst->h, tt) < 2) tt >>= 1;
  // ... but narrow enough to keep every SM busy when the frontier (or what is left of it) is small.  Not for the panel
  // kernel: a tile-iteration costs the same for 1..8 nodes there (the nodes are the N dimension of the mma), so splitting
  // an instance's nodes over two tiles only doubles its HBM stream
  // Nor for the shared-memory-resident kernel: it is bound by instruction issue, and the threads of a missing leaf still issue
  if (!g_tune_tt && !use_panel && !g.use_small) {
    auto count_tiles = [&](int t) { long long c = 0; for (auto &mb : members) c += ((long long)mb.size() + t - 1) / t; return c; };
    while (tt > 1 && count_tiles(tt / 2) <= std::min(capacity, ndev_sms)) tt >>= 1;
  }
  if (use_panel) {
    nslots = 16;
    for (auto *inst : uniq) nslots = std::min(nslots, panel_slots(inst->h));
  } else if (use_stream)
    for (auto *inst : uniq) nslots = std::min(nslots, stream_slots(inst->h, tt));
  // tiles: split each group's nodes evenly; stop at capacity (whole groups only, so siblings stay in the same launch)
  g.tiles.clear();
  g.tile_bytes_iter.clear(); g.tile_bytes_check.clear(); g.tile_check_every.clear(); g.tile_bytes_launch.clear();
  g.nw_max = 0;
  scheduled->clear();
  std::vector<DevInstance> &dinst = g.dinst_host; dinst.clear();
  std::map<bqp_instance *, int> inst_slot;
  size_t work_d = 0, smem = 0;
  for (size_t k = 0; k < members.size(); k++) {
    const HostInstance &h = uniq[k]->h;
    const int cnt = (int)members[k].size(), nt = (cnt + tt - 1) / tt;
    if (!g.tiles.empty() && (long long)g.tiles.size() + nt > capacity) break;
    auto is = inst_slot.find(uniq[k]);
    if (is == inst_slot.end()) { inst_slot[uniq[k]] = (int)dinst.size(); dinst.push_back(uniq[k]->d); is = inst_slot.find(uniq[k]); }
    smem = std::max(smem, use_rows ? rows_smem_bytes(h.npad, nslots, cs) : use_panel ? panel_smem_bytes(h.npad, nslots, cs)
                          : (use_stream ? stream_smem_bytes(h.n, h.m, tt, slot_size(tt), nslots, w_in_stage)
                             : g.use_small ? small_smem_bytes(h.npad, h.m, h.sm.bytes) : tile_smem_bytes(h.n, h.m, tt, g.threads)));
    if (use_panel) g.nw_max = std::max(g.nw_max, h.pn.nw);
    for (int ti = 0; ti < nt; ti++) {
      const int lo = (int)((long long)cnt * ti / nt), hi = (int)((long long)cnt * (ti + 1) / nt);
      DevTile t{};
      t.inst = is->second; t.nn = hi - lo;
      t.iter_begin = gprog[k];
      t.iter_end = rounds ? std::min(gprog[k] + g.round_iters, h.s.max_iter) : h.s.max_iter;
      for (int q = lo; q < hi; q++) {
        const int b = members[k][q];
        t.node[q - lo] = b; t.in_off[q - lo] = g.in_off[b]; t.out_off[q - lo] = g.out_off[b]; t.state_off[q - lo] = g.state_off[b];
        t.corr_off[q - lo] = (size_t)b < g.corr_off.size() ? g.corr_off[b] : -1;
        scheduled->push_back(b);
      }
      t.work_off = (long long)work_d;
      work_d += use_rows ? rows_work_doubles(h.npad, h.m, cs) : use_panel ? cs * panel_work_doubles(h.npad, h.m) : tile_work_doubles(h.n, h.m, tt);
      g.tiles.push_back(t);
      // shared-memory-resident kernel: the blob is read from global memory once per launch, nothing per iteration
      g.tile_bytes_iter.push_back(g.use_small ? 0 : use_panel ? h.pn.iter_bytes() : (use_stream ? h.st.iter_bytes : h.factor_bytes()));
      g.tile_bytes_check.push_back(g.use_small ? 0 : use_panel ? h.pn.check_bytes() : (use_stream ? h.st.check_bytes : h.check_bytes()));
      g.tile_bytes_launch.push_back(g.use_small ? h.sm.bytes : use_panel ? h.pn.launch_bytes() : 0);
      g.tile_check_every.push_back(h.s.check_termination);
    }
  }
  g.small_mask = 0;
  if (g.use_small) for (const DevInstance &di : dinst) g.small_mask |= di.npad <= 32 ? 1 : 2;
  g.ntiles = (int)g.tiles.size(); g.tt = tt; g.smem = smem; g.nslots = nslots; g.slot_bytes = use_stream ? slot_size(tt) : slot_bytes;
  int rc;
  if ((rc = g.h_ti.reserve(sizeof(int) * (size_t)g.ntiles, g.stream))) return rc;
  if ((rc = g.d_ti.reserve(sizeof(int) * (size_t)g.ntiles, g.stream))) return rc;
  if ((rc = g.d_work.reserve(work_d * 8, g.stream))) return rc;
  if ((rc = g.d_tiles.reserve(sizeof(DevTile) * g.tiles.size(), g.stream))) return rc;
  if ((rc = g.d_insts.reserve(sizeof(DevInstance) * dinst.size(), g.stream))) return rc;
  CK(cudaMemcpyAsync(g.d_tiles.p, g.tiles.data(), sizeof(DevTile) * g.tiles.size(), cudaMemcpyHostToDevice, g.stream));
  CK(cudaMemcpyAsync(g.d_insts.p, dinst.data(), sizeof(DevInstance) * dinst.size(), cudaMemcpyHostToDevice, g.stream));
  // no wait here: g.tiles / g.dinst_host are members, overwritten by the next plan only -- after the wait that ends this round
  g.round_h2d_bytes = (long long)(sizeof(DevTile) * g.tiles.size() + sizeof(DevInstance) * dinst.size());
  return BQP_OK;
}

// eq_rho == 2 (per-node rho typing, osqp >= 0.4 update_bounds): a node whose integer-bound rows fall into another rho class
// than at setup (a branched binary variable: l = u, rho x 1e3) changes the reduced KKT matrix by a DIAGONAL term -- the
// integer rows of A_ext are rows of the identity (data.py:5-33) -- K_node = K + sum_{j in S} delta_j e_j e_j'.  The explicit
// inverse is corrected by Woodbury:  K_node^-1 = M - M[:,S] G M[S,:],  G = (diag(1/delta) + M_SS)^-1  (|S| <= depth).
// Per node the kernel gets S and G; M[:,S] = rows of M (symmetric) kept per problem (DevInstance::p_mint).
// Appends the blocks of nodes [first, first + count) to the context's correction buffer.
static int build_corrections(BatchCtx &g, int first, int count, const bqp_handle *handles, const double *const *l, const double *const *u) {
  std::vector<double> stage;
  const size_t base = g.corr_d;
  g.corr_off.resize((size_t)first + count, -1);
  for (int b = 0; b < count; b++) {
    const HostInstance &h = handles[b]->h;
    g.corr_off[(size_t)first + b] = -1;
    if (h.s.eq_rho != 2) continue;
    const int m = h.m, ni = h.n_int;
    auto cls = [&](int i, double lo, double up) {
      lo = std::max(lo, -kInfty) * h.E[i]; up = std::min(up, kInfty) * h.E[i];
      if (lo < -kInfty * kMinScaling && up > kInfty * kMinScaling) return kRhoMin;
      if (up - lo < kRhoTol) return kRhoEqFactor * h.s.rho;
      return h.s.rho;
    };
    for (int i = 0; i < m - ni; i++)
      if (cls(i, l[b][i], u[b][i]) != h.rho[i]) return BQP_E_UNSUPPORTED;     // only the integer-bound rows may change class
    std::vector<int> ks; std::vector<double> delta;
    for (int k = 0; k < ni; k++) {
      const int i = m - ni + k, j = h.i_idx[k];
      const double r = cls(i, l[b][i], u[b][i]);
      if (r == h.rho[i]) continue;
      const double a = h.E[i] * h.D[j];                // the scaled entry of the identity row
      ks.push_back(k); delta.push_back((r - h.rho[i]) * a * a);
    }
    const int nS = (int)ks.size();
    if (nS == 0) continue;
    if (nS > 64) return BQP_E_UNSUPPORTED;             // the kernel's per-node scratch (bqp_rows.cu kMaxS)
    // G = (diag(1/delta) + M_SS)^-1 by Gauss-Jordan with partial pivoting (nS <= n_int)
    std::vector<double> Wm((size_t)nS * 2 * nS, 0.0);
    for (int a = 0; a < nS; a++) {
      for (int c = 0; c < nS; c++) Wm[(size_t)a * 2 * nS + c] = host_panel_M(&h, h.i_idx[ks[a]], h.i_idx[ks[c]]);
      Wm[(size_t)a * 2 * nS + a] += 1.0 / delta[a];
      Wm[(size_t)a * 2 * nS + nS + a] = 1.0;
    }
    for (int c = 0; c < nS; c++) {
      int piv = c;
      for (int a = c + 1; a < nS; a++) if (std::fabs(Wm[(size_t)a * 2 * nS + c]) > std::fabs(Wm[(size_t)piv * 2 * nS + c])) piv = a;
      if (Wm[(size_t)piv * 2 * nS + c] == 0.0) return BQP_E_NONCONVEX;
      if (piv != c) for (int e = 0; e < 2 * nS; e++) std::swap(Wm[(size_t)piv * 2 * nS + e], Wm[(size_t)c * 2 * nS + e]);
      const double inv = 1.0 / Wm[(size_t)c * 2 * nS + c];
      for (int e = 0; e < 2 * nS; e++) Wm[(size_t)c * 2 * nS + e] *= inv;
      for (int a = 0; a < nS; a++) {
        if (a == c) continue;
        const double f = Wm[(size_t)a * 2 * nS + c];
        if (f == 0.0) continue;
        for (int e = 0; e < 2 * nS; e++) Wm[(size_t)a * 2 * nS + e] -= f * Wm[(size_t)c * 2 * nS + e];
      }
    }
    g.corr_off[(size_t)first + b] = (long long)(base + stage.size());
    stage.push_back((double)nS);
    for (int a = 0; a < nS; a++) stage.push_back((double)ks[a]);
    for (int a = 0; a < nS; a++) stage.push_back((double)h.i_idx[ks[a]]);
    for (int a = 0; a < nS; a++)
      for (int c = 0; c < nS; c++) stage.push_back(0.5 * (Wm[(size_t)a * 2 * nS + nS + c] + Wm[(size_t)c * 2 * nS + nS + a]));   // symmetrised
  }
  if (stage.empty()) return BQP_OK;
  int rc;
  if ((rc = g.d_corr.reserve_keep((base + stage.size()) * 8, base * 8, g.stream))) return rc;
  CK(cudaMemcpyAsync((double *)g.d_corr.p + base, stage.data(), stage.size() * 8, cudaMemcpyHostToDevice, g.stream));
  CK(ctx_sync(g));        // `stage` is pageable stack-lifetime memory
  g.corr_d = base + stage.size();
  return BQP_OK;
}

// engine selection over the nodes of the batch / session (g.node_inst): the row-split kernel when every problem has a dense
// panel layout, else the TMA-streamed kernel when every problem has a streamed layout (bqp_set_tuning(threads>0) forces the
// direct-load kernel); CTA shape of the direct-load kernel: one warp per 32-row slice of the widest panel
static void select_kernel(BatchCtx &g) {
  g.use_stream = g_tune_threads == 0;
  g.use_panel = g_tune_threads == 0;
  g.use_rows = true;
  if (const char *e = std::getenv("BQP_KERNEL")) {        // tests / A-B runs: "rows" (default when possible), "panel", "stream", "direct"
    if (!std::strcmp(e, "stream")) g.use_panel = false;
    else if (!std::strcmp(e, "direct")) g.use_panel = g.use_stream = false;
    else if (!std::strcmp(e, "panel")) g.use_rows = false;
  }
  g.stage_bytes = kStageValBytes;
  g.w_in_stage = true;
  g.max_iter_all = 0; g.round_ok = true;
  const int check0 = g.node_inst[0]->h.s.check_termination;
  int want = 2;
  bqp_instance *last = nullptr;
  for (bqp_instance *inst : g.node_inst) {
    if (inst == last) continue;      // nodes of one problem are usually adjacent
    last = inst;
    const HostInstance &h = inst->h;
    const HostStream &st = h.st;
    if (!h.pn.built) g.use_panel = false;
    if (!st.built || (int)st.groups.size() > 96) g.use_stream = false;   // 96 = groups cached in shared memory
    else g.stage_bytes = std::max(g.stage_bytes, st.slot_bytes);
    if (!inst->d.w_in_stage) g.w_in_stage = false;
    g.max_iter_all = std::max(g.max_iter_all, h.s.max_iter);
    if (h.s.check_termination != check0) g.round_ok = false;
    want = std::max(want, std::max(h.Ab.nslices, h.At.nslices));
  }
  g.eq2_unsupported = false;
  for (bqp_instance *inst : g.node_inst) if (inst->h.s.eq_rho == 2 && !(g.use_panel && g.use_rows)) g.eq2_unsupported = true;
  g.rows_ext = false;
  for (bqp_instance *inst : g.node_inst) if (inst->h.s.eq_rho == 2 || inst->h.s.adaptive_rho) g.rows_ext = true;
  // whole-GPU kernel: every problem of the batch has the layout (config 4) and nothing asked for another kernel
  g.use_grid = false;
  if (!g.use_panel && g_tune_threads == 0 && !std::getenv("BQP_KERNEL")) {
    bool all = true;
    size_t smem = 0;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g.device);
    last = nullptr;
    for (bqp_instance *inst : g.node_inst) {
      if (inst == last) continue;
      last = inst;
      if (!inst->h.gd.built) { all = false; break; }
      smem = std::max(smem, grid_smem_bytes(inst->h.npad, inst->h.m, inst->h.n, sms));
    }
    if (all && smem <= (size_t)kMaxSmem) {
      const int ctas = grid_max_ctas(g.device, smem);
      if (ctas > 0) { g.use_grid = true; g.grid_ctas = ctas; g.use_stream = false; }
    }
  }
  if (const char *e = std::getenv("BQP_KERNEL")) if (!std::strcmp(e, "grid")) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g.device);
    g.use_grid = true; g.use_panel = g.use_stream = false; g.grid_ctas = sms;
    for (bqp_instance *inst : g.node_inst) if (!inst->h.gd.built) g.use_grid = false;
  }
  // shared-memory-resident kernel: every problem of the batch carries the layout (small sparse problems that no dense kernel
  // serves: config 3) and no other kernel was asked for
  g.use_small = false;
  {
    const char *e = std::getenv("BQP_KERNEL");
    const bool forced = e && !std::strcmp(e, "small");
    if (g_tune_threads == 0 && !g.use_grid && (forced || (!e && !g.use_panel && !g.use_stream))) {
      bool all = true;
      for (bqp_instance *inst : g.node_inst) all = all && inst->h.sm.built && !inst->h.s.adaptive_rho && inst->h.s.eq_rho != 2;
      if (all) { g.use_small = true; g.use_panel = g.use_stream = false; }
    }
  }
  for (bqp_instance *inst : g.node_inst)
    if (inst->h.s.adaptive_rho && !((g.use_grid && inst->h.gd.spectral) || (g.use_panel && g.use_rows && inst->h.pn.spectral))) g.eq2_unsupported = true;
  if (g.use_panel) g.use_stream = false;
  if (!g.use_stream) g.w_in_stage = false;
  g.threads = g.use_small ? 256 : g.use_panel ? 0 : g.use_stream ? (kStreamWarps + 1) * 32 : (g_tune_threads ? g_tune_threads : 32 * std::min(pow2ceil(want), kMaxThreads / 32));
  // rounds: the streamed kernels run `round_iters` ADMM iterations per launch; finished nodes drop out and the rest are
  // re-tiled (narrower tiles as the frontier drains, so idle SMs pick up the stragglers).  0 = one launch.
  g.round_iters = 0;
  if ((g.use_stream || g.use_panel) && g.round_ok && !g.use_grid) {
    int r = 100;
    if (const char *e = std::getenv("BQP_ROUND_ITERS")) r = std::atoi(e);
    if (g.round_override >= 0) r = g.round_override;
    g.round_iters = r <= 0 ? 0 : ((r + check0 - 1) / check0) * check0;
  }
}

static int batch_upload(BatchCtx &g, int B, const bqp_handle *handles, const double *const *l, const double *const *u,
                        const double *const *x0, const double *const *y0) {
  if (B <= 0 || !handles || !l || !u || !x0 || !y0) return BQP_E_ARG;
  g.resident = false; g.ran = false;
  for (int b = 0; b < B; b++) {
    if (!handles[b] || !l[b] || !u[b] || !x0[b] || !y0[b]) return BQP_E_ARG;
    if (!handles[b]->on_device) return BQP_E_CUDA;
    if (handles[b]->device != handles[0]->device) return BQP_E_ARG;
    const int m = handles[b]->h.m;
    for (int i = 0; i < m; i++)
      if (l[b][i] > u[b][i]) return BQP_E_BOUNDS;   // osqp update_bounds raises
  }
  int rc = ctx_init(g, handles[0]->device);
  if (rc) return rc;
  CK(cudaSetDevice(g.device));

  g.node_inst.assign(B, nullptr); g.in_off.assign(B, 0); g.out_off.assign(B, 0); g.state_off.assign(B, 0);
  size_t in_d = 0, out_d = 0, st_d = 0;
  for (int b = 0; b < B; b++) {
    const HostInstance &h = handles[b]->h;
    g.in_off[b] = (long long)in_d; g.out_off[b] = (long long)out_d; g.state_off[b] = (long long)st_d; g.node_inst[b] = handles[b];
    in_d += 3 * (size_t)h.m + h.n; out_d += (size_t)h.m + h.n; st_d += 2 * (size_t)h.m + h.n + 1;      // scaled x, z, y (+ the leaf's rho: adaptive rho)
  }
  g.B = B; g.in_doubles = in_d; g.out_doubles = out_d;
  g.session = false;
  g.corr_d = 0; g.corr_off.clear();
  if ((rc = build_corrections(g, 0, B, handles, l, u))) return rc;
  g.auto_cluster = g.auto_cluster_default;
  if (const char *e = std::getenv("BQP_ROWS_AUTO_CLUSTER")) g.auto_cluster = std::atoi(e) != 0;
  select_kernel(g);
  if (g.eq2_unsupported) return BQP_E_UNSUPPORTED;
  if ((rc = g.h_in.reserve(in_d * 8, g.stream))) return rc;
  if ((rc = g.h_out.reserve(out_d * 8, g.stream))) return rc;
  if ((rc = g.h_ns.reserve(sizeof(NodeScalars) * (size_t)B, g.stream))) return rc;
  if ((rc = g.d_in.reserve(in_d * 8, g.stream))) return rc;
  if ((rc = g.d_out.reserve(out_d * 8, g.stream))) return rc;
  if ((rc = g.d_ns.reserve(sizeof(NodeScalars) * (size_t)B, g.stream))) return rc;
  if ((rc = g.d_state.reserve(std::max<size_t>(st_d, 1) * 8, g.stream))) return rc;
  // pack the per-node inputs: l[m] u[m] x0[n] y0[m]
  double *hin = (double *)g.h_in.p;
  for (int b = 0; b < B; b++) {
    const HostInstance &h = handles[b]->h;
    double *p = hin + g.in_off[b];
    std::memcpy(p, l[b], 8 * (size_t)h.m);
    std::memcpy(p + h.m, u[b], 8 * (size_t)h.m);
    std::memcpy(p + 2 * (size_t)h.m, x0[b], 8 * (size_t)h.n);
    std::memcpy(p + 2 * (size_t)h.m + h.n, y0[b], 8 * (size_t)h.m);
  }
  static const bool fast_ok = !(std::getenv("BQP_FAST_PATH") && std::atoi(std::getenv("BQP_FAST_PATH")) == 0);
  g.fast = fast_ok && g.want_fast && g.round_iters == 0;
  g.out_fetched = false;
  CK(cudaEventRecord(g.ev[0], g.stream));
  CK(cudaMemcpyAsync(g.d_in.p, hin, in_d * 8, cudaMemcpyHostToDevice, g.stream));
  CK(cudaEventRecord(g.ev[1], g.stream));
  float ms = 0;
  if (!g.fast) {           // h_in is this context's own pinned buffer: nothing touches it before the next upload
    CK(ctx_sync(g));
    cudaEventElapsedTime(&ms, g.ev[0], g.ev[1]);
  }
  g.timing = bqp_timing{};
  g.timing.h2d_ms = ms;
  g.timing.h2d_bytes = (long long)(in_d * 8);
  g.timing.threads = g.threads;
  g.resident = true;
  return BQP_OK;
}

// one launch over (a capacity-bounded subset of) the running nodes; finished nodes leave `alive`
static int run_round(BatchCtx &g, std::vector<int> &alive, std::vector<int> &progress, std::vector<double> &dist,
                     std::vector<double> &remaining, std::vector<int> *finished, long long *tile_iters, long long *bytes) {
  const double kUnknown = 1e30;
  static const bool predict = std::getenv("BQP_NO_PREDICT") == nullptr;
  std::vector<int> scheduled;
  int rc = plan_round(g, alive, progress, remaining, &scheduled);
  if (rc) return rc;
  rc = g.use_grid
           ? launch_admm_grid(g.grid_ctas, (const DevInstance *)g.d_insts.p, (const DevTile *)g.d_tiles.p, g.ntiles, (const double *)g.d_in.p,
                              (double *)g.d_out.p, (double *)g.d_work.p, (NodeScalars *)g.d_ns.p, (int *)g.d_ti.p, (unsigned *)g.d_gbar.p,
                              g.smem, g.stream)
       : (g.use_panel && g.use_rows)
           ? launch_admm_rows(g.cs, g.rows_ext ? 1 : 0, g.nslots, (double *)g.d_state.p, (const DevInstance *)g.d_insts.p, (const DevTile *)g.d_tiles.p, g.ntiles,
                              (const double *)g.d_in.p, (double *)g.d_out.p, (double *)g.d_work.p, (NodeScalars *)g.d_ns.p, (int *)g.d_ti.p,
                              g.smem, (const double *)g.d_corr.p, g.stream)
       : g.use_panel
           ? launch_admm_panel(g.cs, g.nw_max, g.nslots, (double *)g.d_state.p, (const DevInstance *)g.d_insts.p, (const DevTile *)g.d_tiles.p,
                               g.ntiles, (const double *)g.d_in.p, (double *)g.d_out.p, (double *)g.d_work.p, (NodeScalars *)g.d_ns.p,
                               (int *)g.d_ti.p, g.smem, g.stream)
       : g.use_stream
           ? launch_admm_stream(g.tt, g.slot_bytes, g.nslots, g.w_in_stage ? 1 : 0, (double *)g.d_state.p,
                                (const DevInstance *)g.d_insts.p, (const DevTile *)g.d_tiles.p, g.ntiles, (const double *)g.d_in.p,
                                (double *)g.d_out.p, (double *)g.d_work.p, (NodeScalars *)g.d_ns.p, (int *)g.d_ti.p, g.smem, g.stream)
       : g.use_small
           ? launch_admm_small(g.small_mask, (const DevInstance *)g.d_insts.p, (const DevTile *)g.d_tiles.p, g.ntiles, (const double *)g.d_in.p,
                               (double *)g.d_out.p, (NodeScalars *)g.d_ns.p, (int *)g.d_ti.p, g.smem, g.stream)
           : launch_admm(g.tt, g.threads, (const DevInstance *)g.d_insts.p, (const DevTile *)g.d_tiles.p, g.ntiles,
                         (const double *)g.d_in.p, (double *)g.d_out.p, (double *)g.d_work.p, (NodeScalars *)g.d_ns.p,
                         (int *)g.d_ti.p, g.smem, g.stream);
  if (rc) { g_last_cuda = cudaGetLastError(); return rc; }
  if (g.fast) CK(cudaEventRecord(g.ev[2], g.stream));
  CK(cudaMemcpyAsync(g.h_ns.p, g.d_ns.p, sizeof(NodeScalars) * (size_t)g.B, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaMemcpyAsync(g.h_ti.p, g.d_ti.p, sizeof(int) * (size_t)g.ntiles, cudaMemcpyDeviceToHost, g.stream));
  if (g.fast) {            // the one launch of this batch: the iterates come back behind it, one wait for everything
    CK(cudaMemcpyAsync(g.h_out.p, g.d_out.p, g.out_doubles * 8, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaEventRecord(g.ev[3], g.stream));
    g.out_fetched = true;
  }
  CK(ctx_sync(g));
  const NodeScalars *hs = (const NodeScalars *)g.h_ns.p;
  const int *ti = (const int *)g.h_ti.p;
  for (int t = 0; t < g.ntiles; t++) {
    *tile_iters += ti[t];
    const int ce = g.tile_check_every[t];
    *bytes += (long long)ti[t] * g.tile_bytes_iter[t] + (long long)(ti[t] / ce + ((ti[t] % ce) ? 1 : 0)) * g.tile_bytes_check[t] + g.tile_bytes_launch[t];
  }
  std::vector<char> done(g.B, 0);
  for (int b : scheduled) {
    if (hs[b].status == BQP_UNSOLVED) {
      const int before = progress[b];
      progress[b] = hs[b].iters;   // iterations completed so far (its tile's iter_end)
      const double d = hs[b].pri_res;
      if (predict && d == d && d > 0) {
        if (dist[b] == dist[b] && dist[b] > d && d > 1.0 && progress[b] > before)
          remaining[b] = std::log(d) / std::log(dist[b] / d) * (progress[b] - before);
        else if (dist[b] == dist[b]) remaining[b] = kUnknown * 0.5;   // not converging yet: behind the unknown ones only
        dist[b] = d;
      }
    } else { done[b] = 1; if (finished) finished->push_back(b); }
  }
  std::vector<int> next;
  for (int b : alive) if (!done[b]) next.push_back(b);
  if (next.size() == alive.size() && scheduled.empty()) return BQP_E_CUDA;
  alive.swap(next);
  return BQP_OK;
}

static void fill_launch_timing(BatchCtx &g, int first_tiles, int first_tt, long long first_smem, int first_slots) {
  g.timing.tiles = first_tiles; g.timing.tile_nodes = first_tt; g.timing.smem_bytes = first_smem;
  if (g.use_panel && g.use_rows) g.timing.threads = kRowsThreads;
  else if (g.use_panel) {
    const int nwc = g.cs == 2 ? (g.nw_max + 1) / 2 : g.nw_max;
    g.timing.threads = panel_cta_warps(nwc) * 32;
  }
  g.timing.kernel = g.use_grid ? 4 : g.use_small ? 5 : g.use_panel ? (g.use_rows ? 3 : 2) : (g.use_stream ? 1 : 0);
  if (g.use_grid) { g.timing.threads = 512; g.timing.tiles = g.grid_ctas; }
  g.timing.ring_slots = first_slots;
}

static int batch_run(BatchCtx &g) {
  if (!g.resident || g.session) return BQP_E_ARG;
  CK(cudaSetDevice(g.device));
  std::vector<int> alive(g.B), progress(g.B, 0);
  // scheduling hint per running node: distance to the tolerance at its last check (reported by the kernel) and the
  // number of iterations it is predicted to need still, from the geometric decay of that distance between two rounds
  std::vector<double> dist(g.B, NAN), remaining(g.B, 1e30);
  for (int b = 0; b < g.B; b++) alive[b] = b;
  long long tile_iters = 0, bytes = 0, h2d_extra = 0;
  int launches = 0, first_tiles = 0, first_tt = 0, first_slots = 0;
  long long first_smem = 0;
  CK(cudaEventRecord(g.fast ? g.ev[4] : g.ev[1], g.stream));
  while (!alive.empty()) {
    int rc = run_round(g, alive, progress, dist, remaining, nullptr, &tile_iters, &bytes);
    if (rc) return rc;
    h2d_extra += g.round_h2d_bytes;
    if (launches == 0) { first_tiles = g.ntiles; first_tt = g.tt; first_smem = (long long)g.smem; first_slots = (g.use_panel || g.use_stream) ? g.nslots : 0; }
    launches++;
    if (g.fast && !alive.empty()) return BQP_E_CUDA;      // a kernel without rounds finishes every node it is given
  }
  float ms = 0;
  if (g.fast) {            // everything was waited for inside the round
    cudaEventElapsedTime(&ms, g.ev[0], g.ev[1]); g.timing.h2d_ms = ms;
    cudaEventElapsedTime(&ms, g.ev[2], g.ev[3]); g.timing.d2h_ms = ms;
    cudaEventElapsedTime(&ms, g.ev[4], g.ev[2]);
  } else {
    CK(cudaEventRecord(g.ev[2], g.stream));
    CK(ctx_sync(g));
    cudaEventElapsedTime(&ms, g.ev[1], g.ev[2]);
  }
  g.timing.kernel_ms = ms;
  g.timing.launches = launches;
  fill_launch_timing(g, first_tiles, first_tt, first_smem, first_slots);
  g.timing.tile_iters = tile_iters; g.timing.stream_bytes = bytes;
  g.round_h2d_total = h2d_extra;
  g.ran = true;
  return BQP_OK;
}

// ---- rolling session: nodes join a resident batch while others are still iterating; every call of session_round is one
// launch (one round of ADMM iterations over a capacity-bounded, critical-path-first subset of the running nodes)
static int session_begin(BatchCtx &g) {
  g.resident = false; g.ran = false; g.session = true;
  g.fast = false; g.out_fetched = false;
  g.auto_cluster = true;
  if (const char *e = std::getenv("BQP_ROWS_AUTO_CLUSTER")) g.auto_cluster = std::atoi(e) != 0;
  g.B = 0; g.node_inst.clear(); g.in_off.clear(); g.out_off.clear(); g.state_off.clear();
  g.s_alive.clear(); g.s_progress.clear(); g.s_dist.clear(); g.s_remaining.clear();
  g.s_in_d = g.s_out_d = g.s_st_d = 0;
  g.corr_d = 0; g.corr_off.clear();
  g.s_tile_iters = g.s_bytes = 0; g.s_launches = 0; g.s_kernel_ms = 0;
  g.timing = bqp_timing{};
  return BQP_OK;
}

static int session_append(BatchCtx &g, int B, const bqp_handle *handles, const double *const *l, const double *const *u,
                          const double *const *x0, const double *const *y0, int *first_id) {
  if (!g.session || B <= 0 || !handles || !l || !u || !x0 || !y0) return BQP_E_ARG;
  for (int b = 0; b < B; b++) {
    if (!handles[b] || !l[b] || !u[b] || !x0[b] || !y0[b]) return BQP_E_ARG;
    if (!handles[b]->on_device) return BQP_E_CUDA;
    if (handles[b]->device != handles[0]->device || (g.B > 0 && handles[b]->device != g.device)) return BQP_E_ARG;
    const int m = handles[b]->h.m;
    for (int i = 0; i < m; i++)
      if (l[b][i] > u[b][i]) return BQP_E_BOUNDS;
  }
  int rc = ctx_init(g, handles[0]->device);
  if (rc) return rc;
  CK(cudaSetDevice(g.device));
  const int B0 = g.B;
  const size_t in0 = g.s_in_d, out0 = g.s_out_d, st0 = g.s_st_d;
  size_t in_d = in0, out_d = out0, st_d = st0;
  for (int b = 0; b < B; b++) {
    const HostInstance &h = handles[b]->h;
    g.in_off.push_back((long long)in_d); g.out_off.push_back((long long)out_d); g.state_off.push_back((long long)st_d);
    g.node_inst.push_back(handles[b]);
    in_d += 3 * (size_t)h.m + h.n; out_d += (size_t)h.m + h.n; st_d += 2 * (size_t)h.m + h.n + 1;      // scaled x, z, y (+ the leaf's rho: adaptive rho)
    g.s_alive.push_back(B0 + b); g.s_progress.push_back(0); g.s_dist.push_back(NAN); g.s_remaining.push_back(1e30);
  }
  g.B = B0 + B; g.s_in_d = in_d; g.s_out_d = out_d; g.s_st_d = st_d;
  g.in_doubles = in_d; g.out_doubles = out_d;
  select_kernel(g);
  if (g.eq2_unsupported) return BQP_E_UNSUPPORTED;
  if ((rc = build_corrections(g, B0, B, handles, l, u))) return rc;
  // (a kernel without rounds -- the direct-load kernel of small problems -- finishes every running node in one "round")
  if ((rc = g.h_in.reserve((in_d - in0) * 8, g.stream))) return rc;        // staging of the new nodes only
  if ((rc = g.h_out.reserve_keep(out_d * 8, out0 * 8, g.stream))) return rc;
  if ((rc = g.h_ns.reserve(sizeof(NodeScalars) * (size_t)g.B, g.stream))) return rc;
  if ((rc = g.d_in.reserve_keep(in_d * 8, in0 * 8, g.stream))) return rc;
  if ((rc = g.d_out.reserve_keep(out_d * 8, out0 * 8, g.stream))) return rc;
  if ((rc = g.d_ns.reserve_keep(sizeof(NodeScalars) * (size_t)g.B, sizeof(NodeScalars) * (size_t)B0, g.stream))) return rc;
  if ((rc = g.d_state.reserve_keep(std::max<size_t>(st_d, 1) * 8, st0 * 8, g.stream))) return rc;
  double *hin = (double *)g.h_in.p;
  for (int b = 0; b < B; b++) {
    const HostInstance &h = handles[b]->h;
    double *p = hin + (g.in_off[B0 + b] - (long long)in0);
    std::memcpy(p, l[b], 8 * (size_t)h.m);
    std::memcpy(p + h.m, u[b], 8 * (size_t)h.m);
    std::memcpy(p + 2 * (size_t)h.m, x0[b], 8 * (size_t)h.n);
    std::memcpy(p + 2 * (size_t)h.m + h.n, y0[b], 8 * (size_t)h.m);
  }
  CK(cudaMemcpyAsync((double *)g.d_in.p + in0, hin, (in_d - in0) * 8, cudaMemcpyHostToDevice, g.stream));
  CK(ctx_sync(g));        // the staging buffer is reused by the next append
  g.timing.h2d_bytes += (long long)((in_d - in0) * 8);
  g.resident = true;
  if (first_id) *first_id = B0;
  return BQP_OK;
}

static int session_round(BatchCtx &g, std::vector<int> *finished) {
  if (!g.session || !g.resident) return BQP_E_ARG;
  if (g.s_alive.empty()) return BQP_OK;
  CK(cudaSetDevice(g.device));
  CK(cudaEventRecord(g.ev[1], g.stream));
  const size_t f0 = finished->size();
  int rc = run_round(g, g.s_alive, g.s_progress, g.s_dist, g.s_remaining, finished, &g.s_tile_iters, &g.s_bytes);
  if (rc) return rc;
  // results of the nodes that terminated in this launch: x | y into the pinned mirror of the output buffer
  for (size_t k = f0; k < finished->size(); k++) {
    const int b = (*finished)[k];
    const HostInstance &h = g.node_inst[b]->h;
    CK(cudaMemcpyAsync((double *)g.h_out.p + g.out_off[b], (const double *)g.d_out.p + g.out_off[b], 8 * ((size_t)h.n + h.m),
                       cudaMemcpyDeviceToHost, g.stream));
    g.timing.d2h_bytes += (long long)(8 * ((size_t)h.n + h.m));
  }
  CK(cudaEventRecord(g.ev[2], g.stream));
  CK(ctx_sync(g));
  float ms = 0;
  cudaEventElapsedTime(&ms, g.ev[1], g.ev[2]);
  g.s_kernel_ms += ms;
  if (g.s_launches == 0) fill_launch_timing(g, g.ntiles, g.tt, (long long)g.smem, (g.use_panel || g.use_stream) ? g.nslots : 0);
  g.s_launches++;
  g.timing.kernel_ms = g.s_kernel_ms; g.timing.launches = g.s_launches;
  g.timing.tile_iters = g.s_tile_iters; g.timing.stream_bytes = g.s_bytes;
  g.timing.h2d_bytes += g.round_h2d_bytes;
  return BQP_OK;
}

static int batch_download(BatchCtx &g, double *const *x, double *const *y, const bqp_node_out *out) {
  if (!g.resident || !g.ran) return BQP_E_ARG;
  CK(cudaSetDevice(g.device));
  if (!g.out_fetched) {
    CK(cudaEventRecord(g.ev[2], g.stream));
    CK(cudaMemcpyAsync(g.h_out.p, g.d_out.p, g.out_doubles * 8, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaMemcpyAsync(g.h_ns.p, g.d_ns.p, sizeof(NodeScalars) * (size_t)g.B, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaEventRecord(g.ev[3], g.stream));
    CK(ctx_sync(g));
    float ms = 0;
    cudaEventElapsedTime(&ms, g.ev[2], g.ev[3]);
    g.timing.d2h_ms = ms;
  }
  g.timing.d2h_bytes = (long long)(g.out_doubles * 8 + sizeof(NodeScalars) * (size_t)g.B);
  g.timing.h2d_bytes = (long long)(g.in_doubles * 8) + g.round_h2d_total;
  const double *ho = (const double *)g.h_out.p;
  const NodeScalars *hs = (const NodeScalars *)g.h_ns.p;
  long long node_iters = 0;
  for (int b = 0; b < g.B; b++) {
    const HostInstance &h = g.node_inst[b]->h;
    if (x && x[b]) std::memcpy(x[b], ho + g.out_off[b], 8 * (size_t)h.n);
    if (y && y[b]) std::memcpy(y[b], ho + g.out_off[b] + h.n, 8 * (size_t)h.m);
    if (out) {
      if (out->status) out->status[b] = hs[b].status;
      if (out->iters) out->iters[b] = hs[b].iters;
      if (out->obj) out->obj[b] = hs[b].obj;
      if (out->pri_res) out->pri_res[b] = hs[b].pri_res;
      if (out->dua_res) out->dua_res[b] = hs[b].dua_res;
      if (out->lower) out->lower[b] = hs[b].lower;
    }
    node_iters += hs[b].iters;
  }
  g.timing.node_iters = node_iters;
  return BQP_OK;
}

// BQP_API_TIMERS=1: wall time of the three phases of bqp_solve_multi and the device time of its kernels, summed over the process
struct ApiTimers {
  bool on = std::getenv("BQP_API_TIMERS") != nullptr;
  double up = 0, run = 0, down = 0, kernel_ms = 0; long long calls = 0;
  ~ApiTimers() {
    if (on && calls)
      std::fprintf(stderr, "API %lld solve_multi calls: upload %.3f s, run %.3f s (kernels %.3f s), download %.3f s\n", calls, up, run, kernel_ms * 1e-3, down);
  }
};
static ApiTimers g_at;
static inline double at_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int solve_multi(BatchCtx &g, int B, const bqp_handle *handles, const double *const *l, const double *const *u,
                       const double *const *x0, const double *const *y0, double *const *x, double *const *y,
                       const bqp_node_out *out) {
  const double t0 = g_at.on ? at_now() : 0.0;
  g.want_fast = true;
  int rc = batch_upload(g, B, handles, l, u, x0, y0);
  g.want_fast = false;
  if (rc) return rc;
  const double t1 = g_at.on ? at_now() : 0.0;
  if ((rc = batch_run(g))) return rc;
  const double t2 = g_at.on ? at_now() : 0.0;
  rc = batch_download(g, x, y, out);
  if (g_at.on) { const double t3 = at_now(); g_at.up += t1 - t0; g_at.run += t2 - t1; g_at.down += t3 - t2; g_at.kernel_ms += g.timing.kernel_ms; g_at.calls++; }
  return rc;
}

int bqp_batch_upload(int B, const bqp_handle *handles, const double *const *l, const double *const *u,
                     const double *const *x0, const double *const *y0) { return batch_upload(g0, B, handles, l, u, x0, y0); }
int bqp_batch_run(void) { return batch_run(g0); }
int bqp_batch_download(double *const *x, double *const *y, const bqp_node_out *out) { return batch_download(g0, x, y, out); }
int bqp_solve_multi(int B, const bqp_handle *handles, const double *const *l, const double *const *u,
                    const double *const *x0, const double *const *y0, double *const *x, double *const *y,
                    const bqp_node_out *out) { return solve_multi(g0, B, handles, l, u, x0, y0, x, y, out); }

/* ---- explicit contexts: one stream + staging buffers each, usable from different host threads at the same time */
struct bqp_context { BatchCtx c; };
int bqp_ctx_create(int device, int run_to_completion, bqp_ctx *out) {
  if (!out) return BQP_E_ARG;
  *out = nullptr;
  bqp_context *ctx = new (std::nothrow) bqp_context();
  if (!ctx) return BQP_E_ALLOC;
  ctx->c.blocking_sync = true;
  ctx->c.round_override = run_to_completion ? 0 : -1;
  int rc = ctx_init(ctx->c, device);
  if (rc) { ctx_release(ctx->c); delete ctx; return rc; }
  { std::lock_guard<std::mutex> lk(g_ctx_mu); g_ctxs.push_back(&ctx->c); }
  *out = ctx;
  return BQP_OK;
}
int bqp_ctx_free(bqp_ctx ctx) {
  if (!ctx) return BQP_OK;
  { std::lock_guard<std::mutex> lk(g_ctx_mu); g_ctxs.erase(std::remove(g_ctxs.begin(), g_ctxs.end(), &ctx->c), g_ctxs.end()); }
  ctx_release(ctx->c);
  delete ctx;
  return BQP_OK;
}
int bqp_ctx_solve_multi(bqp_ctx ctx, int B, const bqp_handle *handles, const double *const *l, const double *const *u,
                        const double *const *x0, const double *const *y0, double *const *x, double *const *y,
                        const bqp_node_out *out) {
  if (!ctx) return BQP_E_ARG;
  return solve_multi(ctx->c, B, handles, l, u, x0, y0, x, y, out);
}
int bqp_session_begin(bqp_ctx ctx) { return session_begin(ctx ? ctx->c : g0); }
int bqp_session_append(bqp_ctx ctx, int B, const bqp_handle *handles, const double *const *l, const double *const *u,
                       const double *const *x0, const double *const *y0, int *first_id) {
  return session_append(ctx ? ctx->c : g0, B, handles, l, u, x0, y0, first_id);
}
int bqp_session_round(bqp_ctx ctx, int *finished_ids, int cap, int *n_finished, int *running) {
  BatchCtx &g = ctx ? ctx->c : g0;
  std::vector<int> fin;
  const int rc = session_round(g, &fin);
  if (rc) return rc;
  if ((int)fin.size() > cap || (!finished_ids && !fin.empty())) return BQP_E_ARG;
  for (size_t k = 0; k < fin.size(); k++) finished_ids[k] = fin[k];
  if (n_finished) *n_finished = (int)fin.size();
  if (running) *running = (int)g.s_alive.size();
  return BQP_OK;
}
int bqp_session_fetch(bqp_ctx ctx, int id, double *x, double *y, const bqp_node_out *out) {
  BatchCtx &g = ctx ? ctx->c : g0;
  if (!g.session || id < 0 || id >= g.B) return BQP_E_ARG;
  const NodeScalars &r = ((const NodeScalars *)g.h_ns.p)[id];
  if (r.status == BQP_UNSOLVED) return BQP_E_ARG;        // still running
  const HostInstance &h = g.node_inst[id]->h;
  const double *ho = (const double *)g.h_out.p + g.out_off[id];
  if (x) std::memcpy(x, ho, 8 * (size_t)h.n);
  if (y) std::memcpy(y, ho + h.n, 8 * (size_t)h.m);
  if (out) {
    if (out->status) out->status[0] = r.status;
    if (out->iters) out->iters[0] = r.iters;
    if (out->obj) out->obj[0] = r.obj;
    if (out->pri_res) out->pri_res[0] = r.pri_res;
    if (out->dua_res) out->dua_res[0] = r.dua_res;
    if (out->lower) out->lower[0] = r.lower;
  }
  g.timing.node_iters += r.iters;
  return BQP_OK;
}
int bqp_ctx_set_sm_share(bqp_ctx ctx, int parts) {
  BatchCtx &g = ctx ? ctx->c : g0;
  g.sm_parts = parts < 1 ? 1 : parts;
  return BQP_OK;
}
int bqp_ctx_set_auto_cluster(bqp_ctx ctx, int on) {
  BatchCtx &g = ctx ? ctx->c : g0;
  g.auto_cluster_default = on != 0;
  return BQP_OK;
}
int bqp_ctx_last_timing(bqp_ctx ctx, bqp_timing *t) {
  if (!ctx || !t) return BQP_E_ARG;
  *t = ctx->c.timing;
  return BQP_OK;
}

int bqp_solve_batch(bqp_handle h, int B, const double *l, const double *u, const double *x0, const double *y0,
                    double *x, double *y, const bqp_node_out *out) {
  if (!h || B <= 0 || !l || !u || !x0 || !y0) return BQP_E_ARG;
  const int n = h->h.n, m = h->h.m;
  std::vector<bqp_handle> hs(B, h);
  std::vector<const double *> pl(B), pu(B), px0(B), py0(B);
  std::vector<double *> px(B), py(B);
  for (int b = 0; b < B; b++) {
    pl[b] = l + (size_t)b * m; pu[b] = u + (size_t)b * m; px0[b] = x0 + (size_t)b * n; py0[b] = y0 + (size_t)b * m;
    px[b] = x ? x + (size_t)b * n : nullptr; py[b] = y ? y + (size_t)b * m : nullptr;
  }
  return bqp_solve_multi(B, hs.data(), pl.data(), pu.data(), px0.data(), py0.data(), px.data(), py.data(), out);
}

int bqp_last_timing(bqp_timing *t) {
  if (!t) return BQP_E_ARG;
  *t = g0.timing;
  return BQP_OK;
}

int bqp_get_dims(bqp_handle h, int *n, int *m, int *npad, long long *factor_bytes, long long *check_bytes) {
  if (!h) return BQP_E_ARG;
  if (n) *n = h->h.n;
  if (m) *m = h->h.m;
  if (npad) *npad = h->h.npad;
  if (factor_bytes) *factor_bytes = h->h.factor_bytes();
  if (check_bytes) *check_bytes = h->h.check_bytes();
  return BQP_OK;
}

int bqp_get_scaling(bqp_handle h, double *D, double *E, double *c) {
  if (!h) return BQP_E_ARG;
  if (D) std::memcpy(D, h->h.D.data(), 8 * (size_t)h->h.n);
  if (E) std::memcpy(E, h->h.E.data(), 8 * (size_t)h->h.m);
  if (c) *c = h->h.c;
  return BQP_OK;
}

int bqp_get_inverse_guard(bqp_handle h, double *error, int *in_use) {
  if (!h) return BQP_E_ARG;
  if (error) *error = h->h.pn_inverse_error;
  if (in_use) *in_use = (h->h.pn.built || h->h.gd.built || h->h.sm.built) ? 1 : 0;
  return BQP_OK;
}

int bqp_handle_device(bqp_handle h) { return h ? h->device : -1; }

int bqp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char *bqp_strerror(int code) {
  static char buf[160];
  switch (code) {
    case BQP_OK: return "ok";
    case BQP_E_ARG: return "bad argument (null pointer, dimension or setting)";
    case BQP_E_BOUNDS: return "lower bound greater than upper bound";
    case BQP_E_NONCONVEX: return "reduced KKT matrix is not positive definite (non-convex problem)";
    case BQP_E_CUDA:
      std::snprintf(buf, sizeof buf, "CUDA error or no usable sm_100 device (%s)",
                    g_last_cuda == cudaSuccess ? "no device" : cudaGetErrorString(g_last_cuda));
      return buf;
    case BQP_E_ALLOC: return "out of memory";
    case BQP_E_UNSUPPORTED: return "unsupported setting or problem too large for one CTA's shared memory";
    case BQP_BNB_E_EXPLOR_RULE: return "Tree exploring strategy not recognized";
    case BQP_BNB_E_BRANCH_RULE: return "No variable selection rule recognized!";
  }
  return "unknown error";
}

const char *bqp_version(void) { return "bqp 0.1 (sm_100a)"; }

int bqp_debug_host_kkt_solve(bqp_handle h, double *rhs_xz) {
  if (!h || !rhs_xz) return BQP_E_ARG;
  host_kkt_solve(&h->h, rhs_xz);
  return BQP_OK;
}

int bqp_debug_host_stream_kkt_solve(bqp_handle h, double *rhs_xz) {
  if (!h || !rhs_xz) return BQP_E_ARG;
  return host_stream_kkt_solve(&h->h, rhs_xz);
}

int bqp_debug_host_panel_kkt_solve(bqp_handle h, double *rhs_xz) {
  if (!h || !rhs_xz) return BQP_E_ARG;
  return host_panel_kkt_solve(&h->h, rhs_xz);
}

int bqp_debug_host_small_kkt_solve(bqp_handle h, double *rhs_xz) {
  if (!h || !rhs_xz) return BQP_E_ARG;
  return host_small_kkt_solve(&h->h, rhs_xz);
}

int bqp_debug_host_matvec(bqp_handle h, int which, const double *in, double *out) {
  if (!h || !in || !out || which < 0 || which > 5) return BQP_E_ARG;
  if (which == 3) return host_stream_matvec_P(&h->h, in, out);
  if (which == 4) return host_panel_matvec_P(&h->h, in, out);
  if (which == 5) return host_small_matvec_P(&h->h, in, out);
  host_matvec(which == 0 ? h->h.Ab : (which == 1 ? h->h.At : h->h.Pm), in, out);
  return BQP_OK;
}

}  // extern "C"

extern "C" int bqp_debug_dump_groups(bqp_handle h) {
  if (!h) return BQP_E_ARG;
  const HostStream &st = h->h.st;
  std::printf("built=%d groups=%zu slot_bytes=%d iter_bytes=%lld\n", (int)st.built, st.groups.size(), st.slot_bytes, st.iter_bytes);
  for (size_t g = 0; g < st.groups.size(); g++) {
    const StreamGroup &G = st.groups[g];
    std::printf("g%2zu kind=%d row0=%4d nsl=%2d sparse=%d qch=[%d %d %d %d] qcol0=[%d %d %d %d] off=%lld\n", g, G.kind, G.row0, G.nsl,
                G.sparse, G.qch[0], G.qch[1], G.qch[2], G.qch[3], G.qcol0[0], G.qcol0[1], G.qcol0[2], G.qcol0[3], G.data_off);
  }
  return BQP_OK;
}
