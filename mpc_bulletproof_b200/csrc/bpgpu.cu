// libbpgpu: C ABI (include/bpgpu.h) over the sm_100a kernels.
#include "../../include/bpgpu.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <unordered_map>
#include <vector>

#include "msm_kernels.cuh"
#include "ipp_kernels.cuh"
#include "svec_kernels.cuh"
#include "stark_msm.cuh"
#include "ipp_kernels_t.cuh"

using namespace bpg;

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
struct bpg_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  int last_cuda = 0;
  uint64_t launches = 0;
  int forced_c = 0;
  int forced_gsub = 0;  // BPG_MSM_GSUB: bucket groups per set on windowed tables (tuning)
  int sm_count = 148;
  // per-phase device timing (bpg_profile_*): events are recorded on the launch stream
  bool prof = false;
  std::vector<cudaEvent_t> prof_ev;   // pool
  std::vector<int> prof_phase;        // phase id of interval [ev[i], ev[i+1])
  size_t prof_used = 0;
  double prof_ms[BPG_PROF_NPHASE] = {0};
  uint64_t prof_n[BPG_PROF_NPHASE] = {0};
  // workspace arenas (grown on demand, reused across calls): [0] for the launch stream, [1] for the
  // auxiliary stream that runs a small independent MSM beside the main one (verifier: proof points)
  uint8_t* ws = nullptr;
  size_t ws_cap = 0;
  uint8_t* ws_aux = nullptr;
  size_t ws_aux_cap = 0;
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_chunk[8] = {};  // scalar upload in pieces (host_feed), created on first use
  // transient-allocation cache (dev_alloc / dev_free)
  std::vector<std::pair<void*, size_t>> cache;
  std::unordered_map<void*, size_t> live;
  // small staging buffers
  uint8_t* d_small = nullptr;   // device scratch for results (>= 64 KB)
  uint8_t* h_pinned = nullptr;  // pinned host scratch (>= 64 KB)
  // staging for host-buffer calls
  uint8_t* d_stage = nullptr;
  size_t d_stage_cap = 0;
};

struct bpg_table {
  bpg_ctx* ctx;
  uint32_t* niels;  // n * 24 words; windowed: [W][n] * 24 words
  size_t n;
  int win_c = 0;    // 0: plain; otherwise the window width the multiples 2^(c w) P_i were built for
  int win_W = 1;
};

#define CK(call)                                  \
  do {                                            \
    cudaError_t e_ = (call);                      \
    if (e_ != cudaSuccess) {                      \
      ctx->last_cuda = (int)e_;                   \
      return BPG_ERR_CUDA;                        \
    }                                             \
  } while (0)

#define LAUNCH_CHECK()                            \
  do {                                            \
    ctx->launches++;                              \
    cudaError_t e_ = cudaGetLastError();          \
    if (e_ != cudaSuccess) {                      \
      ctx->last_cuda = (int)e_;                   \
      return BPG_ERR_CUDA;                        \
    }                                             \
  } while (0)

// Transient device memory (tables, proof states).  Freed blocks are parked in a small per-context
// cache and handed out again (best fit within 4x) before falling back to the device's stream-ordered
// pool: a proof allocates its state per call, and neither a cudaMalloc nor a pool miss (both cost
// milliseconds at these sizes) may sit on that path.  Everything a context allocates is used on its
// launch stream (the auxiliary lane is fenced by events), so reuse is ordered by the stream.
static constexpr size_t CACHE_SLOTS = 24;
static cudaError_t dev_alloc(bpg_ctx* ctx, void** p, size_t bytes) {
  bytes = std::max<size_t>((bytes + 255) / 256 * 256, 256);
  int best = -1;
  for (size_t i = 0; i < ctx->cache.size(); i++) {
    size_t sz = ctx->cache[i].second;
    if (sz >= bytes && sz <= 4 * bytes && (best < 0 || sz < ctx->cache[best].second)) best = (int)i;
  }
  if (best >= 0) {
    *p = ctx->cache[best].first;
    ctx->live[*p] = ctx->cache[best].second;
    ctx->cache.erase(ctx->cache.begin() + best);
    return cudaSuccess;
  }
  cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
  if (e == cudaSuccess) ctx->live[*p] = bytes;
  return e;
}
template <typename T>
static cudaError_t dev_alloc(bpg_ctx* ctx, T** p, size_t bytes) {
  return dev_alloc(ctx, reinterpret_cast<void**>(p), bytes);
}
static void dev_free(bpg_ctx* ctx, void* p) {
  if (!p) return;
  auto it = ctx->live.find(p);
  size_t sz = it == ctx->live.end() ? 0 : it->second;
  if (it != ctx->live.end()) ctx->live.erase(it);
  if (sz == 0 || sz > ((size_t)1 << 31)) {  // unknown or huge (a user's big table): give it back
    cudaFreeAsync(p, ctx->stream);
    return;
  }
  ctx->cache.push_back({p, sz});
  if (ctx->cache.size() > CACHE_SLOTS) {
    cudaFreeAsync(ctx->cache.front().first, ctx->stream);
    ctx->cache.erase(ctx->cache.begin());
  }
}

static constexpr size_t SMALL_BYTES = 1 << 16;

extern "C" int bpg_init(int device, bpg_ctx** out) {
  if (!out) return BPG_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    return BPG_ERR_CUDA;  // no CUDA device: there is no CPU path
  }
  bpg_ctx* ctx = new (std::nothrow) bpg_ctx();
  if (!ctx) return BPG_ERR_NOMEM;
  ctx->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_small, SMALL_BYTES);
  if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_pinned, SMALL_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_reduce_pairs_final, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RPB_SMEM);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) {
    delete ctx;
    return BPG_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;  // freed blocks stay in the pool for the next proof
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  const char* env = getenv("BPG_MSM_C");
  if (env) ctx->forced_c = atoi(env);
  env = getenv("BPG_MSM_GSUB");
  if (env) ctx->forced_gsub = atoi(env);
  *out = ctx;
  return BPG_OK;
}

extern "C" void bpg_free(bpg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
  for (auto& b : ctx->cache) cudaFreeAsync(b.first, ctx->stream);
  ctx->cache.clear();
  cudaStreamSynchronize(ctx->stream);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->ws_aux) cudaFree(ctx->ws_aux);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  for (auto& e : ctx->ev_chunk)
    if (e) cudaEventDestroy(e);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->d_small) cudaFree(ctx->d_small);
  if (ctx->d_stage) cudaFree(ctx->d_stage);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

extern "C" int bpg_set_stream(bpg_ctx* ctx, void* s, int use_own) {
  if (!ctx) return BPG_ERR_ARG;
  ctx->stream = use_own ? ctx->own_stream : (cudaStream_t)s;
  return BPG_OK;
}

// ---- per-phase profiling ---------------------------------------------------
static void prof_mark(bpg_ctx* ctx, int phase) {
  // closes the previous interval and opens one attributed to `phase` (-1 = close only)
  if (!ctx->prof) return;
  if (ctx->prof_used == ctx->prof_ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    ctx->prof_ev.push_back(e);
    ctx->prof_phase.push_back(-1);
  }
  cudaEventRecord(ctx->prof_ev[ctx->prof_used], ctx->stream);
  ctx->prof_phase[ctx->prof_used] = phase;
  ctx->prof_used++;
}
static void prof_collect(bpg_ctx* ctx) {
  if (ctx->prof_used == 0) return;
  cudaEventSynchronize(ctx->prof_ev[ctx->prof_used - 1]);
  for (size_t i = 0; i + 1 < ctx->prof_used; i++) {
    int ph = ctx->prof_phase[i];
    if (ph < 0 || ph >= BPG_PROF_NPHASE) continue;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->prof_ev[i], ctx->prof_ev[i + 1]) == cudaSuccess) {
      ctx->prof_ms[ph] += ms;
      ctx->prof_n[ph]++;
    }
  }
  ctx->prof_used = 0;
}
extern "C" int bpg_profile_enable(bpg_ctx* ctx, int on) {
  if (!ctx) return BPG_ERR_ARG;
  if (ctx->prof) prof_collect(ctx);
  ctx->prof = on != 0;
  return BPG_OK;
}
extern "C" int bpg_profile_reset(bpg_ctx* ctx) {
  if (!ctx) return BPG_ERR_ARG;
  prof_collect(ctx);
  for (int i = 0; i < BPG_PROF_NPHASE; i++) ctx->prof_ms[i] = 0, ctx->prof_n[i] = 0;
  return BPG_OK;
}
extern "C" int bpg_profile_read(bpg_ctx* ctx, double* ms, uint64_t* count, int n) {
  if (!ctx || !ms || !count) return BPG_ERR_ARG;
  prof_collect(ctx);
  for (int i = 0; i < n && i < BPG_PROF_NPHASE; i++) ms[i] = ctx->prof_ms[i], count[i] = ctx->prof_n[i];
  return BPG_OK;
}
extern "C" const char* bpg_profile_phase_name(int phase) {
  static const char* names[BPG_PROF_NPHASE] = {"hist",   "scan",    "scatter", "accum", "accum_big",
                                               "reduce", "combine", "horner",  "encode", "other"};
  return (phase >= 0 && phase < BPG_PROF_NPHASE) ? names[phase] : "?";
}
extern "C" int bpg_sync(bpg_ctx* ctx) {
  if (!ctx) return BPG_ERR_ARG;
  CK(cudaStreamSynchronize(ctx->stream));
  return BPG_OK;
}
extern "C" int bpg_last_cuda_error(const bpg_ctx* ctx) { return ctx ? ctx->last_cuda : 0; }
extern "C" uint64_t bpg_launch_count(const bpg_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int bpg_set_window(bpg_ctx* ctx, int c) {
  if (!ctx || c < 0 || c > 24) return BPG_ERR_ARG;
  ctx->forced_c = c;
  return BPG_OK;
}

// Page-locked host buffers.  cudaHostAlloc costs milliseconds per call, and a constraint system grows
// its vectors while it is being built, so freed buffers are parked in a small process-wide cache
// and handed out again (best fit within 4x).  Header: [tag, capacity]; tag says how it was obtained.
namespace {
struct HostCache {
  std::mutex mu;
  std::vector<std::pair<void*, size_t>> free_list;  // raw pointers (header included), capacity
  size_t bytes = 0;
  ~HostCache() {
    for (auto& b : free_list) cudaFreeHost(b.first);
  }
};
HostCache& host_cache() {
  static HostCache* c = new HostCache();  // leaked on purpose: buffers may outlive static destruction order
  return *c;
}
constexpr uint64_t TAG_PINNED = 0x50494e4eull, TAG_MALLOC = 0x4d414c4cull;
constexpr size_t HOST_CACHE_MAX = (size_t)2 << 30, HOST_CACHE_SLOTS = 64;
}  // namespace

extern "C" void* bpg_host_alloc(size_t bytes) {
  size_t total = bytes + 64;
  {
    HostCache& hc = host_cache();
    std::lock_guard<std::mutex> lk(hc.mu);
    int best = -1;
    for (size_t i = 0; i < hc.free_list.size(); i++) {
      size_t cap = hc.free_list[i].second;
      if (cap >= total && cap <= 4 * total && (best < 0 || cap < hc.free_list[best].second)) best = (int)i;
    }
    if (best >= 0) {
      void* p = hc.free_list[best].first;
      hc.bytes -= hc.free_list[best].second;
      hc.free_list.erase(hc.free_list.begin() + best);
      return static_cast<uint8_t*>(p) + 64;
    }
  }
  void* p = nullptr;
  bool pinned = cudaHostAlloc(&p, total, cudaHostAllocDefault) == cudaSuccess;
  if (!pinned) {
    cudaGetLastError();
    p = malloc(total);
    if (!p) return nullptr;
  }
  uint64_t* h = reinterpret_cast<uint64_t*>(p);
  h[0] = pinned ? TAG_PINNED : TAG_MALLOC;
  h[1] = total;
  return static_cast<uint8_t*>(p) + 64;
}
extern "C" void bpg_host_free(void* q) {
  if (!q) return;
  uint8_t* p = static_cast<uint8_t*>(q) - 64;
  uint64_t* h = reinterpret_cast<uint64_t*>(p);
  if (h[0] != TAG_PINNED) {
    free(p);
    return;
  }
  HostCache& hc = host_cache();
  {
    std::lock_guard<std::mutex> lk(hc.mu);
    if (hc.free_list.size() < HOST_CACHE_SLOTS && hc.bytes + h[1] <= HOST_CACHE_MAX) {
      hc.free_list.push_back({p, (size_t)h[1]});
      hc.bytes += h[1];
      return;
    }
  }
  cudaFreeHost(p);
}

extern "C" int bpg_set_groups(bpg_ctx* ctx, int gsub) {
  if (!ctx || gsub < 0 || gsub > 128) return BPG_ERR_ARG;
  ctx->forced_gsub = gsub;
  return BPG_OK;
}

extern "C" const char* bpg_strerror(int code) {
  switch (code) {
    case BPG_OK: return "ok";
    case BPG_ERR_ARG: return "bad argument";
    case BPG_ERR_LEN: return "vector lengths differ";
    case BPG_ERR_POW2: return "length is not a power of two";
    case BPG_ERR_CAPACITY: return "generator table too short (InvalidGeneratorsLength)";
    case BPG_ERR_DECODE: return "invalid encoding (FormatError)";
    case BPG_ERR_VERIFY: return "verification failed (VerificationError)";
    case BPG_ERR_CUDA: return "CUDA error or no CUDA device (no CPU fallback exists)";
    case BPG_ERR_NOMEM: return "out of memory";
    default: return "unknown error";
  }
}

static int ensure_ws(bpg_ctx* ctx, size_t bytes, int lane = 0) {
  uint8_t*& ws = lane ? ctx->ws_aux : ctx->ws;
  size_t& cap = lane ? ctx->ws_aux_cap : ctx->ws_cap;
  if (bytes <= cap) return BPG_OK;
  // the arena may still be in use by enqueued work
  CK(cudaStreamSynchronize(lane ? ctx->aux_stream : ctx->stream));
  if (ws) cudaFree(ws);
  ws = nullptr;
  cap = 0;
  size_t want = bytes + bytes / 8;
  cudaError_t e = cudaMalloc(&ws, want);
  if (e != cudaSuccess) {
    ctx->last_cuda = (int)e;
    return e == cudaErrorMemoryAllocation ? BPG_ERR_NOMEM : BPG_ERR_CUDA;
  }
  cap = want;
  return BPG_OK;
}
static int ensure_stage(bpg_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->d_stage_cap) return BPG_OK;
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->d_stage) cudaFree(ctx->d_stage);
  ctx->d_stage = nullptr;
  ctx->d_stage_cap = 0;
  cudaError_t e = cudaMalloc(&ctx->d_stage, bytes);
  if (e != cudaSuccess) {
    ctx->last_cuda = (int)e;
    return e == cudaErrorMemoryAllocation ? BPG_ERR_NOMEM : BPG_ERR_CUDA;
  }
  ctx->d_stage_cap = bytes;
  return BPG_OK;
}

// ---------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------
static int table_from_dev(bpg_ctx* ctx, const uint8_t* d_comp, size_t n, bpg_table** out) {
  bpg_table* t = new (std::nothrow) bpg_table();
  if (!t) return BPG_ERR_NOMEM;
  t->ctx = ctx;
  t->n = n;
  t->niels = nullptr;
  cudaError_t e = dev_alloc(ctx, &t->niels, std::max<size_t>(n, 1) * 96);
  if (e != cudaSuccess) {
    delete t;
    ctx->last_cuda = (int)e;
    return BPG_ERR_NOMEM;
  }
  uint32_t* bad = reinterpret_cast<uint32_t*>(ctx->d_small);
  int rc = BPG_OK;
  do {
    if (cudaMemsetAsync(bad, 0, 4, ctx->stream) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
    if (n) {
      k_decode_to_niels<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d_comp, (uint32_t)n, t->niels, bad);
      ctx->launches++;
      if (cudaGetLastError() != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
    }
    uint32_t* hbad = reinterpret_cast<uint32_t*>(ctx->h_pinned);
    if (cudaMemcpyAsync(hbad, bad, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
    cudaError_t se = cudaStreamSynchronize(ctx->stream);
    if (se != cudaSuccess) { ctx->last_cuda = (int)se; rc = BPG_ERR_CUDA; break; }
    if (*hbad) rc = BPG_ERR_DECODE;
  } while (0);
  if (rc != BPG_OK) {
    dev_free(ctx, t->niels);
    delete t;
    return rc;
  }
  *out = t;
  return BPG_OK;
}

extern "C" int bpg_table_upload_dev(bpg_ctx* ctx, const void* d_comp, size_t n, bpg_table** out) {
  if (!ctx || !out || (!d_comp && n) || n >= (1u << 31)) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  return table_from_dev(ctx, (const uint8_t*)d_comp, n, out);
}

extern "C" int bpg_table_upload(bpg_ctx* ctx, const uint8_t* comp, size_t n, bpg_table** out) {
  if (!ctx || !out || (!comp && n) || n >= (1u << 31)) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_stage(ctx, std::max<size_t>(n, 1) * 32);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->d_stage, comp, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  return table_from_dev(ctx, ctx->d_stage, n, out);
}

extern "C" size_t bpg_table_len(const bpg_table* t) { return t ? t->n : 0; }
extern "C" void bpg_table_free(bpg_table* t) {
  if (!t) return;
  cudaSetDevice(t->ctx->device);
  dev_free(t->ctx, t->niels);  // stream-ordered: work already enqueued on the table finishes first
  delete t;
}

// ---------------------------------------------------------------------------
// MSM launch
// ---------------------------------------------------------------------------
// Plain tables: every window has its own bucket array, reduced separately, then Horner.
static int pick_window(size_t n_per_set, int forced) {
  if (forced >= 2) return forced;
  double best = 1e300;
  int best_c = 4;
  for (int c = 3; c <= 20; c++) {
    int W = (255 + c - 1) / c;
    double nb = (double)(1u << (c - 1));
    // mixed adds (7M) for the terms, ~20M per bucket in the reduction tree
    double cost = W * ((double)n_per_set * 7.0 + nb * 20.0);
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}
// Windowed tables: all windows share one bucket array per set.  Lists of ~32 entries keep the
// accumulation efficient; shorter lists only add merge work.
static uint32_t pick_gsub(size_t n_per_set, int nsets, int c, int W) {
  double nb = (double)(1u << (c - 1));
  double lists_for_len = (double)n_per_set * W / (nb * 32.0);       // groups that make the lists ~32 long
  double lists_for_par = (double)(1u << 17) / (nb * (double)nsets);  // groups that give ~2^17 lists (more only add merge work)
  double avg1 = (double)n_per_set * W / nb;                          // list length with one group
  double g = std::max(lists_for_len, std::min(lists_for_par, avg1 / 8.0));  // never below ~8 entries per list
  uint32_t gs = (uint32_t)(g + 0.5);
  return std::min<uint32_t>(std::max<uint32_t>(gs, 1), (uint32_t)W);
}
static int pick_window_table(size_t n, int forced) {
  if (forced >= 2) return forced;
  double best = 1e300;
  int best_c = 4;
  for (int c = 3; c <= 20; c++) {
    int W = (255 + c - 1) / c;
    double nb = (double)(1u << (c - 1));
    uint32_t gs = pick_gsub(n, 1, c, W);
    double cost = (double)W * n * 7.0 + (gs > 1 ? gs * nb * 8.0 : 0.0) + nb * 20.0;
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

static sc_bias bias_for(int c);
static void make_cfg(MsmCfg& cfg, size_t n_terms, size_t n_points, int nsets, int c, size_t win_stride, int forced_gsub) {
  cfg.win_stride = (uint32_t)win_stride;
  cfg.c = c;
  cfg.W = (255 + c - 1) / c;
  cfg.nb = 1u << (c - 1);
  cfg.nsets = nsets;
  cfg.n_terms = (uint32_t)n_terms;
  cfg.n_points = (uint32_t)std::max<size_t>(n_points, 1);
  if (win_stride) {
    cfg.gsub = forced_gsub > 0 ? std::min<uint32_t>((uint32_t)forced_gsub, (uint32_t)cfg.W)
                               : pick_gsub((n_terms + nsets - 1) / nsets, nsets, c, cfg.W);
  } else {
    cfg.gsub = (uint32_t)cfg.W;
  }
  cfg.narr = (uint32_t)nsets * cfg.gsub;
  cfg.B = cfg.narr * cfg.nb;
  // segments of over-long buckets (> BIG_SEG entries): at most two per BIG_SEG entries
  cfg.big_cap = (uint32_t)(2 * ((uint64_t)n_terms * cfg.W / BIG_SEG) + 2);
  memset(&cfg.bias, 0, sizeof(cfg.bias));
  for (int w = 0; w < cfg.W; w++) {
    int bit = c * w + c - 1;
    cfg.bias.v[bit >> 5] |= 1u << (bit & 31);
  }
}

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Enqueue one Pippenger launch.  d_scalars: n_terms*32 B; d_set_ids / d_point_ids may
// be null (implicit: term t -> point t % n_points of `table_base`, set t / n_points).
static int msm_enqueue(bpg_ctx* ctx, const uint32_t* table_base, size_t n_points, const uint32_t* d_scalars,
                       size_t n_terms, const uint8_t* d_set_ids, const uint32_t* d_point_ids, int nsets,
                       uint32_t* d_out_ext, int win_c = 0, size_t win_stride = 0, int lane = 0, int curve = 0,
                       const uint8_t* h_scalars = nullptr /*scalars still on the host: uploaded here, in pieces*/) {
  // curve 0: ristretto255 (Niels table, 24 words per entry); curve 1: Stark curve (affine table, 16 words
  // per entry, plain tables only).  Sort and schedule are shared; the bucket arithmetic differs.
  if (nsets <= 0) return BPG_ERR_ARG;
  // lane 1: the auxiliary stream and arena (no phase profiling there)
  struct ProfOff {
    bpg_ctx* c;
    bool saved;
    ProfOff(bpg_ctx* c_, bool off) : c(c_), saved(c_->prof) { if (off) c->prof = false; }
    ~ProfOff() { c->prof = saved; }
  } prof_off(ctx, lane != 0);
  cudaStream_t st = lane ? ctx->aux_stream : ctx->stream;
  uint8_t* const& ws = lane ? ctx->ws_aux : ctx->ws;
  if (n_terms == 0) {
    // empty sum: identity for every set
    if (curve == 1) k_stark_set_identity<<<nsets, 32, 0, st>>>(d_out_ext);
    else k_set_identity<<<nsets, 32, 0, st>>>(d_out_ext);
    LAUNCH_CHECK();
    return BPG_OK;
  }
  if (n_terms >= (1u << 31)) return BPG_ERR_ARG;
  if (curve == 0 && win_c == 0 && n_terms <= SMALL_MAX_TERMS && nsets <= SMALL_MAX_SETS && ctx->forced_c < 2) {
    // a handful of terms over a plain table: one quad per term walks the doubling chain (k_msm_small);
    // a forced window width (bpg_set_window) keeps the bucket pipeline, which is how the tests reach it
    if (h_scalars) CK(cudaMemcpyAsync((void*)d_scalars, h_scalars, n_terms * 32, cudaMemcpyHostToDevice, st));
    unsigned nblk = (unsigned)((n_terms + SMALL_QUADS - 1) / SMALL_QUADS);
    prof_mark(ctx, BPG_PROF_ACCUM);
    if (nblk == 1) {
      k_msm_small<<<1, SMALL_THREADS, 0, st>>>(table_base, d_scalars, d_set_ids, d_point_ids, (uint32_t)n_terms,
                                               (uint32_t)std::max<size_t>(n_points, 1), nsets, bias_for(4), d_out_ext);
      LAUNCH_CHECK();
    } else {
      int rc = ensure_ws(ctx, (size_t)nblk * nsets * 128, lane);
      if (rc) return rc;
      uint32_t* parts = (uint32_t*)ws;
      k_msm_small<<<nblk, SMALL_THREADS, 0, st>>>(table_base, d_scalars, d_set_ids, d_point_ids, (uint32_t)n_terms,
                                                  (uint32_t)std::max<size_t>(n_points, 1), nsets, bias_for(4), parts);
      LAUNCH_CHECK();
      k_msm_small_fin<<<1, SMALL_THREADS, 0, st>>>(parts, nblk, nsets, d_out_ext);
      LAUNCH_CHECK();
    }
    prof_mark(ctx, -1);
    return BPG_OK;
  }
  MsmCfg cfg;
  int c = win_c ? win_c : pick_window((n_terms + nsets - 1) / nsets, ctx->forced_c);
  make_cfg(cfg, n_terms, n_points, nsets, c, win_c ? win_stride : 0, ctx->forced_gsub);
  if ((uint64_t)cfg.narr * cfg.nb >= (1ull << 31)) return BPG_ERR_ARG;
  const bool windowed = cfg.win_stride != 0;

  // reduction geometry: `rarr` arrays of nb buckets.  Small arrays: a leaf block of RT_QUADS quad
  // chunks of LC buckets plus its in-block tree; large arrays: one thread per chunk of 16.
  uint32_t rarr = windowed ? (uint32_t)nsets : cfg.narr;
  // (measured at 2 x 2^16 buckets, an IPP round at n = 2^16: quad chunks of 8 and thread chunks of 4 tie,
  // thread chunks of 8 / 16 and quad chunks of 4 are 2-11 % of a proof slower; gpurun_out/r1i_tune.jsonl)
  const bool thread_leaf = cfg.nb >= (1u << 17);
  const uint32_t LC = thread_leaf ? 16 : (cfg.nb > (1u << 15) ? 8 : 4);
  uint32_t tiles0 = thread_leaf ? (cfg.nb + LC - 1) / LC : (cfg.nb + RT_QUADS * LC - 1) / (RT_QUADS * LC);
  size_t ntiles = (cfg.B + SCAN_TILE - 1) / SCAN_TILE;
  size_t off = 0;
  // counts and the schedule's control words are adjacent: ONE memset per launch
  size_t o_counts = off;  off += align_up((size_t)cfg.B * 4);
  size_t o_bins = off;    off += align_up((2 * SIZE_BINS + 4) * 4);  // bins | n_items, part, multi, big_count | cursors
  size_t o_offsets = off; off += align_up(((size_t)cfg.B + 1) * 4);
  size_t o_tiles = off;   off += align_up(ntiles * 4);
  size_t o_big = off;     off += align_up(3 * (size_t)cfg.big_cap * 4);
  size_t o_bigpart = off; off += align_up((size_t)cfg.big_cap * 128);
  size_t o_entries = off; off += align_up((size_t)n_terms * cfg.W * 4);
  size_t o_buckets = off; off += align_up((size_t)cfg.B * 128);
  size_t o_merged = off;  off += (windowed && cfg.gsub > 1) ? align_up((size_t)nsets * cfg.nb * 128) : 0;
  if (curve == 1) {
    rarr = windowed ? (uint32_t)nsets : cfg.narr;
    tiles0 = (cfg.nb + SLEAF_LC - 1) / SLEAF_LC;
  }
  size_t o_pairs = off;   off += 4 * align_up((size_t)rarr * tiles0 * 128);  // (A, Y) x ping-pong
  size_t o_wins = off;    off += align_up((size_t)rarr * 128);
  // accumulation schedule: at most one item per bucket plus one per ACC_SEG entries
  size_t max_items = (size_t)cfg.B + (size_t)n_terms * cfg.W / ACC_SEG + 1;
  size_t max_multi = (size_t)n_terms * cfg.W / ACC_SEG + 1;  // buckets longer than ACC_SEG
  size_t o_items = off;   off += align_up(max_items * 8);
  size_t o_segslot = off; off += align_up((size_t)cfg.B * 4);
  size_t o_multi = off;   off += align_up(max_multi * 4);
  size_t o_segpart = off; off += align_up(2 * max_multi * 128);  // sum of nseg over multi-segment buckets <= 2 max_multi
  int rc = ensure_ws(ctx, off, lane);
  if (rc) return rc;
  uint32_t* counts = (uint32_t*)(ws + o_counts);
  uint32_t* offsets = (uint32_t*)(ws + o_offsets);
  uint32_t* tiles = (uint32_t*)(ws + o_tiles);
  uint32_t* big_list = (uint32_t*)(ws + o_big);
  uint32_t* big_part = (uint32_t*)(ws + o_bigpart);
  uint32_t* entries = (uint32_t*)(ws + o_entries);
  uint32_t* buckets = (uint32_t*)(ws + o_buckets);
  uint32_t* merged = (uint32_t*)(ws + o_merged);
  size_t pair_words = align_up((size_t)rarr * tiles0 * 128) / 4;
  uint32_t* pairs = (uint32_t*)(ws + o_pairs);
  uint32_t* wins = (uint32_t*)(ws + o_wins);
  uint32_t* bins = (uint32_t*)(ws + o_bins);
  AccSched sched;
  uint32_t* big_count = bins + SIZE_BINS + 3;
  sched.bins = bins;
  sched.cursors = bins + SIZE_BINS + 4;
  sched.n_items = bins + SIZE_BINS;
  sched.part_count = bins + SIZE_BINS + 1;
  sched.multi_count = bins + SIZE_BINS + 2;
  sched.items = (uint2*)(ws + o_items);
  sched.seg_slot = (uint32_t*)(ws + o_segslot);
  sched.multi_list = (uint32_t*)(ws + o_multi);
  uint32_t* seg_part = (uint32_t*)(ws + o_segpart);

  prof_mark(ctx, BPG_PROF_HIST);
  CK(cudaMemsetAsync(counts, 0, o_bins - o_counts + (2 * SIZE_BINS + 4) * 4, st));
  unsigned gt = (unsigned)((n_terms + 255) / 256);
  if (h_scalars && lane == 0 && n_terms >= (1u << 18)) {
    // Host scalars: the copy runs on the auxiliary stream in pieces and the digit histogram of
    // piece i runs while piece i+1 is still on the bus (hides the 0.12 ms histogram of a 2^20-term
    // launch; measured against one copy on two boxes: 2.26-2.45 vs 2.39-2.62 ms end to end, the
    // spread being the PCIe rate of the box).
    const int pieces = 4;
    size_t per = (((n_terms + pieces - 1) / pieces) + 255) / 256 * 256;
    CK(cudaEventRecord(ctx->ev_fork, st));  // d_scalars (staging) is free once earlier work is done
    CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    for (int i = 0; i < pieces; i++) {
      size_t t0 = (size_t)i * per, t1 = std::min(n_terms, t0 + per);
      if (t0 >= t1) break;
      if (!ctx->ev_chunk[i]) CK(cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming));
      CK(cudaMemcpyAsync((uint8_t*)d_scalars + t0 * 32, h_scalars + t0 * 32, (t1 - t0) * 32, cudaMemcpyHostToDevice,
                         ctx->aux_stream));
      CK(cudaEventRecord(ctx->ev_chunk[i], ctx->aux_stream));
      CK(cudaStreamWaitEvent(st, ctx->ev_chunk[i], 0));
      k_hist<<<(unsigned)((t1 - t0 + 255) / 256), 256, 0, st>>>(d_scalars, d_set_ids, cfg, counts, (uint32_t)t0, (uint32_t)t1);
      LAUNCH_CHECK();
    }
  } else {
    if (h_scalars) CK(cudaMemcpyAsync((void*)d_scalars, h_scalars, n_terms * 32, cudaMemcpyHostToDevice, st));
    k_hist<<<gt, 256, 0, st>>>(d_scalars, d_set_ids, cfg, counts, 0u, (uint32_t)n_terms);
    LAUNCH_CHECK();
  }
  prof_mark(ctx, BPG_PROF_SCAN);
  k_scan_tiles<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(counts, cfg.B, tiles);
  LAUNCH_CHECK();
  k_scan_spine<<<1, 1024, 0, st>>>(tiles, (uint32_t)ntiles, offsets, cfg.B);
  LAUNCH_CHECK();
  k_scan_apply<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(counts, cfg.B, tiles, offsets, bins);
  LAUNCH_CHECK();
  prof_mark(ctx, BPG_PROF_SCATTER);
  k_scatter<<<gt, 256, 0, st>>>(d_scalars, d_set_ids, d_point_ids, cfg, offsets, counts, entries);
  LAUNCH_CHECK();
  // accumulation schedule: (bucket, segment) items by decreasing length; over-long buckets -> big list
  k_size_scatter<<<(cfg.B + 255) / 256, 256, 0, st>>>(offsets, cfg, sched, big_count, big_list);
  LAUNCH_CHECK();
  if (curve == 1) {
    // ---- Stark-curve policy: same stages, short-Weierstrass bucket arithmetic (stark_msm.cuh) ----
    prof_mark(ctx, BPG_PROF_ACCUM);
    k_stark_accum<<<(unsigned)((max_items + SACC_THREADS - 1) / SACC_THREADS), SACC_THREADS, 0, st>>>(
        table_base, offsets, entries, sched, buckets, seg_part);
    LAUNCH_CHECK();
    prof_mark(ctx, BPG_PROF_ACCUM_BIG);
    k_stark_fix<<<(unsigned)std::min<size_t>((max_multi + 127) / 128, (size_t)ctx->sm_count * 8), 128, 0, st>>>(
        offsets, sched, seg_part, buckets);
    LAUNCH_CHECK();
    unsigned gb = std::min<unsigned>(cfg.big_cap, (unsigned)ctx->sm_count * 4);
    k_stark_big<<<gb, SBIG_THREADS, 0, st>>>(table_base, offsets, entries, cfg, buckets, big_count, big_list, big_part);
    LAUNCH_CHECK();
    k_stark_big_fin<<<std::min<unsigned>((cfg.big_cap + 127) / 128, (unsigned)ctx->sm_count), 128, 0, st>>>(
        cfg, buckets, big_count, big_list, big_part);
    LAUNCH_CHECK();
    const uint32_t* lvl0 = buckets;
    if (windowed && cfg.gsub > 1) {
      prof_mark(ctx, BPG_PROF_COMBINE);
      k_stark_merge<<<((unsigned)nsets * cfg.nb + 127) / 128, 128, 0, st>>>(buckets, cfg, merged);
      LAUNCH_CHECK();
      lvl0 = merged;
    }
    prof_mark(ctx, BPG_PROF_REDUCE);
    uint32_t arrays = windowed ? (uint32_t)nsets : cfg.narr;
    uint32_t* fin = windowed ? d_out_ext : wins;
    // leaf: large arrays one thread per chunk of 8 (throughput), small ones one quad per chunk of 4
    // plus the in-block tree (latency); the pairs levels and Horner are quad-cooperative
    const bool sthread_leaf = cfg.nb >= (1u << 17);
    uint32_t t = sthread_leaf ? (cfg.nb + SLEAF_LC - 1) / SLEAF_LC : (cfg.nb + SRT_QUADS * 4 - 1) / (SRT_QUADS * 4);
    uint32_t* pa[2] = {pairs, pairs + 2 * pair_words};
    int cur = 0;
    uint32_t* oa = t == 1 ? fin : pa[cur];
    if (sthread_leaf) k_stark_leaf<<<(arrays * t + 127) / 128, 128, 0, st>>>(lvl0, cfg.nb, t, arrays, oa, pa[cur] + pair_words);
    else k_stark_leaf4<4><<<arrays * t, SRT_THREADS, 0, st>>>(lvl0, cfg.nb, t, oa, pa[cur] + pair_words);
    LAUNCH_CHECK();
    while (t > 1) {
      uint32_t n = t;
      t = (n + SRP_PAIRS - 1) / SRP_PAIRS;
      const uint32_t* ia = pa[cur];
      const uint32_t* iy = pa[cur] + pair_words;
      cur ^= 1;
      oa = t == 1 ? fin : pa[cur];
      k_stark_pairs4<<<arrays * t, SRP_THREADS, 0, st>>>(ia, iy, n, t, oa, pa[cur] + pair_words);
      LAUNCH_CHECK();
    }
    if (!windowed) {
      prof_mark(ctx, BPG_PROF_HORNER);
      k_stark_horner4<<<nsets, 32, 0, st>>>(wins, cfg, d_out_ext);
      LAUNCH_CHECK();
    }
    prof_mark(ctx, -1);
    return BPG_OK;
  }
  prof_mark(ctx, BPG_PROF_ACCUM);
  k_accum<<<(unsigned)((max_items + ACC_THREADS - 1) / ACC_THREADS), ACC_THREADS, 0, st>>>(table_base, offsets, entries,
                                                                                          sched, buckets, seg_part);
  LAUNCH_CHECK();
  prof_mark(ctx, BPG_PROF_ACCUM_BIG);
  k_accum_fix<<<(unsigned)std::min<size_t>((max_multi * 4 + FIX_THREADS - 1) / FIX_THREADS, (size_t)ctx->sm_count * 8),
                FIX_THREADS, 0, st>>>(offsets, sched, seg_part, buckets);
  LAUNCH_CHECK();
  unsigned gbig = std::min<unsigned>(cfg.big_cap, (unsigned)ctx->sm_count * 4);
  k_accum_big<<<gbig, BIG_THREADS, 0, st>>>(table_base, offsets, entries, cfg, buckets, big_count, big_list, big_part);
  LAUNCH_CHECK();
  k_accum_big_fin<<<gbig, BIG_THREADS, 0, st>>>(cfg, buckets, big_count, big_list, big_part);
  LAUNCH_CHECK();
  const uint32_t* level0 = buckets;
  if (windowed && cfg.gsub > 1) {
    prof_mark(ctx, BPG_PROF_COMBINE);
    unsigned nq = (unsigned)nsets * cfg.nb;
    k_merge<<<(nq * 4 + MERGE_THREADS - 1) / MERGE_THREADS, MERGE_THREADS, 0, st>>>(buckets, cfg, merged);
    LAUNCH_CHECK();
    level0 = merged;
  }
  prof_mark(ctx, BPG_PROF_REDUCE);
  {
    uint32_t* final_out = windowed ? d_out_ext : wins;
    uint32_t t = tiles0;
    uint32_t* pa[2] = {pairs, pairs + 2 * pair_words};
    int cur = 0;
    uint32_t* oa = t == 1 ? final_out : pa[cur];
    if (thread_leaf) {
      k_reduce_leaf_thread<16><<<(rarr * t + RL_THREADS - 1) / RL_THREADS, RL_THREADS, 0, st>>>(level0, cfg.nb, t, rarr, oa,
                                                                                               pa[cur] + pair_words);
    } else if (LC == 8) {
      k_reduce_leaf<8><<<rarr * t, RT_THREADS, 0, st>>>(level0, cfg.nb, t, oa, pa[cur] + pair_words);
    } else {
      k_reduce_leaf<4><<<rarr * t, RT_THREADS, 0, st>>>(level0, cfg.nb, t, oa, pa[cur] + pair_words);
    }
    LAUNCH_CHECK();
    while (t > 1) {
      uint32_t n = t;
      const uint32_t* ia = pa[cur];
      const uint32_t* iy = pa[cur] + pair_words;
      if (n <= RPB_PAIRS) {
        // what is left fits one block per array: finish here
        k_reduce_pairs_final<<<rarr, RPB_THREADS, RPB_SMEM, st>>>(ia, iy, n, final_out);
        LAUNCH_CHECK();
        break;
      }
      t = (n + RP_PAIRS - 1) / RP_PAIRS;
      cur ^= 1;
      oa = t == 1 ? final_out : pa[cur];
      k_reduce_pairs<<<rarr * t, RP_THREADS, 0, st>>>(ia, iy, n, t, oa, pa[cur] + pair_words);
      LAUNCH_CHECK();
    }
  }
  if (!windowed) {
    prof_mark(ctx, BPG_PROF_HORNER);
    k_horner<<<nsets, 32, 0, st>>>(wins, cfg, d_out_ext);
    LAUNCH_CHECK();
  }
  prof_mark(ctx, -1);
  return BPG_OK;
}

extern "C" int bpg_dev_msm_table(bpg_ctx* ctx, const bpg_table* table, size_t offset, size_t n,
                                 const void* d_scalars, int n_sets, void* d_out_ext) {
  if (!ctx || !table || !d_out_ext || (!d_scalars && n) || n_sets <= 0) return BPG_ERR_ARG;
  if (offset + n > table->n) return BPG_ERR_CAPACITY;
  CK(cudaSetDevice(ctx->device));
  return msm_enqueue(ctx, table->niels + offset * 24, n, (const uint32_t*)d_scalars, n * (size_t)n_sets, nullptr,
                     nullptr, n_sets, (uint32_t*)d_out_ext, table->win_c, table->n);
}

extern "C" int bpg_dev_sum_encode(bpg_ctx* ctx, const void* d_parts, int n_parts, int n_sets, void* d_out_bytes,
                                  void* d_out_ext) {
  if (!ctx || !d_parts || n_parts <= 0 || n_sets <= 0) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  prof_mark(ctx, BPG_PROF_ENCODE);
  k_sum_encode<<<n_sets, ENC_THREADS, 0, ctx->stream>>>((const uint32_t*)d_parts, n_parts, n_sets,
                                                           (uint8_t*)d_out_bytes, (uint32_t*)d_out_ext);
  LAUNCH_CHECK();
  prof_mark(ctx, -1);
  return BPG_OK;
}

// ---------------------------------------------------------------------------
// peer exchange: the ranks' partial sums meet in peer-mapped buffers (one process per GPU)
// ---------------------------------------------------------------------------
struct bpg_peer {
  bpg_ctx* ctx;
  int world, rank, max_sets;
  uint8_t* local;      // cudaMalloc: parts [2][world][max_sets][32] words | flags [2][world] | status
  size_t parts_bytes;
  void* opened[8];     // peers' buffers opened through IPC (null for our own)
  PeerPtrs ptrs;
  uint32_t seq;
  bool connected;
};

extern "C" int bpg_peer_create(bpg_ctx* ctx, int world, int rank, int max_sets, bpg_peer** out, uint8_t handle_out[64]) {
  if (!ctx || !out || !handle_out || world < 1 || world > 8 || rank < 0 || rank >= world || max_sets < 1 ||
      max_sets > XCH_THREADS)
    return BPG_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CK(cudaSetDevice(ctx->device));
  bpg_peer* p = new (std::nothrow) bpg_peer();
  if (!p) return BPG_ERR_NOMEM;
  memset(p, 0, sizeof *p);
  p->ctx = ctx;
  p->world = world;
  p->rank = rank;
  p->max_sets = max_sets;
  p->parts_bytes = align_up((size_t)2 * world * max_sets * 128);
  size_t total = p->parts_bytes + align_up((size_t)2 * world * 4) + 256;
  cudaError_t e = cudaMalloc(&p->local, total);  // IPC needs a plain allocation, not the pool
  if (e == cudaSuccess) e = cudaMemset(p->local, 0, total);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->local);
  if (e != cudaSuccess) {
    ctx->last_cuda = (int)e;
    if (p->local) cudaFree(p->local);
    delete p;
    return BPG_ERR_CUDA;
  }
  memcpy(handle_out, &h, 64);
  *out = p;
  return BPG_OK;
}
// handles: world x 64 bytes, in rank order (every rank's bpg_peer_create output, exchanged by the caller)
extern "C" int bpg_peer_connect(bpg_peer* p, const uint8_t* handles) {
  if (!p || !handles) return BPG_ERR_ARG;
  bpg_ctx* ctx = p->ctx;
  CK(cudaSetDevice(ctx->device));
  for (int r = 0; r < p->world; r++) {
    uint8_t* base = p->local;
    if (r != p->rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + 64 * (size_t)r, 64);
      void* ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        ctx->last_cuda = (int)e;
        return BPG_ERR_CUDA;
      }
      p->opened[r] = ptr;
      base = static_cast<uint8_t*>(ptr);
    }
    p->ptrs.parts[r] = reinterpret_cast<uint32_t*>(base);
    p->ptrs.flags[r] = reinterpret_cast<uint32_t*>(base + p->parts_bytes);
  }
  p->connected = true;
  return BPG_OK;
}
extern "C" void bpg_peer_free(bpg_peer* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  for (int r = 0; r < p->world; r++)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->opened[r]);
  cudaFree(p->local);
  delete p;
}
// One kernel: push this rank's partial sums to every rank, wait for all, add, encode.  Every rank
// must call it the same number of times (the step counter is part of the protocol).
extern "C" int bpg_dev_exchange_sum_encode(bpg_ctx* ctx, bpg_peer* p, const void* d_part, int n_sets, void* d_out_bytes,
                                           void* d_out_ext) {
  if (!ctx || !p || !d_part || n_sets <= 0 || n_sets > p->max_sets || !p->connected) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  p->seq++;
  uint32_t* status = reinterpret_cast<uint32_t*>(p->local + p->parts_bytes + align_up((size_t)2 * p->world * 4));
  prof_mark(ctx, BPG_PROF_ENCODE);
  k_exchange_sum_encode<<<1, XCH_THREADS, 0, ctx->stream>>>((const uint32_t*)d_part, p->ptrs, p->world, p->rank, n_sets,
                                                            p->max_sets, p->seq, (uint8_t*)d_out_bytes, (uint32_t*)d_out_ext,
                                                            status);
  LAUNCH_CHECK();
  prof_mark(ctx, -1);
  return BPG_OK;
}
// 0 = every exchange so far completed; 1 = a peer did not show up within the kernel's bound
extern "C" int bpg_peer_status(bpg_peer* p, int* status_out) {
  if (!p || !status_out) return BPG_ERR_ARG;
  bpg_ctx* ctx = p->ctx;
  CK(cudaSetDevice(ctx->device));
  uint32_t* status = reinterpret_cast<uint32_t*>(p->local + p->parts_bytes + align_up((size_t)2 * p->world * 4));
  CK(cudaMemcpyAsync(ctx->h_pinned + 2048, status, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *status_out = (int)*reinterpret_cast<uint32_t*>(ctx->h_pinned + 2048);
  return BPG_OK;
}

extern "C" int bpg_msm_table(bpg_ctx* ctx, const bpg_table* table, size_t offset, size_t n,
                             const uint8_t* scalars_le, int n_sets, uint8_t* out) {
  if (!ctx || !table || !out || (!scalars_le && n) || n_sets <= 0) return BPG_ERR_ARG;
  if ((size_t)n_sets * 160 > SMALL_BYTES) return BPG_ERR_ARG;
  if (offset + n > table->n) return BPG_ERR_CAPACITY;
  CK(cudaSetDevice(ctx->device));
  size_t sbytes = n * (size_t)n_sets * 32;
  int rc = ensure_stage(ctx, std::max<size_t>(sbytes, 32));
  if (rc) return rc;
  uint32_t* d_ext = (uint32_t*)ctx->d_small;
  uint8_t* d_bytes = ctx->d_small + (size_t)n_sets * 128;
  rc = msm_enqueue(ctx, table->niels + offset * 24, n, (const uint32_t*)ctx->d_stage, n * (size_t)n_sets, nullptr,
                   nullptr, n_sets, d_ext, table->win_c, table->n, 0, 0, sbytes ? scalars_le : nullptr);
  if (rc) return rc;
  rc = bpg_dev_sum_encode(ctx, d_ext, 1, n_sets, d_bytes, nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->h_pinned, d_bytes, (size_t)n_sets * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->h_pinned, (size_t)n_sets * 32);
  return BPG_OK;
}

// Partial sums for a caller that combines them itself (a rank of a sharded MSM, a party of the
// MPC prover working on its share): out_ext = n_sets extended points (X|Y|Z|T, 4 x 32 bytes LE).
extern "C" int bpg_msm_table_partial(bpg_ctx* ctx, const bpg_table* table, size_t offset, size_t n,
                                     const uint8_t* scalars_le, int n_sets, uint8_t* out_ext) {
  if (!ctx || !table || !out_ext || (!scalars_le && n) || n_sets <= 0) return BPG_ERR_ARG;
  if ((size_t)n_sets * 128 > SMALL_BYTES) return BPG_ERR_ARG;
  if (offset + n > table->n) return BPG_ERR_CAPACITY;
  CK(cudaSetDevice(ctx->device));
  size_t sbytes = n * (size_t)n_sets * 32;
  int rc = ensure_stage(ctx, std::max<size_t>(sbytes, 32));
  if (rc) return rc;
  if (sbytes) CK(cudaMemcpyAsync(ctx->d_stage, scalars_le, sbytes, cudaMemcpyHostToDevice, ctx->stream));
  uint32_t* d_ext = (uint32_t*)ctx->d_small;
  rc = msm_enqueue(ctx, table->niels + offset * 24, n, (const uint32_t*)ctx->d_stage, n * (size_t)n_sets, nullptr,
                   nullptr, n_sets, d_ext, table->win_c, table->n);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->h_pinned, d_ext, (size_t)n_sets * 128, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  memcpy(out_ext, ctx->h_pinned, (size_t)n_sets * 128);
  return BPG_OK;
}
// out[s] = encode(sum_p parts[p][s]): the combine step after the ranks' / parties' partial sums met
extern "C" int bpg_sum_encode(bpg_ctx* ctx, const uint8_t* parts_ext, int n_parts, int n_sets, uint8_t* out) {
  if (!ctx || !parts_ext || !out || n_parts <= 0 || n_sets <= 0) return BPG_ERR_ARG;
  size_t in_bytes = (size_t)n_parts * n_sets * 128, out_bytes = (size_t)n_sets * 32;
  if (in_bytes + out_bytes > SMALL_BYTES) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  memcpy(ctx->h_pinned, parts_ext, in_bytes);
  CK(cudaMemcpyAsync(ctx->d_small, ctx->h_pinned, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  int rc = bpg_dev_sum_encode(ctx, ctx->d_small, n_parts, n_sets, ctx->d_small + in_bytes, nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->h_pinned, ctx->d_small + in_bytes, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->h_pinned, out_bytes);
  return BPG_OK;
}

extern "C" int bpg_msm(bpg_ctx* ctx, const uint8_t* scalars_le, const uint8_t* points_compressed, size_t n,
                       uint8_t out[32]) {
  if (!ctx || !out || ((!scalars_le || !points_compressed) && n)) return BPG_ERR_ARG;
  bpg_table* t = nullptr;
  int rc = bpg_table_upload(ctx, points_compressed, n, &t);
  if (rc) return rc;
  rc = bpg_msm_table(ctx, t, 0, n, scalars_le, 1, out);
  bpg_table_free(t);
  return rc;
}

// ---------------------------------------------------------------------------
// fixed-base (comb) multiplication
// ---------------------------------------------------------------------------
struct bpg_comb {
  bpg_ctx* ctx;
  uint32_t* tables;  // nbases * COMB_ENTRIES * 24 words
  int nbases;
};

extern "C" int bpg_comb_create(bpg_ctx* ctx, const uint8_t* bases_compressed, int nbases, bpg_comb** out) {
  if (!ctx || !bases_compressed || nbases <= 0 || nbases > 64 || !out) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  bpg_comb* c = new (std::nothrow) bpg_comb();
  if (!c) return BPG_ERR_NOMEM;
  c->ctx = ctx;
  c->nbases = nbases;
  c->tables = nullptr;
  cudaError_t e = cudaMalloc(&c->tables, (size_t)nbases * COMB_ENTRIES * 96);
  if (e != cudaSuccess) {
    delete c;
    ctx->last_cuda = (int)e;
    return BPG_ERR_NOMEM;
  }
  uint32_t* bad = reinterpret_cast<uint32_t*>(ctx->d_small);
  uint8_t* d_bases = ctx->d_small + 256;
  int rc = BPG_OK;
  do {
    memcpy(ctx->h_pinned + 256, bases_compressed, (size_t)nbases * 32);
    if (cudaMemsetAsync(bad, 0, 4, ctx->stream) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
    if (cudaMemcpyAsync(d_bases, ctx->h_pinned + 256, (size_t)nbases * 32, cudaMemcpyHostToDevice, ctx->stream) !=
        cudaSuccess) { rc = BPG_ERR_CUDA; break; }
    for (int t = 0; t < nbases; t++) {
      k_comb_build<<<1, COMB_WINDOWS, 0, ctx->stream>>>(d_bases + 32 * t, c->tables + (size_t)t * COMB_ENTRIES * 24, bad);
      ctx->launches++;
    }
    if (cudaGetLastError() != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
    uint32_t* hbad = reinterpret_cast<uint32_t*>(ctx->h_pinned);
    if (cudaMemcpyAsync(hbad, bad, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
    cudaError_t se = cudaStreamSynchronize(ctx->stream);
    if (se != cudaSuccess) { ctx->last_cuda = (int)se; rc = BPG_ERR_CUDA; break; }
    if (*hbad) rc = BPG_ERR_DECODE;
  } while (0);
  if (rc != BPG_OK) {
    cudaFree(c->tables);
    delete c;
    return rc;
  }
  *out = c;
  return BPG_OK;
}

extern "C" void bpg_comb_free(bpg_comb* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->ctx->stream);
  cudaFree(c->tables);
  delete c;
}

static sc_bias bias_for(int c) {
  sc_bias b;
  memset(&b, 0, sizeof b);
  int W = (255 + c - 1) / c;
  for (int w = 0; w < W; w++) {
    int bit = c * w + c - 1;
    b.v[bit >> 5] |= 1u << (bit & 31);
  }
  return b;
}

extern "C" int bpg_dev_comb_mul(bpg_ctx* ctx, const bpg_comb* comb, const void* d_scalars, size_t n,
                                void* d_out_bytes, void* d_out_ext) {
  if (!ctx || !comb || (!d_scalars && n) || n >= (1u << 31)) return BPG_ERR_ARG;
  if (n == 0) return BPG_OK;
  CK(cudaSetDevice(ctx->device));
  if (n <= 2048) {
    // latency form: one warp per output
    k_comb_mul_warp<<<(unsigned)((n * 32 + 127) / 128), 128, 0, ctx->stream>>>(comb->tables, comb->nbases,
                                                                              (const uint32_t*)d_scalars, (uint32_t)n,
                                                                              bias_for(4), (uint8_t*)d_out_bytes,
                                                                              (uint32_t*)d_out_ext);
  } else {
    k_comb_mul<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(comb->tables, comb->nbases,
                                                                    (const uint32_t*)d_scalars, (uint32_t)n,
                                                                    bias_for(4), (uint8_t*)d_out_bytes,
                                                                    (uint32_t*)d_out_ext);
  }
  LAUNCH_CHECK();
  return BPG_OK;
}

extern "C" int bpg_comb_mul(bpg_ctx* ctx, const bpg_comb* comb, const uint8_t* scalars_le, size_t n, uint8_t* out) {
  if (!ctx || !comb || !out || (!scalars_le && n)) return BPG_ERR_ARG;
  if (n == 0) return BPG_OK;
  CK(cudaSetDevice(ctx->device));
  size_t sbytes = n * (size_t)comb->nbases * 32;
  int rc = ensure_stage(ctx, sbytes + n * 32);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->d_stage, scalars_le, sbytes, cudaMemcpyHostToDevice, ctx->stream));
  uint8_t* d_out = ctx->d_stage + sbytes;
  rc = bpg_dev_comb_mul(ctx, comb, ctx->d_stage, n, d_out, nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(out, d_out, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return BPG_OK;
}

// Element derivation for generator chains: out[i] = from_uniform_bytes(uniform[64 i .. 64 i + 64)).
extern "C" int bpg_points_from_uniform(bpg_ctx* ctx, const uint8_t* uniform64, size_t n, uint8_t* out_compressed) {
  if (!ctx || (n && (!uniform64 || !out_compressed))) return BPG_ERR_ARG;
  if (n == 0) return BPG_OK;
  if (n >= (1ull << 31)) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_stage(ctx, n * 96);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->d_stage, uniform64, n * 64, cudaMemcpyHostToDevice, ctx->stream));
  uint8_t* d_out = ctx->d_stage + n * 64;
  k_from_uniform<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_stage, (uint32_t)n, d_out);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(out_compressed, d_out, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return BPG_OK;
}

// ---------------------------------------------------------------------------
// windowed tables
// ---------------------------------------------------------------------------
extern "C" int bpg_table_set_windows(bpg_ctx* ctx, bpg_table* t, int c) {
  if (!ctx || !t || c < 0 || c > 20) return BPG_ERR_ARG;
  if (t->win_c) return BPG_ERR_ARG;  // already windowed
  CK(cudaSetDevice(ctx->device));
  if (c == 0) c = pick_window_table(t->n, ctx->forced_c);
  if (c < 2) c = 2;
  int W = (255 + c - 1) / c;
  if ((uint64_t)W * t->n >= (1ull << 31)) return BPG_ERR_ARG;
  if (t->n == 0) {
    t->win_c = c;
    t->win_W = W;
    return BPG_OK;
  }
  uint32_t* out = nullptr;
  cudaError_t e = dev_alloc(ctx, &out, (size_t)W * t->n * 96);
  if (e != cudaSuccess) {
    ctx->last_cuda = (int)e;
    return BPG_ERR_NOMEM;
  }
  const size_t CH = 1 << 16;
  size_t chunk = std::min(CH, t->n);
  size_t ext_bytes = align_up((size_t)(W - 1) * chunk * 128);
  size_t zp_bytes = align_up((size_t)(W - 1) * chunk * 32);
  int rc = ensure_ws(ctx, ext_bytes + zp_bytes);
  if (rc) {
    dev_free(ctx, out);
    return rc;
  }
  for (size_t first = 0; first < t->n; first += chunk) {
    size_t cnt = std::min(chunk, t->n - first);
    k_window_chain<<<(unsigned)((cnt + 127) / 128), 128, 0, ctx->stream>>>(
        t->niels, (uint32_t)t->n, (uint32_t)first, (uint32_t)cnt, c, W, (uint32_t*)ctx->ws,
        (uint32_t*)(ctx->ws + ext_bytes), out);
    ctx->launches++;
  }
  cudaError_t se = cudaStreamSynchronize(ctx->stream);
  if (se != cudaSuccess || cudaGetLastError() != cudaSuccess) {
    ctx->last_cuda = (int)se;
    dev_free(ctx, out);
    return BPG_ERR_CUDA;
  }
  dev_free(ctx, t->niels);
  t->niels = out;
  t->win_c = c;
  t->win_W = W;
  return BPG_OK;
}
extern "C" int bpg_table_window(const bpg_table* t) { return t ? t->win_c : 0; }

// ---------------------------------------------------------------------------
// inner-product argument: device-resident state, one MSM per round
// ---------------------------------------------------------------------------
struct bpg_ipp {
  bpg_ctx* ctx;
  size_t n;        // original length (power of two)
  size_t m;        // current length
  const bpg_table* tab;  // windowed table the round MSMs run over
  bpg_table* own_tab;    // non-null when the state built its own [G | H | Q] table
  bool has_qmul;         // cross terms are multiplied by q_mul (Q = q_mul * table[q_id])
  bool q_sep;            // Q is outside the (caller's windowed) table: c_L Q, c_R Q come from a comb of Q
  uint32_t *q_comb, *q_side;
  uint8_t* buf;    // one allocation for everything below
  uint32_t *a, *b, *wG, *wH, *scalars, *point_ids, *partials, *u_pair, *out_ext;
  uint32_t* q_mul;
  uint8_t *set_ids, *out_bytes;
  bool lr_done;
};

static int table_alloc_plain(bpg_ctx* ctx, size_t n, bpg_table** out) {
  bpg_table* t = new (std::nothrow) bpg_table();
  if (!t) return BPG_ERR_NOMEM;
  t->ctx = ctx;
  t->n = n;
  cudaError_t e = dev_alloc(ctx, &t->niels, std::max<size_t>(n, 1) * 96);
  if (e != cudaSuccess) {
    delete t;
    ctx->last_cuda = (int)e;
    return BPG_ERR_NOMEM;
  }
  *out = t;
  return BPG_OK;
}

// d_* pointers are device pointers; factors may be null (all ones).
// Either (G, H, Q_host) are given and the state builds its own windowed [G | H | Q] table,
// or `shared` is a windowed table that already holds the generators at g_base/h_base and a
// base point at q_id with Q = q_mul * shared[q_id] (the R1CS prover's Q = w*B).
static int ipp_begin_dev(bpg_ctx* ctx, const bpg_table* G, size_t g_off, const bpg_table* H, size_t h_off, size_t n,
                         const uint8_t* Q_host, const bpg_table* shared, size_t g_base, size_t h_base, size_t q_id,
                         const uint8_t* q_mul_host, const uint32_t* d_gf, const uint32_t* d_hf, const uint32_t* d_a,
                         const uint32_t* d_b, bpg_ipp** out) {
  if (n == 0 || (n & (n - 1))) return BPG_ERR_POW2;
  if (n >= (1u << 28)) return BPG_ERR_ARG;
  if (shared) {
    if (g_base + n > shared->n || h_base + n > shared->n || q_id >= shared->n) return BPG_ERR_CAPACITY;
    if (!shared->win_c && n > 1) return BPG_ERR_ARG;
  } else if (g_off + n > G->n || h_off + n > H->n) {
    return BPG_ERR_CAPACITY;
  }
  bpg_ipp* st = new (std::nothrow) bpg_ipp();
  if (!st) return BPG_ERR_NOMEM;
  memset(st, 0, sizeof *st);
  st->ctx = ctx;
  st->n = st->m = n;
  int rc = BPG_OK;
  size_t T = 2 * n + 2;
  size_t nparts = 256;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  size_t o_a = take(n * 32), o_b = take(n * 32), o_wG = take(n * 32), o_wH = take(n * 32);
  size_t o_sc = take(T * 32), o_pid = take(T * 4), o_set = take(T), o_part = take(nparts * 64);
  size_t o_u = take(64), o_ext = take(4 * 128), o_bytes = take(64), o_q = take(32), o_qm = take(32);
  size_t o_qside = take(64), o_qcomb = take((size_t)COMB_ENTRIES * 96);
  do {
    cudaError_t e = dev_alloc(ctx, &st->buf, off);
    if (e != cudaSuccess) { ctx->last_cuda = (int)e; rc = BPG_ERR_NOMEM; break; }
    st->a = (uint32_t*)(st->buf + o_a); st->b = (uint32_t*)(st->buf + o_b);
    st->wG = (uint32_t*)(st->buf + o_wG); st->wH = (uint32_t*)(st->buf + o_wH);
    st->scalars = (uint32_t*)(st->buf + o_sc); st->point_ids = (uint32_t*)(st->buf + o_pid);
    st->set_ids = st->buf + o_set; st->partials = (uint32_t*)(st->buf + o_part);
    st->u_pair = (uint32_t*)(st->buf + o_u); st->out_ext = (uint32_t*)(st->buf + o_ext);
    st->out_bytes = st->buf + o_bytes;
    st->q_mul = (uint32_t*)(st->buf + o_qm);
    st->q_side = (uint32_t*)(st->buf + o_qside);
    st->q_comb = (uint32_t*)(st->buf + o_qcomb);
    uint8_t* d_q = st->buf + o_q;
    cudaStream_t s = ctx->stream;
    if (cudaMemcpyAsync(st->a, d_a, n * 32, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(st->b, d_b, n * 32, cudaMemcpyDeviceToDevice, s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
    unsigned gn = (unsigned)((n + 255) / 256);
    k_ipp_init_weights<<<gn, 256, 0, s>>>(d_gf, d_hf, (uint32_t)n, st->wG, st->wH);
    ctx->launches++;
    if (shared) {
      st->tab = shared;
      st->has_qmul = q_mul_host != nullptr;
      if (q_mul_host) {
        memcpy(ctx->h_pinned + 512, q_mul_host, 32);
        if (cudaMemcpyAsync(st->q_mul, ctx->h_pinned + 512, 32, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
      }
      k_ipp_point_ids<<<gn, 256, 0, s>>>(st->point_ids, (uint32_t)n, (uint32_t)g_base, (uint32_t)h_base, (uint32_t)q_id);
      ctx->launches++;
      if (cudaStreamSynchronize(s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
    } else if (G == H && G->win_c && n > 1) {
      // The caller's generators already live in ONE windowed table (a resident BulletproofGens):
      // run the round MSMs over it as they are and form c_L Q, c_R Q from a fixed-base comb of Q
      // built here once, on the auxiliary stream beside each round's MSM.
      st->tab = G;
      st->q_sep = true;
      k_ipp_point_ids<<<gn, 256, 0, s>>>(st->point_ids, (uint32_t)n, (uint32_t)g_off, (uint32_t)h_off, (uint32_t)g_off);
      ctx->launches++;
      memcpy(ctx->h_pinned + 512, Q_host, 32);
      uint32_t* bad = reinterpret_cast<uint32_t*>(ctx->d_small);
      if (cudaMemsetAsync(bad, 0, 4, s) != cudaSuccess ||
          cudaMemcpyAsync(d_q, ctx->h_pinned + 512, 32, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
      k_comb_build<<<1, COMB_WINDOWS, 0, s>>>(d_q, st->q_comb, bad);
      ctx->launches++;
      uint32_t* hbad = reinterpret_cast<uint32_t*>(ctx->h_pinned);
      if (cudaMemcpyAsync(hbad, bad, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
          cudaStreamSynchronize(s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
      if (*hbad) { rc = BPG_ERR_DECODE; break; }
    } else {
      k_ipp_point_ids<<<gn, 256, 0, s>>>(st->point_ids, (uint32_t)n, 0u, (uint32_t)n, (uint32_t)(2 * n));
      ctx->launches++;
      // combined table [G | H | Q]
      rc = table_alloc_plain(ctx, 2 * n + 1, &st->own_tab);
      if (rc) break;
      st->tab = st->own_tab;
      if (cudaMemcpyAsync(st->own_tab->niels, G->niels + g_off * 24, n * 96, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
          cudaMemcpyAsync(st->own_tab->niels + n * 24, H->niels + h_off * 24, n * 96, cudaMemcpyDeviceToDevice, s) !=
              cudaSuccess) { rc = BPG_ERR_CUDA; break; }
      memcpy(ctx->h_pinned + 512, Q_host, 32);
      uint32_t* bad = reinterpret_cast<uint32_t*>(ctx->d_small);
      if (cudaMemsetAsync(bad, 0, 4, s) != cudaSuccess ||
          cudaMemcpyAsync(d_q, ctx->h_pinned + 512, 32, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
      k_decode_to_niels<<<1, 128, 0, s>>>(d_q, 1, st->own_tab->niels + 2 * n * 24, bad);
      ctx->launches++;
      uint32_t* hbad = reinterpret_cast<uint32_t*>(ctx->h_pinned);
      if (cudaMemcpyAsync(hbad, bad, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
          cudaStreamSynchronize(s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
      if (*hbad) { rc = BPG_ERR_DECODE; break; }
      if (n > 1) {
        rc = bpg_table_set_windows(ctx, st->own_tab, pick_window(n + 1, ctx->forced_c));
        if (rc) break;
      }
    }
  } while (0);
  if (rc != BPG_OK) {
    if (st->own_tab) bpg_table_free(st->own_tab);
    dev_free(ctx, st->buf);
    delete st;
    return rc;
  }
  *out = st;
  return BPG_OK;
}

extern "C" int bpg_ipp_begin(bpg_ctx* ctx, const bpg_table* G, size_t g_off, const bpg_table* H, size_t h_off,
                             size_t n, const uint8_t Q[32], const uint8_t* G_factors, const uint8_t* H_factors,
                             const uint8_t* a, const uint8_t* b, bpg_ipp** out) {
  if (!ctx || !G || !H || !Q || !a || !b || !out) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_stage(ctx, 4 * n * 32 + 64);
  if (rc) return rc;
  uint8_t* d = ctx->d_stage;
  CK(cudaMemcpyAsync(d, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d + n * 32, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (G_factors) CK(cudaMemcpyAsync(d + 2 * n * 32, G_factors, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (H_factors) CK(cudaMemcpyAsync(d + 3 * n * 32, H_factors, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  return ipp_begin_dev(ctx, G, g_off, H, h_off, n, Q, nullptr, 0, 0, 0, nullptr,
                       G_factors ? (const uint32_t*)(d + 2 * n * 32) : nullptr,
                       H_factors ? (const uint32_t*)(d + 3 * n * 32) : nullptr, (const uint32_t*)d,
                       (const uint32_t*)(d + n * 32), out);
}

extern "C" int bpg_ipp_begin_dev(bpg_ctx* ctx, const bpg_table* G, size_t g_off, const bpg_table* H, size_t h_off,
                                 size_t n, const uint8_t Q[32], const void* d_G_factors, const void* d_H_factors,
                                 const void* d_a, const void* d_b, bpg_ipp** out) {
  if (!ctx || !G || !H || !Q || !d_a || !d_b || !out) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  return ipp_begin_dev(ctx, G, g_off, H, h_off, n, Q, nullptr, 0, 0, 0, nullptr, (const uint32_t*)d_G_factors,
                       (const uint32_t*)d_H_factors, (const uint32_t*)d_a, (const uint32_t*)d_b, out);
}

// Generators and the base of Q live in one windowed table (the R1CS prover: G at g_base, H at
// h_base, Q = q_mul * shared[q_id] with q_id the Pedersen base B and q_mul the challenge w).
extern "C" int bpg_ipp_begin_shared(bpg_ctx* ctx, const bpg_table* shared, size_t g_base, size_t h_base, size_t q_id,
                                    const uint8_t q_mul[32], size_t n, const uint8_t* G_factors,
                                    const uint8_t* H_factors, const uint8_t* a, const uint8_t* b, bpg_ipp** out) {
  if (!ctx || !shared || !a || !b || !out) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_stage(ctx, 4 * n * 32 + 64);
  if (rc) return rc;
  uint8_t* d = ctx->d_stage;
  CK(cudaMemcpyAsync(d, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d + n * 32, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (G_factors) CK(cudaMemcpyAsync(d + 2 * n * 32, G_factors, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (H_factors) CK(cudaMemcpyAsync(d + 3 * n * 32, H_factors, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  return ipp_begin_dev(ctx, nullptr, 0, nullptr, 0, n, nullptr, shared, g_base, h_base, q_id, q_mul,
                       G_factors ? (const uint32_t*)(d + 2 * n * 32) : nullptr,
                       H_factors ? (const uint32_t*)(d + 3 * n * 32) : nullptr, (const uint32_t*)d,
                       (const uint32_t*)(d + n * 32), out);
}

extern "C" size_t bpg_ipp_rounds_left(const bpg_ipp* st) {
  size_t r = 0;
  if (!st) return 0;
  for (size_t m = st->m; m > 1; m >>= 1) r++;
  return r;
}

extern "C" int bpg_ipp_round_LR(bpg_ipp* st, uint8_t L[32], uint8_t R[32]) {
  if (!st || !L || !R) return BPG_ERR_ARG;
  if (st->m <= 1 || st->lr_done) return BPG_ERR_ARG;
  bpg_ctx* ctx = st->ctx;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  size_t n = st->n, m = st->m, h = m / 2;
  unsigned gcross = (unsigned)std::min<size_t>(256, (h + IPP_THREADS - 1) / IPP_THREADS);
  prof_mark(ctx, BPG_PROF_OTHER);
  k_ipp_cross<<<gcross, IPP_THREADS, 0, s>>>(st->a, st->b, (uint32_t)h, st->partials);
  LAUNCH_CHECK();
  k_ipp_round_scalars<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(st->a, st->b, st->wG, st->wH, (uint32_t)n,
                                                                  (uint32_t)m, st->scalars, st->set_ids);
  LAUNCH_CHECK();
  k_ipp_cross_finish<<<1, IPP_THREADS, 0, s>>>(st->partials, gcross, (uint32_t)n, st->has_qmul ? st->q_mul : nullptr,
                                               st->scalars, st->set_ids, st->q_sep ? st->q_side : nullptr);
  LAUNCH_CHECK();
  if (st->q_sep) {
    // c_L Q, c_R Q: 64 mixed additions each from the comb of Q, beside the MSM
    CK(cudaEventRecord(ctx->ev_fork, s));
    CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    k_comb_mul<<<1, 128, 0, ctx->aux_stream>>>(st->q_comb, 1, st->q_side, 2u, bias_for(4), nullptr, st->out_ext + 64);
    LAUNCH_CHECK();
    CK(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
  }
  int rc = msm_enqueue(ctx, st->tab->niels, st->tab->n, st->scalars, 2 * n + 2, st->set_ids, st->point_ids, 2,
                       st->out_ext, st->tab->win_c, st->tab->n);
  if (rc) return rc;
  if (st->q_sep) CK(cudaStreamWaitEvent(s, ctx->ev_join, 0));
  rc = bpg_dev_sum_encode(ctx, st->out_ext, st->q_sep ? 2 : 1, 2, st->out_bytes, nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->h_pinned, st->out_bytes, 64, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  memcpy(L, ctx->h_pinned, 32);
  memcpy(R, ctx->h_pinned + 32, 32);
  st->lr_done = true;
  return BPG_OK;
}

extern "C" int bpg_ipp_round_fold(bpg_ipp* st, const uint8_t u[32], const uint8_t u_inv[32]) {
  if (!st || !u || !u_inv) return BPG_ERR_ARG;
  if (st->m <= 1 || !st->lr_done) return BPG_ERR_ARG;
  bpg_ctx* ctx = st->ctx;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  // the challenge pair travels as kernel arguments: no staging copy, and no wait here -- the
  // next round's launches queue up behind the fold
  ScPair up;
  memcpy(up.v, u, 32);
  memcpy(up.v + 8, u_inv, 32);
  prof_mark(ctx, BPG_PROF_OTHER);
  k_ipp_fold<<<(unsigned)((st->n + 255) / 256), 256, 0, s>>>(st->a, st->b, st->wG, st->wH, (uint32_t)st->n,
                                                             (uint32_t)st->m, up);
  LAUNCH_CHECK();
  prof_mark(ctx, -1);
  st->m /= 2;
  st->lr_done = false;
  return BPG_OK;
}

extern "C" int bpg_ipp_finish(bpg_ipp* st, uint8_t a[32], uint8_t b[32]) {
  if (!st || !a || !b) return BPG_ERR_ARG;
  if (st->m != 1) return BPG_ERR_ARG;
  bpg_ctx* ctx = st->ctx;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(ctx->h_pinned, st->a, 32, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_pinned + 32, st->b, 32, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  memcpy(a, ctx->h_pinned, 32);
  memcpy(b, ctx->h_pinned + 32, 32);
  return BPG_OK;
}

extern "C" void bpg_ipp_free(bpg_ipp* st) {
  if (!st) return;
  cudaSetDevice(st->ctx->device);
  if (st->q_sep) cudaStreamSynchronize(st->ctx->aux_stream);
  if (st->own_tab) bpg_table_free(st->own_tab);
  dev_free(st->ctx, st->buf);
  delete st;
}

// ---------------------------------------------------------------------------
// indexed MSM: term t = scalars[t] * table[point_ids[t]] accumulated into out[set_ids[t]]
// ---------------------------------------------------------------------------
extern "C" int bpg_msm_table_indexed(bpg_ctx* ctx, const bpg_table* table, const uint32_t* point_ids,
                                     const uint8_t* set_ids, const uint8_t* scalars_le, size_t n_terms, int n_sets,
                                     uint8_t* out) {
  if (!ctx || !table || !out || n_sets <= 0 || n_sets > 255 || (n_terms && (!point_ids || !scalars_le))) return BPG_ERR_ARG;
  if ((size_t)n_sets * 160 > SMALL_BYTES) return BPG_ERR_ARG;
  for (size_t t = 0; t < n_terms; t++) {
    if (point_ids[t] >= table->n) return BPG_ERR_CAPACITY;
    if (set_ids && set_ids[t] >= n_sets) return BPG_ERR_ARG;
  }
  CK(cudaSetDevice(ctx->device));
  size_t o_pid = align_up(n_terms * 32), o_set = o_pid + align_up(n_terms * 4);
  int rc = ensure_stage(ctx, o_set + align_up(n_terms) + 64);
  if (rc) return rc;
  uint8_t* d = ctx->d_stage;
  cudaStream_t s = ctx->stream;
  if (n_terms) {
    CK(cudaMemcpyAsync(d, scalars_le, n_terms * 32, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d + o_pid, point_ids, n_terms * 4, cudaMemcpyHostToDevice, s));
    if (set_ids) CK(cudaMemcpyAsync(d + o_set, set_ids, n_terms, cudaMemcpyHostToDevice, s));
    else CK(cudaMemsetAsync(d + o_set, 0, n_terms, s));
  }
  uint32_t* d_ext = (uint32_t*)ctx->d_small;
  uint8_t* d_bytes = ctx->d_small + (size_t)n_sets * 128;
  rc = msm_enqueue(ctx, table->niels, table->n, (const uint32_t*)d, n_terms, d + o_set, (const uint32_t*)(d + o_pid),
                   n_sets, d_ext, table->win_c, table->n);
  if (rc) return rc;
  rc = bpg_dev_sum_encode(ctx, d_ext, 1, n_sets, d_bytes, nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->h_pinned, d_bytes, (size_t)n_sets * 32, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  memcpy(out, ctx->h_pinned, (size_t)n_sets * 32);
  return BPG_OK;
}

// ---------------------------------------------------------------------------
// one MSM over ad-hoc (compressed) points followed by ranges of resident tables
// ---------------------------------------------------------------------------
// core of the mixed MSM: scalars for all `total` terms are already in d_scalars (device)
static int msm_mixed_core(bpg_ctx* ctx, const uint8_t* d_adhoc_points, size_t n_adhoc, const bpg_table* const* tabs,
                          const size_t* offs, const size_t* lens, int nsegs, const uint32_t* d_scalars, size_t total,
                          uint8_t out[32], bool identity_only = false /*out: zeros iff the sum is the identity*/) {
  cudaStream_t s = ctx->stream;
  uint32_t* bad = reinterpret_cast<uint32_t*>(ctx->d_small + 1024);
  uint32_t* d_ext = (uint32_t*)ctx->d_small;  // up to two partial sums
  uint8_t* d_bytes = ctx->d_small + 256;
  CK(cudaMemsetAsync(bad, 0, 4, s));
  // Fast path: every range lies in ONE windowed table (the R1CS verifier: B, B_blinding, G, H of the
  // generator table).  The table terms run as an indexed MSM over the window multiples (no doublings)
  // on the launch stream while the few ad-hoc points (proof points: decoded, no precomputation) run
  // as a small plain MSM on the auxiliary stream; the two partial sums are added at the end.
  bool one_windowed = nsegs >= 1 && nsegs <= 4 && tabs[0]->win_c != 0;
  for (int i = 1; i < nsegs && one_windowed; i++) one_windowed = tabs[i] == tabs[0];
  size_t n_tab = total - n_adhoc;
  if (one_windowed && n_tab > 0) {
    const bpg_table* T = tabs[0];
    bpg_table* ta = nullptr;
    uint32_t* d_ids = nullptr;
    int rc = BPG_OK;
    int n_parts = 1;
    do {
      if (dev_alloc(ctx, &d_ids, n_tab * 4) != cudaSuccess) { rc = BPG_ERR_NOMEM; break; }
      if (n_adhoc) {
        rc = table_alloc_plain(ctx, n_adhoc, &ta);
        if (rc) break;
        rc = BPG_ERR_CUDA;
        if (cudaEventRecord(ctx->ev_fork, s) != cudaSuccess) break;
        if (cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0) != cudaSuccess) break;
        k_decode_to_niels<<<(unsigned)((n_adhoc + 127) / 128), 128, 0, ctx->aux_stream>>>(d_adhoc_points, (uint32_t)n_adhoc,
                                                                                          ta->niels, bad);
        ctx->launches++;
        rc = msm_enqueue(ctx, ta->niels, n_adhoc, d_scalars, n_adhoc, nullptr, nullptr, 1, d_ext + 32, 0, 0, /*lane=*/1);
        if (rc) break;
        rc = BPG_ERR_CUDA;
        if (cudaEventRecord(ctx->ev_join, ctx->aux_stream) != cudaSuccess) break;
        n_parts = 2;
      }
      SegIds sg;
      sg.n = nsegs;
      for (int i = 0; i < 4; i++) {
        sg.off[i] = i < nsegs ? (uint32_t)offs[i] : 0;
        sg.len[i] = i < nsegs ? (uint32_t)lens[i] : 0;
      }
      k_seg_point_ids<<<(unsigned)((n_tab + 255) / 256), 256, 0, s>>>(sg, (uint32_t)n_tab, d_ids);
      ctx->launches++;
      rc = msm_enqueue(ctx, T->niels, T->n, d_scalars + n_adhoc * 8, n_tab, nullptr, d_ids, 1, d_ext, T->win_c, T->n);
      if (rc) break;
      rc = BPG_ERR_CUDA;
      if (n_adhoc && cudaStreamWaitEvent(s, ctx->ev_join, 0) != cudaSuccess) break;
      if (identity_only) {
        prof_mark(ctx, BPG_PROF_ENCODE);
        k_sum_is_identity<<<1, 32, 0, s>>>(d_ext, n_parts, d_bytes);
        ctx->launches++;
        prof_mark(ctx, -1);
      } else {
        rc = bpg_dev_sum_encode(ctx, d_ext, n_parts, 1, d_bytes, nullptr);
        if (rc) break;
      }
      rc = BPG_ERR_CUDA;
      if (cudaMemcpyAsync(ctx->h_pinned, d_bytes, 32, cudaMemcpyDeviceToHost, s) != cudaSuccess) break;
      if (cudaMemcpyAsync(ctx->h_pinned + 64, bad, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess) break;
      cudaError_t se = cudaStreamSynchronize(s);
      if (se != cudaSuccess) { ctx->last_cuda = (int)se; break; }
      if (*reinterpret_cast<uint32_t*>(ctx->h_pinned + 64)) { rc = BPG_ERR_DECODE; break; }
      memcpy(out, ctx->h_pinned, 32);
      rc = BPG_OK;
    } while (0);
    if (rc != BPG_OK && n_parts == 2) cudaStreamSynchronize(ctx->aux_stream);  // do not free under the aux lane
    if (ta) bpg_table_free(ta);
    dev_free(ctx, d_ids);
    return rc;
  }
  // General path: one plain table [adhoc | range copies], one MSM with Horner.
  bpg_table* t = nullptr;
  int rc = table_alloc_plain(ctx, total, &t);
  if (rc) return rc;
  do {
    rc = BPG_ERR_CUDA;
    if (n_adhoc) {
      k_decode_to_niels<<<(unsigned)((n_adhoc + 127) / 128), 128, 0, s>>>(d_adhoc_points, (uint32_t)n_adhoc, t->niels, bad);
      ctx->launches++;
    }
    size_t pos = n_adhoc;
    bool ok = true;
    for (int i = 0; i < nsegs && ok; i++) {
      if (lens[i] && cudaMemcpyAsync(t->niels + pos * 24, tabs[i]->niels + offs[i] * 24, lens[i] * 96,
                                     cudaMemcpyDeviceToDevice, s) != cudaSuccess)
        ok = false;
      pos += lens[i];
    }
    if (!ok) break;
    rc = msm_enqueue(ctx, t->niels, total, d_scalars, total, nullptr, nullptr, 1, d_ext);
    if (rc) break;
    rc = bpg_dev_sum_encode(ctx, d_ext, 1, 1, d_bytes, nullptr);
    if (rc) break;
    rc = BPG_ERR_CUDA;
    if (cudaMemcpyAsync(ctx->h_pinned, d_bytes, 32, cudaMemcpyDeviceToHost, s) != cudaSuccess) break;
    if (cudaMemcpyAsync(ctx->h_pinned + 64, bad, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess) break;
    cudaError_t se = cudaStreamSynchronize(s);
    if (se != cudaSuccess) { ctx->last_cuda = (int)se; break; }
    if (*reinterpret_cast<uint32_t*>(ctx->h_pinned + 64)) { rc = BPG_ERR_DECODE; break; }
    memcpy(out, ctx->h_pinned, 32);
    rc = BPG_OK;
  } while (0);
  bpg_table_free(t);
  return rc;
}

static int check_segs(const bpg_table* const* tabs, const size_t* offs, const size_t* lens, int nsegs, size_t* total) {
  for (int i = 0; i < nsegs; i++) {
    if (!tabs[i]) return BPG_ERR_ARG;
    if (offs[i] + lens[i] > tabs[i]->n) return BPG_ERR_CAPACITY;
    *total += lens[i];
  }
  return BPG_OK;
}

extern "C" int bpg_msm_mixed(bpg_ctx* ctx, const uint8_t* adhoc_points, size_t n_adhoc,
                             const bpg_table* const* tabs, const size_t* offs, const size_t* lens, int nsegs,
                             const uint8_t* scalars_le, uint8_t out[32]) {
  if (!ctx || !out || (n_adhoc && !adhoc_points) || nsegs < 0 || (nsegs && (!tabs || !offs || !lens))) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  size_t total = n_adhoc;
  int rc = check_segs(tabs, offs, lens, nsegs, &total);
  if (rc) return rc;
  if (total && !scalars_le) return BPG_ERR_ARG;
  rc = ensure_stage(ctx, std::max<size_t>(total * 32 + n_adhoc * 32, 64));
  if (rc) return rc;
  uint8_t* d_sc = ctx->d_stage;
  uint8_t* d_pts = ctx->d_stage + total * 32;
  if (total) CK(cudaMemcpyAsync(d_sc, scalars_le, total * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (n_adhoc) CK(cudaMemcpyAsync(d_pts, adhoc_points, n_adhoc * 32, cudaMemcpyHostToDevice, ctx->stream));
  return msm_mixed_core(ctx, d_pts, n_adhoc, tabs, offs, lens, nsegs, (const uint32_t*)d_sc, total, out);
}

#include "r1cs_dev.inc"
#include "stark_msm.inc"
#include "stark_ipp.inc"
