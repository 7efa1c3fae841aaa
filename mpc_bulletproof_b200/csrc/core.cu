// libbpgpu: context, memory, tables, fixed-base comb, peer exchange and the host-buffer MSM entry points
// of the C ABI (include/bpgpu.h).  The Pippenger launch itself is msm.cu.
#define BPG_FE_OUTLINE 1  // latency-bound kernels: products are calls, not 1.5 KB of inline code each
#include "internal.cuh"
#include "point_kernels.cuh"

using namespace bpg;

// Transient device memory (tables, proof states).  Freed blocks are parked in a small per-context
// cache and handed out again (best fit within 4x) before falling back to the device's stream-ordered
// pool: a proof allocates its state per call, and neither a cudaMalloc nor a pool miss (both cost
// milliseconds at these sizes) may sit on that path.  Everything a context allocates is used on its
// launch stream (the auxiliary lane is fenced by events), so reuse is ordered by the stream.
static constexpr size_t CACHE_SLOTS = 24;
cudaError_t dev_alloc(bpg_ctx* ctx, void** p, size_t bytes) {
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
void dev_free(bpg_ctx* ctx, void* p) {
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
  if (e == cudaSuccess) e = msm_kernels_init();
  if (e == cudaSuccess) e = msm_sort_kernels_init();
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

void pipe_release(bpg_ctx* ctx);
extern "C" void bpg_free(bpg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  pipe_release(ctx);
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
  if (ctx->d_terms) cudaFree(ctx->d_terms);
  if (ctx->d_adhoc) cudaFree(ctx->d_adhoc);
  if (ctx->q_cache_comb) cudaFree(ctx->q_cache_comb);
  if (ctx->ev_terms) cudaEventDestroy(ctx->ev_terms);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

extern "C" int bpg_set_stream(bpg_ctx* ctx, void* s, int use_own) {
  if (!ctx) return BPG_ERR_ARG;
  cudaStream_t next = use_own ? ctx->own_stream : (cudaStream_t)s;
  if (next != ctx->stream) {
    // the arenas, the staging buffers and the block cache are ordered by the launch stream: work already
    // enqueued on the old stream must finish before anything on the new one may reuse them
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
    CK(cudaStreamWaitEvent(next, ctx->ev_fork, 0));
  }
  ctx->stream = next;
  return BPG_OK;
}

// ---- per-phase profiling ---------------------------------------------------
void prof_mark(bpg_ctx* ctx, int phase) {
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

int ensure_ws(bpg_ctx* ctx, size_t bytes, int lane) {
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
int ensure_stage(bpg_ctx* ctx, size_t bytes) {
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
  cudaError_t e = dev_alloc(ctx, &t->niels, std::max<size_t>(n, 1) * NIELS_BYTES);
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
extern "C" size_t bpg_table_entry_bytes(const bpg_table* t) { return t ? NIELS_BYTES : 0; }
extern "C" void bpg_table_free(bpg_table* t) {
  if (!t) return;
  cudaSetDevice(t->ctx->device);
  dev_free(t->ctx, t->niels);  // stream-ordered: work already enqueued on the table finishes first
  if (t->comb) {
    cudaStreamSynchronize(t->ctx->stream);
    cudaFree(t->comb);
  }
  delete t;
}
extern "C" int bpg_dev_msm_table(bpg_ctx* ctx, const bpg_table* table, size_t offset, size_t n,
                                 const void* d_scalars, int n_sets, void* d_out_ext) {
  if (!ctx || !table || !d_out_ext || (!d_scalars && n) || n_sets <= 0) return BPG_ERR_ARG;
  if (offset + n > table->n) return BPG_ERR_CAPACITY;
  CK(cudaSetDevice(ctx->device));
  return msm_enqueue(ctx, table->niels + offset * NIELS_WORDS, n, (const uint32_t*)d_scalars, n * (size_t)n_sets, nullptr,
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
  uint32_t* status = reinterpret_cast<uint32_t*>(p->local + p->parts_bytes + align_up((size_t)2 * p->world * 4));
  prof_mark(ctx, BPG_PROF_ENCODE);
  k_exchange_sum_encode<<<1, XCH_THREADS, 0, ctx->stream>>>((const uint32_t*)d_part, p->ptrs, p->world, p->rank, n_sets,
                                                            p->max_sets, p->seq + 1, (uint8_t*)d_out_bytes,
                                                            (uint32_t*)d_out_ext, status);
  LAUNCH_CHECK();
  p->seq++;  // only a launch that went out advances the step counter the ranks share
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
  rc = msm_enqueue(ctx, table->niels + offset * NIELS_WORDS, n, (const uint32_t*)ctx->d_stage, n * (size_t)n_sets, nullptr,
                   nullptr, n_sets, d_ext, table->win_c, table->n, 0, 0, sbytes ? scalars_le : nullptr);
  if (rc) return rc;
  rc = bpg_dev_sum_encode(ctx, d_ext, 1, n_sets, d_bytes, nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->h_pinned, d_bytes, (size_t)n_sets * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->h_pinned, (size_t)n_sets * 32);
  return BPG_OK;
}

// ---------------------------------------------------------------------------
// Pipelined host-buffer MSMs over a resident table: submit returns once the work is enqueued, wait returns the
// result.  Two jobs may be in flight per context: the scalars of job i + 1 travel host -> device on the copy
// stream into the second staging buffer while job i computes, so a stream of MSMs runs at the rate of the
// slower of the two (the 2^20-point step: 32 MiB of upload under 1.6 ms of kernels) instead of their sum.
// ---------------------------------------------------------------------------
struct bpg_msm_job {
  bpg_ctx* ctx;
  int slot;
  int n_sets;
  bool waited;
};
struct MsmPipe {
  uint8_t* d_sc[2] = {nullptr, nullptr};
  size_t cap[2] = {0, 0};
  uint8_t* d_res[2] = {nullptr, nullptr};   // per slot: ext sums | encodings
  uint8_t* h_res[2] = {nullptr, nullptr};   // pinned
  cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  bool busy[2] = {false, false};
  int next = 0;
  cudaStream_t copy = nullptr;
};
static std::unordered_map<bpg_ctx*, MsmPipe*>& pipes() {
  static auto* m = new std::unordered_map<bpg_ctx*, MsmPipe*>();
  return *m;
}
static std::mutex& pipes_mu() {
  static std::mutex* m = new std::mutex();
  return *m;
}
static constexpr size_t PIPE_MAX_SETS = 64;
static int pipe_get(bpg_ctx* ctx, MsmPipe** out) {
  std::lock_guard<std::mutex> lk(pipes_mu());
  auto it = pipes().find(ctx);
  if (it != pipes().end()) {
    *out = it->second;
    return BPG_OK;
  }
  MsmPipe* p = new (std::nothrow) MsmPipe();
  if (!p) return BPG_ERR_NOMEM;
  cudaError_t e = cudaStreamCreateWithFlags(&p->copy, cudaStreamNonBlocking);
  for (int k = 0; k < 2 && e == cudaSuccess; k++) {
    e = cudaEventCreateWithFlags(&p->ev_up[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_done[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_res[k], PIPE_MAX_SETS * 160);
    if (e == cudaSuccess) e = cudaMallocHost(&p->h_res[k], PIPE_MAX_SETS * 32);
  }
  if (e != cudaSuccess) {
    ctx->last_cuda = (int)e;
    delete p;
    return BPG_ERR_CUDA;
  }
  pipes()[ctx] = p;
  *out = p;
  return BPG_OK;
}
void pipe_release(bpg_ctx* ctx) {  // bpg_free
  std::lock_guard<std::mutex> lk(pipes_mu());
  auto it = pipes().find(ctx);
  if (it == pipes().end()) return;
  MsmPipe* p = it->second;
  cudaStreamSynchronize(p->copy);
  for (int k = 0; k < 2; k++) {
    if (p->d_sc[k]) cudaFree(p->d_sc[k]);
    if (p->d_res[k]) cudaFree(p->d_res[k]);
    if (p->h_res[k]) cudaFreeHost(p->h_res[k]);
    if (p->ev_up[k]) cudaEventDestroy(p->ev_up[k]);
    if (p->ev_done[k]) cudaEventDestroy(p->ev_done[k]);
  }
  cudaStreamDestroy(p->copy);
  delete p;
  pipes().erase(it);
}

extern "C" int bpg_msm_table_submit(bpg_ctx* ctx, const bpg_table* table, size_t offset, size_t n, const uint8_t* scalars_le,
                                    int n_sets, bpg_msm_job** out) {
  if (!ctx || !table || !out || (!scalars_le && n) || n_sets <= 0 || (size_t)n_sets > PIPE_MAX_SETS) return BPG_ERR_ARG;
  if (offset + n > table->n) return BPG_ERR_CAPACITY;
  CK(cudaSetDevice(ctx->device));
  MsmPipe* p = nullptr;
  int rc = pipe_get(ctx, &p);
  if (rc) return rc;
  const int k = p->next;
  if (p->busy[k]) return BPG_ERR_ARG;  // both slots in flight: wait for the older job first
  const size_t sbytes = std::max<size_t>(n * (size_t)n_sets * 32, 32);
  if (sbytes > p->cap[k]) {
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(p->copy));
    if (p->d_sc[k]) cudaFree(p->d_sc[k]);
    p->d_sc[k] = nullptr;
    p->cap[k] = 0;
    cudaError_t e = cudaMalloc(&p->d_sc[k], sbytes);
    if (e != cudaSuccess) {
      ctx->last_cuda = (int)e;
      return BPG_ERR_NOMEM;
    }
    p->cap[k] = sbytes;
  }
  // the staging buffer of this slot was last read by the job that used it two submissions ago: its ev_done
  // (recorded on the launch stream) orders the upload behind it
  CK(cudaStreamWaitEvent(p->copy, p->ev_done[k], 0));
  if (n) CK(cudaMemcpyAsync(p->d_sc[k], scalars_le, n * (size_t)n_sets * 32, cudaMemcpyHostToDevice, p->copy));
  CK(cudaEventRecord(p->ev_up[k], p->copy));
  CK(cudaStreamWaitEvent(ctx->stream, p->ev_up[k], 0));
  uint32_t* d_ext = (uint32_t*)p->d_res[k];
  uint8_t* d_bytes = p->d_res[k] + (size_t)n_sets * 128;
  rc = msm_enqueue(ctx, table->niels + offset * NIELS_WORDS, n, (const uint32_t*)p->d_sc[k], n * (size_t)n_sets, nullptr, nullptr,
                   n_sets, d_ext, table->win_c, table->n);
  if (rc) return rc;
  rc = bpg_dev_sum_encode(ctx, d_ext, 1, n_sets, d_bytes, nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(p->h_res[k], d_bytes, (size_t)n_sets * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaEventRecord(p->ev_done[k], ctx->stream));
  bpg_msm_job* job = new (std::nothrow) bpg_msm_job();
  if (!job) return BPG_ERR_NOMEM;
  job->ctx = ctx;
  job->slot = k;
  job->n_sets = n_sets;
  job->waited = false;
  p->busy[k] = true;
  p->next = k ^ 1;
  *out = job;
  return BPG_OK;
}
// out: n_sets x 32 bytes.  Frees the job.
extern "C" int bpg_msm_job_wait(bpg_msm_job* job, uint8_t* out) {
  if (!job || !out) return BPG_ERR_ARG;
  bpg_ctx* ctx = job->ctx;
  MsmPipe* p = nullptr;
  int rc = pipe_get(ctx, &p);
  if (rc) return rc;
  cudaError_t e = cudaEventSynchronize(p->ev_done[job->slot]);
  p->busy[job->slot] = false;
  if (e == cudaSuccess) memcpy(out, p->h_res[job->slot], (size_t)job->n_sets * 32);
  delete job;
  if (e != cudaSuccess) {
    ctx->last_cuda = (int)e;
    return BPG_ERR_CUDA;
  }
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
  rc = msm_enqueue(ctx, table->niels + offset * NIELS_WORDS, n, (const uint32_t*)ctx->d_stage, n * (size_t)n_sets, nullptr,
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

// out[s] = encode(sum_p decode(points[p][s])), points laid out [part][set], n_parts <= 32: the open of
// additively shared points that travelled as compressed encodings (the MPC prover's party-to-party link)
extern "C" int bpg_points_sum(bpg_ctx* ctx, const uint8_t* points, int n_parts, int n_sets, uint8_t* out) {
  if (!ctx || !points || !out || n_parts <= 0 || n_parts > 32 || n_sets <= 0) return BPG_ERR_ARG;
  size_t in_bytes = (size_t)n_parts * n_sets * 32, out_bytes = (size_t)n_sets * 32;
  if (in_bytes + out_bytes + 256 > SMALL_BYTES) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  uint32_t* bad = reinterpret_cast<uint32_t*>(ctx->d_small);
  uint8_t* d_in = ctx->d_small + 256;
  uint8_t* d_out = d_in + in_bytes;
  memcpy(ctx->h_pinned + 256, points, in_bytes);
  CK(cudaMemsetAsync(bad, 0, 4, s));
  CK(cudaMemcpyAsync(d_in, ctx->h_pinned + 256, in_bytes, cudaMemcpyHostToDevice, s));
  k_points_sum<<<n_sets, ENC_THREADS, 0, s>>>(d_in, n_parts, n_sets, d_out, bad);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(ctx->h_pinned, bad, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(ctx->h_pinned + 256, d_out, out_bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (*reinterpret_cast<uint32_t*>(ctx->h_pinned)) return BPG_ERR_DECODE;
  memcpy(out, ctx->h_pinned + 256, out_bytes);
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
  cudaError_t e = dev_alloc(ctx, &out, (size_t)W * t->n * NIELS_BYTES);
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
int table_alloc_plain(bpg_ctx* ctx, size_t n, bpg_table** out) {
  bpg_table* t = new (std::nothrow) bpg_table();
  if (!t) return BPG_ERR_NOMEM;
  t->ctx = ctx;
  t->n = n;
  cudaError_t e = dev_alloc(ctx, &t->niels, std::max<size_t>(n, 1) * NIELS_BYTES);
  if (e != cudaSuccess) {
    delete t;
    ctx->last_cuda = (int)e;
    return BPG_ERR_NOMEM;
  }
  *out = t;
  return BPG_OK;
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
// ---------------------------------------------------------------------------
// Ad-hoc points ahead of their scalars.  A verifier knows every point of its final check as soon as it holds the
// proof (verifier.rs:516-547: A_*, S_*, V_*, T_*, L_*, R_*), and the doublings of a variable-base multiplication do
// not depend on the scalar: decoding, the 252-step doubling chain and the digit multiples run on the auxiliary
// stream while the transcript is replayed and the constraints are flattened.  The MSM that later names the same
// points (same encodings, same order) adds comb entries instead of running its own double-and-add.
// ---------------------------------------------------------------------------
extern "C" int bpg_adhoc_prefetch(bpg_ctx* ctx, const uint8_t* points, size_t n) {
  if (!ctx || (n && !points)) return BPG_ERR_ARG;
  ctx->adhoc_n = 0;
  if (n == 0 || n > ADHOC_MAX_POINTS) return BPG_OK;  // larger sets keep the plain path
  CK(cudaSetDevice(ctx->device));
  if (n > ctx->adhoc_cap) {
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->aux_stream));
    if (ctx->d_adhoc) cudaFree(ctx->d_adhoc);
    ctx->d_adhoc = nullptr;
    ctx->adhoc_cap = 0;
    const size_t cap = std::max<size_t>(64, std::min<size_t>(ADHOC_MAX_POINTS, 2 * n));
    CK(cudaMalloc(&ctx->d_adhoc, AdhocLayout(cap).total));
    ctx->adhoc_cap = cap;
  }
  const AdhocLayout L(ctx->adhoc_cap);
  uint8_t* d = ctx->d_adhoc;
  cudaStream_t a = ctx->aux_stream;
  // the buffer's only reader is the auxiliary stream itself (in order)
  CK(cudaMemcpyAsync(d + L.comp, points, n * 32, cudaMemcpyHostToDevice, a));
  CK(cudaMemsetAsync(d + L.flags, 0, 8, a));
  int rc = comb_from_points(ctx, a, d + L.comp, n, (uint32_t*)(d + L.ext), (uint32_t*)(d + L.chain), (uint32_t*)(d + L.comb),
                            (uint32_t*)(d + L.flags));
  if (rc) return rc;
  ctx->adhoc_src.assign(points, points + n * 32);
  ctx->adhoc_n = n;
  return BPG_OK;
}
bool adhoc_matches(const bpg_ctx* ctx, const uint8_t* host_points, size_t n) {
  return n && ctx->adhoc_n == n && ctx->adhoc_src.size() == n * 32 && memcmp(ctx->adhoc_src.data(), host_points, n * 32) == 0;
}

int msm_mixed_core(bpg_ctx* ctx, const uint8_t* d_adhoc_points, size_t n_adhoc, const bpg_table* const* tabs,
                          const size_t* offs, const size_t* lens, int nsegs, const uint32_t* d_scalars, size_t total,
                          uint8_t out[32], bool identity_only /*out: zeros iff the sum is the identity*/,
                          bool adhoc_resident /*combs of exactly these points are in ctx->d_adhoc*/) {
  cudaStream_t s = ctx->stream;
  uint32_t* bad = reinterpret_cast<uint32_t*>(ctx->d_small + 1024);
  uint32_t* d_ext = (uint32_t*)ctx->d_small;  // up to two partial sums
  uint8_t* d_bytes = ctx->d_small + 256;
  CK(cudaMemsetAsync(bad, 0, 8, s));  // [0] encodings rejected here, [1] encodings rejected by bpg_adhoc_prefetch
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
      if (n_adhoc && adhoc_resident) {
        const AdhocLayout L(ctx->adhoc_cap);
        uint8_t* da = ctx->d_adhoc;
        rc = BPG_ERR_CUDA;
        if (cudaEventRecord(ctx->ev_fork, s) != cudaSuccess) break;  // the scalars were uploaded on the launch stream
        if (cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0) != cudaSuccess) break;
        rc = launch_comb_msm(ctx, ctx->aux_stream, (const uint32_t*)(da + L.comb), d_scalars, n_adhoc, (uint32_t*)(da + L.parts),
                             (uint32_t*)(da + L.flags) + 1, d_ext + 32);
        if (rc) break;
        rc = BPG_ERR_CUDA;
        if (cudaMemcpyAsync(bad + 1, da + L.flags, 4, cudaMemcpyDeviceToDevice, ctx->aux_stream) != cudaSuccess) break;
        if (cudaEventRecord(ctx->ev_join, ctx->aux_stream) != cudaSuccess) break;
        n_parts = 2;
      } else if (n_adhoc) {
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
      // a few thousand table terms: 64 comb additions each in two launches instead of the bucket pipeline's dozen
      static const size_t comb_max = env_size("BPG_MIXED_COMB_MAX", 4100);
      if (T->comb && n_tab <= comb_max) {
        const uint32_t single[4] = {0, 0, 0, 0}, lo[4] = {1, 0, 0, 0}, hi[4] = {(uint32_t)n_tab, 0, 0, 0};
        rc = launch_comb_terms(ctx, s, T->comb, d_scalars + n_adhoc * 8, d_ids, single, lo, hi, 1, nullptr, d_ext);
      } else {
        rc = msm_enqueue(ctx, T->niels, T->n, d_scalars + n_adhoc * 8, n_tab, nullptr, d_ids, 1, d_ext, T->win_c, T->n);
      }
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
      if (cudaMemcpyAsync(ctx->h_pinned + 64, bad, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess) break;
      cudaError_t se = cudaStreamSynchronize(s);
      if (se != cudaSuccess) { ctx->last_cuda = (int)se; break; }
      if (reinterpret_cast<uint32_t*>(ctx->h_pinned + 64)[0] | reinterpret_cast<uint32_t*>(ctx->h_pinned + 64)[1]) { rc = BPG_ERR_DECODE; break; }
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
      if (lens[i] && cudaMemcpyAsync(t->niels + pos * NIELS_WORDS, tabs[i]->niels + offs[i] * NIELS_WORDS, lens[i] * NIELS_BYTES,
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
  return msm_mixed_core(ctx, d_pts, n_adhoc, tabs, offs, lens, nsegs, (const uint32_t*)d_sc, total, out, false,
                        adhoc_matches(ctx, adhoc_points, n_adhoc));
}

// ---------------------------------------------------------------------------
// launch helpers for other translation units (the kernels live here)
// ---------------------------------------------------------------------------
void launch_decode_to_niels(bpg_ctx* ctx, cudaStream_t s, const uint8_t* d_comp, size_t n, uint32_t* niels, uint32_t* bad) {
  k_decode_to_niels<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_comp, (uint32_t)n, niels, bad);
  ctx->launches++;
}
void launch_comb_build(bpg_ctx* ctx, cudaStream_t s, const uint8_t* d_base32, uint32_t* table, uint32_t* bad) {
  k_comb_build<<<1, COMB_WINDOWS, 0, s>>>(d_base32, table, bad);
  ctx->launches++;
}
void launch_comb_mul(bpg_ctx* ctx, cudaStream_t s, const uint32_t* tables, int nbases, const uint32_t* d_scalars, size_t n,
                     uint8_t* d_out_bytes, uint32_t* d_out_ext) {
  k_comb_mul<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(tables, nbases, d_scalars, (uint32_t)n, bias_for(4), d_out_bytes,
                                                        d_out_ext);
  ctx->launches++;
}
