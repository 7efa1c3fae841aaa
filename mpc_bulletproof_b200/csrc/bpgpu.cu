// libbpgpu: C ABI (include/bpgpu.h) over the sm_100a kernels.
#include "../../include/bpgpu.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "msm_kernels.cuh"

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
  int sm_count = 148;
  // per-phase device timing (bpg_profile_*): events are recorded on the launch stream
  bool prof = false;
  std::vector<cudaEvent_t> prof_ev;   // pool
  std::vector<int> prof_phase;        // phase id of interval [ev[i], ev[i+1])
  size_t prof_used = 0;
  double prof_ms[BPG_PROF_NPHASE] = {0};
  uint64_t prof_n[BPG_PROF_NPHASE] = {0};
  // workspace arena (grown on demand, reused across calls)
  uint8_t* ws = nullptr;
  size_t ws_cap = 0;
  // small staging buffers
  uint8_t* d_small = nullptr;   // device scratch for results (>= 64 KB)
  uint8_t* h_pinned = nullptr;  // pinned host scratch (>= 64 KB)
  // staging for host-buffer calls
  uint8_t* d_stage = nullptr;
  size_t d_stage_cap = 0;
};

struct bpg_table {
  bpg_ctx* ctx;
  uint32_t* niels;  // n * 24 words
  size_t n;
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
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_small, SMALL_BYTES);
  if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_pinned, SMALL_BYTES);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) {
    delete ctx;
    return BPG_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  const char* env = getenv("BPG_MSM_C");
  if (env) ctx->forced_c = atoi(env);
  *out = ctx;
  return BPG_OK;
}

extern "C" void bpg_free(bpg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->ws) cudaFree(ctx->ws);
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

static int ensure_ws(bpg_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_cap) return BPG_OK;
  // the arena may still be in use by enqueued work
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->ws) cudaFree(ctx->ws);
  ctx->ws = nullptr;
  ctx->ws_cap = 0;
  size_t want = bytes + bytes / 8;
  cudaError_t e = cudaMalloc(&ctx->ws, want);
  if (e != cudaSuccess) {
    ctx->last_cuda = (int)e;
    return e == cudaErrorMemoryAllocation ? BPG_ERR_NOMEM : BPG_ERR_CUDA;
  }
  ctx->ws_cap = want;
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
  cudaError_t e = cudaMalloc(&t->niels, std::max<size_t>(n, 1) * 96);
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
    cudaFree(t->niels);
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
  cudaStreamSynchronize(t->ctx->stream);
  cudaFree(t->niels);
  delete t;
}

// ---------------------------------------------------------------------------
// MSM launch
// ---------------------------------------------------------------------------
static int pick_window(size_t n_per_set, int forced) {
  if (forced >= 2) return forced;
  double best = 1e300;
  int best_c = 4;
  for (int c = 3; c <= 20; c++) {
    int W = (255 + c - 1) / c;
    double nb = (double)(1u << (c - 1));
    // mixed adds (7M) for the terms, two full adds (9M) per bucket in the reduction,
    // c doublings per window on the serial tail (charged as if 64 lanes idle)
    double cost = W * ((double)n_per_set * 7.0 + nb * 18.0);
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

static void make_cfg(MsmCfg& cfg, size_t n_terms, size_t n_points, int nsets, int c) {
  cfg.c = c;
  cfg.W = (255 + c - 1) / c;
  cfg.nb = 1u << (c - 1);
  cfg.nsets = nsets;
  cfg.n_terms = (uint32_t)n_terms;
  cfg.n_points = (uint32_t)std::max<size_t>(n_points, 1);
  cfg.nwin = (uint32_t)nsets * cfg.W;
  cfg.B = cfg.nwin * cfg.nb;
  cfg.chunk = std::min<uint32_t>(cfg.nb, 32);
  cfg.nchunks = cfg.nb / cfg.chunk;
  double avg = (double)n_terms * cfg.W / (double)cfg.B;
  cfg.big_thresh = (uint32_t)std::max(256.0, 16.0 * avg);
  cfg.big_cap = (uint32_t)std::min<uint64_t>(cfg.B, (uint64_t)n_terms * cfg.W / cfg.big_thresh + 1);
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
                       uint32_t* d_out_ext) {
  if (nsets <= 0) return BPG_ERR_ARG;
  if (n_terms == 0) {
    // empty sum: identity for every set
    std::vector<uint32_t> id(32 * (size_t)nsets, 0);
    for (int s = 0; s < nsets; s++) id[32 * s + 8] = id[32 * s + 16] = 1;
    if ((size_t)nsets * 128 > SMALL_BYTES) return BPG_ERR_ARG;
    memcpy(ctx->h_pinned, id.data(), id.size() * 4);
    CK(cudaMemcpyAsync(d_out_ext, ctx->h_pinned, id.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BPG_OK;
  }
  if (n_terms >= (1u << 31)) return BPG_ERR_ARG;
  MsmCfg cfg;
  int c = pick_window((n_terms + nsets - 1) / nsets, ctx->forced_c);
  make_cfg(cfg, n_terms, n_points, nsets, c);
  if ((uint64_t)cfg.nwin * cfg.nb >= (1ull << 31)) return BPG_ERR_ARG;

  size_t ntiles = (cfg.B + SCAN_TILE - 1) / SCAN_TILE;
  size_t off = 0;
  size_t o_counts = off;  off += align_up((size_t)cfg.B * 4);
  size_t o_offsets = off; off += align_up(((size_t)cfg.B + 1) * 4);
  size_t o_tiles = off;   off += align_up(ntiles * 4);
  size_t o_big = off;     off += align_up(((size_t)cfg.big_cap + 1) * 4);
  size_t o_entries = off; off += align_up((size_t)n_terms * cfg.W * 4);
  size_t o_buckets = off; off += align_up((size_t)cfg.B * 128);
  size_t o_chunks = off;  off += align_up((size_t)cfg.nwin * cfg.nchunks * 128);
  size_t o_wins = off;    off += align_up((size_t)cfg.nwin * 128);
  int rc = ensure_ws(ctx, off);
  if (rc) return rc;
  uint32_t* counts = (uint32_t*)(ctx->ws + o_counts);
  uint32_t* offsets = (uint32_t*)(ctx->ws + o_offsets);
  uint32_t* tiles = (uint32_t*)(ctx->ws + o_tiles);
  uint32_t* big_count = (uint32_t*)(ctx->ws + o_big);
  uint32_t* big_list = big_count + 1;
  uint32_t* entries = (uint32_t*)(ctx->ws + o_entries);
  uint32_t* buckets = (uint32_t*)(ctx->ws + o_buckets);
  uint32_t* chunks = (uint32_t*)(ctx->ws + o_chunks);
  uint32_t* wins = (uint32_t*)(ctx->ws + o_wins);
  cudaStream_t st = ctx->stream;

  prof_mark(ctx, BPG_PROF_HIST);
  CK(cudaMemsetAsync(counts, 0, (size_t)cfg.B * 4, st));
  CK(cudaMemsetAsync(big_count, 0, 4, st));
  unsigned gt = (unsigned)((n_terms + 255) / 256);
  k_hist<<<gt, 256, 0, st>>>(d_scalars, d_set_ids, cfg, counts);
  LAUNCH_CHECK();
  prof_mark(ctx, BPG_PROF_SCAN);
  k_scan_tiles<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(counts, cfg.B, tiles);
  LAUNCH_CHECK();
  k_scan_spine<<<1, 1024, 0, st>>>(tiles, (uint32_t)ntiles, offsets, cfg.B);
  LAUNCH_CHECK();
  k_scan_apply<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(counts, cfg.B, tiles, offsets);
  LAUNCH_CHECK();
  prof_mark(ctx, BPG_PROF_SCATTER);
  k_scatter<<<gt, 256, 0, st>>>(d_scalars, d_set_ids, d_point_ids, cfg, offsets, counts, entries);
  LAUNCH_CHECK();
  prof_mark(ctx, BPG_PROF_ACCUM);
  k_accum<<<(cfg.B + ACC_THREADS - 1) / ACC_THREADS, ACC_THREADS, 0, st>>>(table_base, offsets, entries, cfg,
                                                                           buckets, big_count, big_list);
  LAUNCH_CHECK();
  prof_mark(ctx, BPG_PROF_ACCUM_BIG);
  unsigned gbig = std::min<unsigned>(cfg.big_cap, (unsigned)ctx->sm_count * 2);
  k_accum_big<<<gbig, BIG_THREADS, 0, st>>>(table_base, offsets, entries, cfg, buckets, big_count, big_list);
  LAUNCH_CHECK();
  prof_mark(ctx, BPG_PROF_REDUCE);
  unsigned nred = cfg.nwin * cfg.nchunks;
  k_reduce<<<(nred + RED_THREADS - 1) / RED_THREADS, RED_THREADS, 0, st>>>(buckets, cfg, chunks);
  LAUNCH_CHECK();
  prof_mark(ctx, BPG_PROF_COMBINE);
  k_combine<<<cfg.nwin, COMB_THREADS, 0, st>>>(chunks, cfg, wins);
  LAUNCH_CHECK();
  prof_mark(ctx, BPG_PROF_HORNER);
  k_horner<<<(nsets + 31) / 32, 32, 0, st>>>(wins, cfg, d_out_ext);
  LAUNCH_CHECK();
  prof_mark(ctx, -1);
  return BPG_OK;
}

extern "C" int bpg_dev_msm_table(bpg_ctx* ctx, const bpg_table* table, size_t offset, size_t n,
                                 const void* d_scalars, int n_sets, void* d_out_ext) {
  if (!ctx || !table || !d_out_ext || (!d_scalars && n) || n_sets <= 0) return BPG_ERR_ARG;
  if (offset + n > table->n) return BPG_ERR_CAPACITY;
  CK(cudaSetDevice(ctx->device));
  return msm_enqueue(ctx, table->niels + offset * 24, n, (const uint32_t*)d_scalars, n * (size_t)n_sets, nullptr,
                     nullptr, n_sets, (uint32_t*)d_out_ext);
}

extern "C" int bpg_dev_sum_encode(bpg_ctx* ctx, const void* d_parts, int n_parts, int n_sets, void* d_out_bytes,
                                  void* d_out_ext) {
  if (!ctx || !d_parts || n_parts <= 0 || n_sets <= 0) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  prof_mark(ctx, BPG_PROF_ENCODE);
  k_sum_encode<<<(n_sets + 31) / 32, 32, 0, ctx->stream>>>((const uint32_t*)d_parts, n_parts, n_sets,
                                                           (uint8_t*)d_out_bytes, (uint32_t*)d_out_ext);
  LAUNCH_CHECK();
  prof_mark(ctx, -1);
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
  if (sbytes) CK(cudaMemcpyAsync(ctx->d_stage, scalars_le, sbytes, cudaMemcpyHostToDevice, ctx->stream));
  uint32_t* d_ext = (uint32_t*)ctx->d_small;
  uint8_t* d_bytes = ctx->d_small + (size_t)n_sets * 128;
  rc = msm_enqueue(ctx, table->niels + offset * 24, n, (const uint32_t*)ctx->d_stage, n * (size_t)n_sets, nullptr,
                   nullptr, n_sets, d_ext);
  if (rc) return rc;
  rc = bpg_dev_sum_encode(ctx, d_ext, 1, n_sets, d_bytes, nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->h_pinned, d_bytes, (size_t)n_sets * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->h_pinned, (size_t)n_sets * 32);
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
  k_comb_mul<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(comb->tables, comb->nbases,
                                                                  (const uint32_t*)d_scalars, (uint32_t)n,
                                                                  bias_for(4), (uint8_t*)d_out_bytes,
                                                                  (uint32_t*)d_out_ext);
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
