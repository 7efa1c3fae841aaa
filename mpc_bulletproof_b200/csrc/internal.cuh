// Internal declarations shared by the translation units of libbpgpu (not part of the ABI).
// Kernels are `static __global__` (BPG_GLOBAL): every translation unit compiles only the ones it launches.
#pragma once
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

#include "sc.cuh"

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
  // constraint terms uploaded ahead of their use (bpg_r1cs_terms_prefetch): consumed once by bpg_r1cs_dev_flatten
  uint8_t* d_terms = nullptr;
  size_t d_terms_cap = 0;
  bpg_terms terms_res = {};      // host arrays the resident copy was taken from (identity check)
  bool terms_resident = false;
  bpg_terms terms_pending = {};  // a prefetch waiting for the next commitment's own uploads to be queued first
  bool terms_is_pending = false;
  cudaEvent_t ev_terms = nullptr;
  // comb of the last Q given to a standalone InnerProductProof::create (callers tend to reuse one Q: building its
  // affine comb is a 1.2 ms chain of doublings and an inversion, a copy of 48 KB is not)
  uint32_t* q_cache_comb = nullptr;
  uint8_t q_cache_key[32] = {0};
  bool q_cache_valid = false;
  // combs of ad-hoc points built ahead of their MSM (bpg_adhoc_prefetch): [comp | bad, ticket | parts | ext | chain | comb]
  uint8_t* d_adhoc = nullptr;
  size_t adhoc_cap = 0;            // points the buffer holds
  size_t adhoc_n = 0;              // points of the resident combs (0: none)
  std::vector<uint8_t> adhoc_src;  // their encodings (identity check)
};

struct bpg_table {
  bpg_ctx* ctx;
  uint32_t* niels;  // n * 24 words; windowed: [W][n] * 24 words
  size_t n;
  int win_c = 0;    // 0: plain; otherwise the window width the multiples 2^(c w) P_i were built for
  int win_W = 1;
  uint32_t* comb = nullptr;  // optional: 64 x 8 affine-Niels multiples (d+1) 16^j P_i per point (bpg_table_build_comb)
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

cudaError_t dev_alloc(bpg_ctx* ctx, void** p, size_t bytes);
template <typename T>
static inline cudaError_t dev_alloc(bpg_ctx* ctx, T** p, size_t bytes) {
  return dev_alloc(ctx, reinterpret_cast<void**>(p), bytes);
}
void dev_free(bpg_ctx* ctx, void* p);
constexpr size_t SMALL_BYTES = 1 << 16;
void prof_mark(bpg_ctx* ctx, int phase);
int terms_issue_pending(bpg_ctx* ctx);  // r1cs_dev.inc
int ensure_ws(bpg_ctx* ctx, size_t bytes, int lane = 0);
int ensure_stage(bpg_ctx* ctx, size_t bytes);
static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
// tuning knobs read from the environment (documented where they are used)
static inline size_t env_size(const char* name, size_t dflt) {
  const char* e = getenv(name);
  return e ? (size_t)strtoull(e, nullptr, 10) : dflt;
}
int pick_window(size_t n_per_set, int forced);
int pick_window_table(size_t n, int forced);
int table_alloc_plain(bpg_ctx* ctx, size_t n, bpg_table** out);
static inline bpg::sc_bias bias_for(int c) {
  bpg::sc_bias b;
  memset(&b, 0, sizeof b);
  int W = (255 + c - 1) / c;
  for (int w = 0; w < W; w++) {
    int bit = c * w + c - 1;
    b.v[bit >> 5] |= 1u << (bit & 31);
  }
  return b;
}
// Enqueue one Pippenger launch (msm.cu).  d_scalars: n_terms*32 B; d_set_ids / d_point_ids may be null
// (implicit: term t -> point t % n_points of `table_base`, set t / n_points).
int msm_enqueue(bpg_ctx* ctx, const uint32_t* table_base, size_t n_points, const uint32_t* d_scalars, size_t n_terms,
                const uint8_t* d_set_ids, const uint32_t* d_point_ids, int nsets, uint32_t* d_out_ext, int win_c = 0,
                size_t win_stride = 0, int lane = 0, int curve = 0, const uint8_t* h_scalars = nullptr);
// one sum over ad-hoc compressed points + ranges of resident tables, scalars already on the device (core.cu)
int msm_mixed_core(bpg_ctx* ctx, const uint8_t* d_adhoc_points, size_t n_adhoc, const bpg_table* const* tabs,
                   const size_t* offs, const size_t* lens, int nsegs, const uint32_t* d_scalars, size_t total,
                   uint8_t out[32], bool identity_only = false, bool adhoc_resident = false);
// ad-hoc points with combs built ahead (bpg_adhoc_prefetch, core.cu): layout of ctx->d_adhoc for `cap` points
constexpr size_t ADHOC_MAX_POINTS = 256;
constexpr size_t ADHOC_PARTS = 64;
struct AdhocLayout {
  size_t comp, flags, parts, ext, chain, comb, total;
  explicit AdhocLayout(size_t cap) {
    comp = 0;
    flags = align_up(cap * 32);                       // [0] bad encodings, [1] ticket of the comb MSM
    parts = flags + 256;
    ext = parts + ADHOC_PARTS * 128;
    chain = ext + align_up(cap * 128);
    comb = chain + align_up(cap * 64 * 128);
    total = comb + cap * (size_t)512 * 128;
  }
};
bool adhoc_matches(const bpg_ctx* ctx, const uint8_t* host_points, size_t n);
// comb_build.cu: decode, doubling chains and cached combs of `n` compressed points
int comb_from_points(bpg_ctx* ctx, cudaStream_t s, const uint8_t* d_comp, size_t n, uint32_t* ext, uint32_t* chain,
                     uint32_t* comb, uint32_t* bad);
// ipp.cu: up to four sets of indexed terms over the affine combs of a table, encoded (set s: term single[s] and terms [lo[s], hi[s]))
int launch_comb_terms(bpg_ctx* ctx, cudaStream_t s, const uint32_t* comb_affine, const uint32_t* d_scalars,
                      const uint32_t* d_point_ids, const uint32_t single[4], const uint32_t lo[4], const uint32_t hi[4],
                      int nsets, uint8_t* d_out_bytes /*encodings, or null*/, uint32_t* d_out_ext /*extended sums, or null*/);
// ipp.cu: sum_k scalars[k] * P_k from cached combs (canonical scalars on the device), one extended point
int launch_comb_msm(bpg_ctx* ctx, cudaStream_t s, const uint32_t* comb_cached, const uint32_t* d_scalars, size_t n,
                    uint32_t* parts /*ADHOC_PARTS x 32 words*/, uint32_t* ticket, uint32_t* out_ext);
cudaError_t msm_kernels_init();
cudaError_t msm_sort_kernels_init();  // msm.cu  // msm.cu: function attributes of the pipeline kernels
// point-level kernels launched on behalf of other translation units (core.cu)
void launch_decode_to_niels(bpg_ctx* ctx, cudaStream_t s, const uint8_t* d_comp, size_t n, uint32_t* niels, uint32_t* bad);
void launch_comb_build(bpg_ctx* ctx, cudaStream_t s, const uint8_t* d_base32, uint32_t* table, uint32_t* bad);
void launch_comb_mul(bpg_ctx* ctx, cudaStream_t s, const uint32_t* tables, int nbases, const uint32_t* d_scalars, size_t n,
                     uint8_t* d_out_bytes, uint32_t* d_out_ext);
// folded generators of a long inner-product argument and their combs (comb_build.cu); `folded` holds
// COMB_MAT_SPLIT partial sums per point
constexpr uint32_t COMB_MAT_SPLIT = 2;
int comb_materialize(bpg_ctx* ctx, cudaStream_t s, const uint32_t* gen_comb, uint32_t g_id, uint32_t h_id, const uint32_t* wG,
                     const uint32_t* wH, size_t n, size_t m0, uint32_t* folded, uint32_t* chain, uint32_t* comb);
// IPP state over device-resident vectors (ipp.cu)
int ipp_begin_dev(bpg_ctx* ctx, const bpg_table* G, size_t g_off, const bpg_table* H, size_t h_off, size_t n,
                  const uint8_t* Q_host, const bpg_table* shared, size_t g_base, size_t h_base, size_t q_id,
                  const uint8_t* q_mul_host, const uint32_t* d_gf, const uint32_t* d_hf, const uint32_t* d_a,
                  const uint32_t* d_b, bpg_ipp** out, int lanes = 1);
constexpr int IPP_MAX_LANES = 4;
