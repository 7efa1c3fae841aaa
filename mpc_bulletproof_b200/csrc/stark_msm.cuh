// Stark-curve policy, MSM kernels: the same signed-digit Pippenger pipeline as
// msm_kernels.cuh — its sort (k_hist, scan, k_scatter) and its accumulation schedule
// (k_size_*: segments of <= 64 entries by decreasing length) are curve-agnostic and are reused as
// they are — with the bucket arithmetic of the short-Weierstrass policy (stark_pt.cuh):
//   k_stark_decode   affine x||y bytes -> validated Montgomery table entries
//   k_stark_accum    one thread per (bucket, segment): XYZZ += affine entry (8M + 2S)
//   k_stark_fix / k_stark_big / k_stark_big_fin   partial sums of long buckets
//   k_stark_leaf, k_stark_pairs   T_w = sum_j (j+1) B_j per window: pair hierarchy (A, Y) as in
//                    msm_kernels.cuh, one thread per chunk / per pair
//   k_stark_horner   sum_w 2^(c w) T_w
//   k_stark_finish   sum of the ranks' partial sums, to affine bytes
// Replaces `StarkPoint::msm_iter` / `::msm` (mpc-stark over ark-ec) at the call sites of
// SURVEY.md 2.2 for the instantiation the mounted fork itself uses.
#pragma once
#include "msm_sort_kernels.cuh"
#include "stark_pt.cuh"
#include "stark_pt4.cuh"

namespace bpg {

constexpr int SACC_THREADS = 128;
static __global__ void __launch_bounds__(SACC_THREADS, 1) k_stark_accum(const uint32_t* __restrict__ table,
                                                                  const uint32_t* __restrict__ offsets,
                                                                  const uint32_t* __restrict__ entries, AccSched sc,
                                                                  uint32_t* __restrict__ bucket_sums,
                                                                  uint32_t* __restrict__ seg_part) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *sc.n_items) return;
  uint2 it = sc.items[t];
  uint32_t b = it.x;
  uint32_t b_beg = offsets[b], b_end = offsets[b + 1];
  uint32_t beg = b_beg + it.y * ACC_SEG, end = min(b_end, beg + ACC_SEG);
  sp_xyzz acc = sp_identity();
  if (beg < end) {
    uint32_t e = __ldg(entries + beg);
    sp_aff q, qn;
    sp_aff_load(q, table + (size_t)(e & ~ENTRY_NEG) * 16);
    uint32_t e_next = beg + 1 < end ? __ldg(entries + beg + 1) : e;
    sp_aff_load(qn, table + (size_t)(e_next & ~ENTRY_NEG) * 16);
    q.y = fp_sel((e & ENTRY_NEG) != 0, fp_neg(q.y), q.y);
    acc = sp_from_aff(q);
    for (uint32_t i = beg + 1; i < end; i++) {
      e = e_next;
      q = qn;
      e_next = i + 1 < end ? __ldg(entries + i + 1) : e;
      sp_aff_load(qn, table + (size_t)(e_next & ~ENTRY_NEG) * 16);
      acc = sp_madd(acc, q, (e & ENTRY_NEG) != 0);
    }
  }
  if (b_end - b_beg <= ACC_SEG) sp_store(bucket_sums + (size_t)b * 32, acc);
  else sp_store(seg_part + (size_t)(sc.seg_slot[b] + it.y) * 32, acc);
}

// multi-segment buckets (<= BIG_SEG / ACC_SEG partial sums): one thread adds them
static __global__ void __launch_bounds__(128) k_stark_fix(const uint32_t* __restrict__ offsets, AccSched sc,
                                                    const uint32_t* __restrict__ seg_part,
                                                    uint32_t* __restrict__ bucket_sums) {
  uint32_t nmulti = *sc.multi_count;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nmulti; k += gridDim.x * blockDim.x) {
    uint32_t b = sc.multi_list[k];
    uint32_t nseg = acc_nseg(offsets[b + 1] - offsets[b]);
    const uint32_t* src = seg_part + (size_t)sc.seg_slot[b] * 32;
    sp_xyzz acc;
    sp_load(acc, src);
    for (uint32_t j = 1; j < nseg; j++) {
      sp_xyzz o;
      sp_load(o, src + (size_t)j * 32);
      acc = sp_add(acc, o);
    }
    sp_store(bucket_sums + (size_t)b * 32, acc);
  }
}

// over-long buckets: one block per segment of BIG_SEG entries, strided accumulation, shared-memory tree
constexpr int SBIG_THREADS = 256;
static __global__ void __launch_bounds__(SBIG_THREADS) k_stark_big(const uint32_t* __restrict__ table,
                                                             const uint32_t* __restrict__ offsets,
                                                             const uint32_t* __restrict__ entries, MsmCfg cfg,
                                                             uint32_t* __restrict__ bucket_sums,
                                                             const uint32_t* __restrict__ big_count,
                                                             const uint32_t* __restrict__ big_list,
                                                             uint32_t* __restrict__ big_part) {
  __shared__ uint32_t pts[SBIG_THREADS][32];
  uint32_t nbig = min(*big_count, cfg.big_cap);
  for (uint32_t k = blockIdx.x; k < nbig; k += gridDim.x) {
    uint32_t b = big_list[3 * (size_t)k], j = big_list[3 * (size_t)k + 1], nseg = big_list[3 * (size_t)k + 2];
    uint32_t beg = offsets[b] + j * BIG_SEG, end = min(offsets[b + 1], beg + BIG_SEG);
    sp_xyzz acc = sp_identity();
    for (uint32_t i = beg + threadIdx.x; i < end; i += SBIG_THREADS) {
      uint32_t e = __ldg(entries + i);
      sp_aff q;
      sp_aff_load(q, table + (size_t)(e & ~ENTRY_NEG) * 16);
      acc = sp_madd(acc, q, (e & ENTRY_NEG) != 0);
    }
    sp_store(pts[threadIdx.x], acc);
    __syncthreads();
    for (int half = SBIG_THREADS / 2; half >= 1; half >>= 1) {
      if (threadIdx.x < (uint32_t)half) {
        sp_xyzz x, y;
        sp_load(x, pts[threadIdx.x]);
        sp_load(y, pts[threadIdx.x + half]);
        sp_store(pts[threadIdx.x], sp_add(x, y));
      }
      __syncthreads();
    }
    if (threadIdx.x < 32) {
      uint32_t* dst = nseg == 1 ? bucket_sums + (size_t)b * 32 : big_part + (size_t)k * 32;
      dst[threadIdx.x] = pts[0][threadIdx.x];
    }
    __syncthreads();
  }
}
static __global__ void __launch_bounds__(128) k_stark_big_fin(MsmCfg cfg, uint32_t* __restrict__ bucket_sums,
                                                        const uint32_t* __restrict__ big_count,
                                                        const uint32_t* __restrict__ big_list,
                                                        const uint32_t* __restrict__ big_part) {
  uint32_t nbig = min(*big_count, cfg.big_cap);
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nbig; k += gridDim.x * blockDim.x) {
    uint32_t b = big_list[3 * (size_t)k], j = big_list[3 * (size_t)k + 1], nseg = big_list[3 * (size_t)k + 2];
    if (j != 0 || nseg == 1) continue;
    sp_xyzz acc;
    sp_load(acc, big_part + (size_t)k * 32);
    for (uint32_t s = 1; s < nseg; s++) {
      sp_xyzz o;
      sp_load(o, big_part + (size_t)(k + s) * 32);
      acc = sp_add(acc, o);
    }
    sp_store(bucket_sums + (size_t)b * 32, acc);
  }
}

// ---- bucket reduction: T = sum_{j<n} (j+1) X_j per array, pairs (A, Y) with T = sum A_q + sum q Y_q ----
constexpr uint32_t SLEAF_LC = 8;
static __global__ void __launch_bounds__(128) k_stark_leaf(const uint32_t* __restrict__ in /*[narr][n] XYZZ*/, uint32_t n,
                                                     uint32_t chunks, uint32_t narr, uint32_t* __restrict__ out_a,
                                                     uint32_t* __restrict__ out_y) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= narr * chunks) return;
  uint32_t arr = t / chunks, q = t % chunks;
  uint32_t first = q * SLEAF_LC;
  int valid = (int)min(SLEAF_LC, n - first);
  const uint32_t* src = in + ((size_t)arr * n + first) * 32;
  sp_xyzz run = sp_identity(), acc = sp_identity();
  for (int k = valid - 1; k >= 0; k--) {
    sp_xyzz x;
    sp_load(x, src + (size_t)k * 32);
    run = sp_add(run, x);
    acc = sp_add(acc, run);
  }
#pragma unroll
  for (uint32_t i = 1; i < SLEAF_LC; i <<= 1) run = sp_dbl(run);
  sp_store(out_a + (size_t)t * 32, acc);
  sp_store(out_y + (size_t)t * 32, run);
}

// up to 64 pairs per block -> one: binary tree A' = A0 + A1 + Y1, Y' = 2 (Y0 + Y1), one thread per output pair
constexpr uint32_t SPAIR_N = 64;
static __global__ void __launch_bounds__(SPAIR_N) k_stark_pairs(const uint32_t* __restrict__ in_a,
                                                          const uint32_t* __restrict__ in_y, uint32_t n,
                                                          uint32_t tiles, uint32_t* __restrict__ out_a,
                                                          uint32_t* __restrict__ out_y) {
  __shared__ uint32_t sa[SPAIR_N][32], sy[SPAIR_N][32];
  uint32_t arr = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  uint32_t first = tile * SPAIR_N;
  uint32_t m = min(SPAIR_N, n - first);
  const uint32_t* ga = in_a + ((size_t)arr * n + first) * 32;
  const uint32_t* gy = in_y + ((size_t)arr * n + first) * 32;
  for (uint32_t w = threadIdx.x; w < m * 32; w += blockDim.x) {
    sa[w >> 5][w & 31] = ga[w];
    sy[w >> 5][w & 31] = gy[w];
  }
  __syncthreads();
  while (m > 1) {
    uint32_t half = (m + 1) >> 1;
    sp_xyzz A, Y;
    bool live = threadIdx.x < half;
    if (live) {
      uint32_t i0 = 2 * threadIdx.x, i1 = i0 + 1;
      sp_xyzz a0, y0;
      sp_load(a0, sa[i0]);
      sp_load(y0, sy[i0]);
      if (i1 < m) {
        sp_xyzz a1, y1;
        sp_load(a1, sa[i1]);
        sp_load(y1, sy[i1]);
        A = sp_add(sp_add(a0, a1), y1);
        Y = sp_dbl(sp_add(y0, y1));
      } else {
        A = a0;
        Y = sp_dbl(y0);
      }
    }
    __syncthreads();
    if (live) {
      sp_store(sa[threadIdx.x], A);
      sp_store(sy[threadIdx.x], Y);
    }
    __syncthreads();
    m = half;
  }
  if (threadIdx.x < 32) {
    size_t o = ((size_t)arr * tiles + tile) * 32;
    out_a[o + threadIdx.x] = sa[0][threadIdx.x];
    out_y[o + threadIdx.x] = sy[0][threadIdx.x];
  }
}

// one thread per set: sum_w 2^(c w) S_w, top window first
static __global__ void __launch_bounds__(32) k_stark_horner(const uint32_t* __restrict__ window_sums, MsmCfg cfg,
                                                      uint32_t* __restrict__ out) {
  uint32_t set = blockIdx.x * blockDim.x + threadIdx.x;
  if (set >= (uint32_t)cfg.nsets) return;
  const uint32_t* src = window_sums + (size_t)set * cfg.W * 32;
  sp_xyzz acc;
  sp_load(acc, src + (size_t)(cfg.W - 1) * 32);
  for (int w = cfg.W - 2; w >= 0; w--) {
    for (int i = 0; i < cfg.c; i++) acc = sp_dbl(acc);
    sp_xyzz o;
    sp_load(o, src + (size_t)w * 32);
    acc = sp_add(acc, o);
  }
  sp_store(out + (size_t)set * 32, acc);
}

static __global__ void k_stark_set_identity(uint32_t* __restrict__ out) {
  out[(size_t)blockIdx.x * 32 + threadIdx.x] = 0u;
}

// windowed tables with several bucket groups per set: merged[set][b] = sum_g bucket_sums[set][g][b]
static __global__ void __launch_bounds__(128) k_stark_merge(const uint32_t* __restrict__ bucket_sums, MsmCfg cfg,
                                                      uint32_t* __restrict__ merged) {
  uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t total = (uint32_t)cfg.nsets * cfg.nb;
  if (q >= total) return;
  uint32_t set = q / cfg.nb, b = q % cfg.nb;
  const uint32_t* src = bucket_sums + ((size_t)set * cfg.gsub * cfg.nb + b) * 32;
  sp_xyzz acc;
  sp_load(acc, src);
  for (uint32_t g = 1; g < cfg.gsub; g++) {
    sp_xyzz o;
    sp_load(o, src + (size_t)g * cfg.nb * 32);
    acc = sp_add(acc, o);
  }
  sp_store(merged + (size_t)q * 32, acc);
}

// ---- quad-cooperative tails (stark_pt4.cuh): the same pair hierarchy with four lanes per point ----
constexpr int SRT_THREADS = 256;
constexpr int SRT_QUADS = SRT_THREADS / 4;

// in-block binary tree over `m` pairs in shared memory (m <= blockDim/4): A' = A0 + A1 + Y1, Y' = 2 (Y0 + Y1)
__device__ __forceinline__ void srt_block_tree(uint32_t (*sa)[32], uint32_t (*sy)[32], uint32_t m) {
  uint32_t quad = threadIdx.x >> 2, warp = threadIdx.x >> 5;
  while (m > 1) {
    uint32_t half = (m + 1) >> 1;
    bool warp_live = warp * 8 < half;  // warp-uniform
    sp4 A, Y;
    if (warp_live) {
      bool live = quad < half;
      uint32_t q = live ? quad : 0;
      bool have1 = 2 * q + 1 < m;
      uint32_t i0 = 2 * q, i1 = have1 ? 2 * q + 1 : 2 * q;
      sp4 a0 = sp4_load(sa[i0]), a1 = sp4_load(sa[i1]), y0 = sp4_load(sy[i0]), y1 = sp4_load(sy[i1]);
      a1.c = fp_sel(have1, a1.c, fp_zero());
      y1.c = fp_sel(have1, y1.c, fp_zero());
      A = sp4_add(sp4_add(a0, a1), y1);
      Y = sp4_dbl(sp4_add(y0, y1));
    }
    __syncthreads();
    if (warp_live && quad < half) {
      sp4_store(sa[quad], A);
      sp4_store(sy[quad], Y);
    }
    __syncthreads();
    m = half;
  }
}

// leaf pass, one quad per chunk of LC buckets, then the in-block tree: a block reduces 64 LC buckets to one pair
template <int LC>
__global__ void __launch_bounds__(SRT_THREADS) k_stark_leaf4(const uint32_t* __restrict__ in /*[narr][n] XYZZ*/,
                                                              uint32_t n, uint32_t tiles, uint32_t* __restrict__ out_a,
                                                              uint32_t* __restrict__ out_y) {
  __shared__ uint32_t sa[SRT_QUADS][32], sy[SRT_QUADS][32];
  uint32_t arr = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  uint32_t quad = threadIdx.x >> 2;
  uint32_t first = (tile * SRT_QUADS + quad) * LC;
  int valid = first >= n ? 0 : (int)min((uint32_t)LC, n - first);
  const uint32_t* src = in + ((size_t)arr * n + min(first, n - 1)) * 32;
  sp4 run = sp4_identity(), acc = sp4_identity();
#pragma unroll 2
  for (int k = LC - 1; k >= 0; k--) {
    bool have = k < valid;
    sp4 x = sp4_load(src + (size_t)(have ? k : 0) * 32);
    x.c = fp_sel(have, x.c, fp_zero());
    run = sp4_add(run, x);
    acc = sp4_add(acc, run);
  }
#pragma unroll
  for (int i = 1; i < LC; i <<= 1) run = sp4_dbl(run);
  sp4_store(sa[quad], acc);
  sp4_store(sy[quad], run);
  __syncthreads();
  srt_block_tree(sa, sy, SRT_QUADS);
  if (threadIdx.x < 32) {
    size_t o = ((size_t)arr * tiles + tile) * 32;
    out_a[o + threadIdx.x] = sa[0][threadIdx.x];
    out_y[o + threadIdx.x] = sy[0][threadIdx.x];
  }
}

// up to 64 pairs per block -> one (the final launch has tiles == 1 and writes T to out_a)
constexpr int SRP_THREADS = 128;
constexpr uint32_t SRP_PAIRS = SRP_THREADS / 2;
static __global__ void __launch_bounds__(SRP_THREADS) k_stark_pairs4(const uint32_t* __restrict__ in_a,
                                                               const uint32_t* __restrict__ in_y, uint32_t n,
                                                               uint32_t tiles, uint32_t* __restrict__ out_a,
                                                               uint32_t* __restrict__ out_y) {
  __shared__ uint32_t sa[SRP_PAIRS][32], sy[SRP_PAIRS][32];
  uint32_t arr = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  uint32_t first = tile * SRP_PAIRS;
  uint32_t m = min(SRP_PAIRS, n - first);
  const uint32_t* ga = in_a + ((size_t)arr * n + first) * 32;
  const uint32_t* gy = in_y + ((size_t)arr * n + first) * 32;
  for (uint32_t w = threadIdx.x; w < m * 32; w += blockDim.x) {
    sa[w >> 5][w & 31] = ga[w];
    sy[w >> 5][w & 31] = gy[w];
  }
  __syncthreads();
  srt_block_tree(sa, sy, m);
  if (threadIdx.x < 32) {
    size_t o = ((size_t)arr * tiles + tile) * 32;
    out_a[o + threadIdx.x] = sa[0][threadIdx.x];
    out_y[o + threadIdx.x] = sy[0][threadIdx.x];
  }
}

// plain tables, one warp per set: sum_w 2^(c w) S_w, every quad runs the same chain
static __global__ void __launch_bounds__(32) k_stark_horner4(const uint32_t* __restrict__ window_sums, MsmCfg cfg,
                                                       uint32_t* __restrict__ out) {
  uint32_t set = blockIdx.x;
  const uint32_t* src = window_sums + (size_t)set * cfg.W * 32;
  sp4 acc = sp4_load(src + (size_t)(cfg.W - 1) * 32);
  for (int w = cfg.W - 2; w >= 0; w--) {
    for (int i = 0; i < cfg.c; i++) acc = sp4_dbl(acc);
    acc = sp4_add(acc, sp4_load(src + (size_t)w * 32));
  }
  if (threadIdx.x < 4) sp4_store(out + (size_t)set * 32, acc);
}

}  // namespace bpg
