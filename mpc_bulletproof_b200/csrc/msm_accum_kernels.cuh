// Pippenger pipeline, bucket accumulation (see msm_sort_kernels.cuh for the pipeline overview).
#pragma once
#include "msm_sort_kernels.cuh"

namespace bpg {

// ---------------------------------------------------------------------------
// bucket accumulation: one thread per bucket
// ---------------------------------------------------------------------------
constexpr int ACC_THREADS = 128;

// bucket sums are parked in the "cached" operand layout of ge4_add_cached:
// [Y-X | Y+X | 2Z | 2dT], so that the reduction's first addition needs no conversion
__device__ __forceinline__ void ge_store_cached(uint32_t* p, const ge_ext& a) {
  fe_store(p, fe_sub(a.Y, a.X));
  fe_store(p + 8, fe_add_nc(a.Y, a.X));
  fe_store(p + 16, fe_add_nc(a.Z, a.Z));
  fe_store(p + 24, fe_mul(a.T, fe_const(BPG_K(K_D2))));
}


// Software pipeline: the Niels entry of step k+1 (and the entry word of step k+2) are loaded
// before step k multiplies.  Measured at 2^20 points (13.6 M additions): 0.899 ms, against 0.947 ms
// with only the entry word prefetched (112 registers), the same 0.948 ms when that form is
// compiled for 5 blocks/SM (96 registers: occupancy is not the limiter), and 0.925 ms with two
// entries in flight (150 registers).
static __global__ void __launch_bounds__(ACC_THREADS, 1) k_accum(const uint32_t* __restrict__ table,
                                                        const uint32_t* __restrict__ offsets,
                                                        const uint32_t* __restrict__ entries, AccSched sc,
                                                        uint32_t* __restrict__ bucket_sums,
                                                        uint32_t* __restrict__ seg_part /*[slots][32] ext*/) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *sc.n_items) return;
  uint2 it = sc.items[t];
  uint32_t b = it.x;
  uint32_t b_beg = offsets[b], b_end = offsets[b + 1];
  uint32_t beg = b_beg + it.y * ACC_SEG, end = min(b_end, beg + ACC_SEG);
  ge_ext acc = ge_identity();
  if (beg < end) {
    uint32_t e = __ldg(entries + beg);
    ge_niels q;
    ge_load_niels(q, table + (size_t)(e & ~ENTRY_NEG) * NIELS_WORDS);
    uint32_t e_next = beg + 1 < end ? __ldg(entries + beg + 1) : 0;
    acc = ge_from_niels(q, (e & ENTRY_NEG) != 0);
    ge_niels qn;
    if (beg + 1 < end) ge_load_niels(qn, table + (size_t)(e_next & ~ENTRY_NEG) * NIELS_WORDS);
    for (uint32_t i = beg + 1; i < end; i++) {
      e = e_next;
      q = qn;
      e_next = i + 1 < end ? __ldg(entries + i + 1) : e;
      ge_load_niels(qn, table + (size_t)(e_next & ~ENTRY_NEG) * NIELS_WORDS);  // next point (or a harmless re-read)
      acc = ge_madd(acc, q, (e & ENTRY_NEG) != 0);
    }
  }
  if (b_end - b_beg <= ACC_SEG) ge_store_cached(bucket_sums + (size_t)b * 32, acc);
  else ge_store_ext(seg_part + (size_t)(sc.seg_slot[b] + it.y) * 32, acc);
}

// multi-segment buckets: one quad adds the (at most BIG_SEG / ACC_SEG) partial sums
constexpr int FIX_THREADS = 128;
static __global__ void __launch_bounds__(FIX_THREADS) k_accum_fix(const uint32_t* __restrict__ offsets, AccSched sc,
                                                            const uint32_t* __restrict__ seg_part,
                                                            uint32_t* __restrict__ bucket_sums) {
  uint32_t nmulti = *sc.multi_count;
  uint32_t quads = gridDim.x * (FIX_THREADS / 4);
  uint32_t rounds = (nmulti + quads - 1) / quads;
  uint32_t q0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  for (uint32_t r = 0; r < rounds; r++) {
    uint32_t k = r * quads + q0;
    bool live = k < nmulti;
    uint32_t b = sc.multi_list[live ? k : 0];
    uint32_t nseg = acc_nseg(offsets[b + 1] - offsets[b]);
    // warp-uniform trip count (the quad arithmetic shuffles warp-wide)
    uint32_t nmax = nseg;
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1) nmax = max(nmax, __shfl_xor_sync(BPG_FULL_MASK, nmax, o));
    const uint32_t* src = seg_part + (size_t)sc.seg_slot[b] * 32;
    ge4 acc = ge4_identity();
    const ge4 id = ge4_identity();
    for (uint32_t j = 0; j < nmax; j++) {
      bool have = j < nseg;
      ge4 x = ge4_load(src + (size_t)(have ? j : 0) * 32);
      x.c = fe_sel(have, x.c, id.c);
      acc = ge4_add(acc, x);
    }
    acc = ge4_to_cached(acc);
    if (live) ge4_store(bucket_sums + (size_t)b * 32, acc);
  }
}

// over-long buckets (structured scalars: bit vectors, the nearly empty top window): one block per
// segment of BIG_SEG entries, strided accumulation, then a quad-cooperative block sum.  A bucket
// of one segment is finished here; longer ones leave per-segment partial sums for k_accum_big_fin.
constexpr int BIG_THREADS = 256;
static __global__ void __launch_bounds__(BIG_THREADS) k_accum_big(const uint32_t* __restrict__ table,
                                                            const uint32_t* __restrict__ offsets,
                                                            const uint32_t* __restrict__ entries, MsmCfg cfg,
                                                            uint32_t* __restrict__ bucket_sums,
                                                            const uint32_t* __restrict__ big_count,
                                                            const uint32_t* __restrict__ big_list,
                                                            uint32_t* __restrict__ big_part /*[big_cap][32] ext*/) {
  __shared__ uint32_t pts[BIG_THREADS][32];
  __shared__ uint32_t sm[BIG_THREADS / 32][32];
  uint32_t nbig = min(*big_count, cfg.big_cap);
  for (uint32_t k = blockIdx.x; k < nbig; k += gridDim.x) {
    uint32_t b = big_list[3 * (size_t)k], j = big_list[3 * (size_t)k + 1], nseg = big_list[3 * (size_t)k + 2];
    uint32_t beg = offsets[b] + j * BIG_SEG, end = min(offsets[b + 1], beg + BIG_SEG);
    ge_ext acc = ge_identity();
    for (uint32_t i = beg + threadIdx.x; i < end; i += BIG_THREADS) {
      uint32_t e = __ldg(entries + i);
      ge_niels q;
      ge_load_niels(q, table + (size_t)(e & ~ENTRY_NEG) * NIELS_WORDS);
      acc = ge_madd(acc, q, (e & ENTRY_NEG) != 0);
    }
    ge_store_ext(pts[threadIdx.x], acc);
    __syncthreads();
    // quad g sums points 4g..4g+3, then the block sum
    int g = threadIdx.x >> 2;
    ge4 t = ge4_load(pts[4 * g]);
#pragma unroll
    for (int jj = 1; jj < 4; jj++) t = ge4_add(t, ge4_load(pts[4 * g + jj]));
    t = block_sum_quads(t, sm);
    if (nseg == 1) {
      ge4 c = ge4_to_cached(t);  // park in cached layout like k_accum (all lanes: it shuffles)
      if (threadIdx.x < 4) ge4_store(bucket_sums + (size_t)b * 32, c);
    } else {
      if (threadIdx.x < 4) ge4_store(big_part + (size_t)k * 32, t);
    }
    __syncthreads();
  }
}
// buckets of several segments: the block that owns segment 0 sums the partials
static __global__ void __launch_bounds__(BIG_THREADS) k_accum_big_fin(MsmCfg cfg, uint32_t* __restrict__ bucket_sums,
                                                                const uint32_t* __restrict__ big_count,
                                                                const uint32_t* __restrict__ big_list,
                                                                const uint32_t* __restrict__ big_part) {
  __shared__ uint32_t sm[BIG_THREADS / 32][32];
  uint32_t nbig = min(*big_count, cfg.big_cap);
  for (uint32_t k = blockIdx.x; k < nbig; k += gridDim.x) {
    uint32_t b = big_list[3 * (size_t)k], j = big_list[3 * (size_t)k + 1], nseg = big_list[3 * (size_t)k + 2];
    if (j != 0 || nseg == 1) continue;  // block-uniform
    uint32_t quad = threadIdx.x >> 2;
    ge4 t = ge4_identity();
    for (uint32_t base = 0; base < nseg; base += BIG_THREADS / 4) {
      uint32_t i = base + quad;
      bool have = i < nseg;
      ge4 o = ge4_load(big_part + (size_t)(k + (have ? i : 0)) * 32);
      o.c = fe_sel(have, o.c, ge4_identity().c);
      t = ge4_add(t, o);
    }
    t = block_sum_quads(t, sm);
    ge4 c = ge4_to_cached(t);
    if (threadIdx.x < 4) ge4_store(bucket_sums + (size_t)b * 32, c);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// windowed tables: merged[set][b] = sum_g bucket_sums[set][g][b].  One quad per (set, bucket);
// operands and result in the cached layout.
// ---------------------------------------------------------------------------
constexpr int MERGE_THREADS = 128;
static __global__ void __launch_bounds__(MERGE_THREADS) k_merge(const uint32_t* __restrict__ bucket_sums, MsmCfg cfg,
                                                          uint32_t* __restrict__ merged) {
  uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  uint32_t total = (uint32_t)cfg.nsets * cfg.nb;
  bool live = q < total;
  uint32_t qq = live ? q : total - 1;  // idle quads shadow the last one: shuffles need every lane
  uint32_t set = qq / cfg.nb, b = qq % cfg.nb;
  const uint32_t* src = bucket_sums + ((size_t)set * cfg.gsub * cfg.nb + b) * 32;
  ge4 acc = ge4_identity();
  ge4 x = ge4_load(src);
  for (uint32_t g = 0; g < cfg.gsub; g++) {
    ge4 cur = x;
    if (g + 1 < cfg.gsub) x = ge4_load(src + (size_t)(g + 1) * cfg.nb * 32);
    acc = ge4_add_cached(acc, cur);
  }
  acc = ge4_to_cached(acc);
  if (live) ge4_store(merged + (size_t)qq * 32, acc);
}

}  // namespace bpg
