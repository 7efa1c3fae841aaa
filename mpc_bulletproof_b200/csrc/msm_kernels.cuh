// Signed-digit windowed Pippenger multiscalar multiplication, sm_100a.
//
// Replaces `StarkPoint::msm_iter(scalars, points)` / `StarkPoint::msm(..)`
// (reference call sites: src/inner_product_proof.rs:90-114,159-172,353;
// src/r1cs/prover.rs:465-494,532-565; src/r1cs/verifier.rs:516-547) for the
// ristretto255 instantiation.
//
// Pipeline for one launch (T terms, `nsets` independent output sums that share
// one point table; term t belongs to set t / n_points unless set_ids is given):
//   k_hist      digits of every scalar -> per-bucket counts (atomics)
//   scan        exclusive prefix over the nsets*W*2^(c-1) buckets
//   k_scatter   counting-sort of (bucket -> point index | sign) entries
//   k_size_*    accumulation schedule: (bucket, segment <= 64 entries) items by decreasing length
//   k_accum     one thread per item: 7-mul mixed additions from the Niels table
//   k_accum_fix / k_accum_big  partial sums of multi-segment buckets; block-cooperative path for
//               over-long buckets (structured scalars)
//   k_merge     windowed tables: the sub-bucket groups of a bucket -> one sum per (set, bucket)
//   k_reduce_tree  radix-8 hierarchy of running sums: sum_j (j+1) B_j per bucket array
//   k_horner    plain tables: sum_w 2^(c w) S_w per set -> extended point (the partial sum a rank owns)
// A windowed table holds 2^(c w) P_i for every window, so all windows of a set feed ONE
// array of 2^(c-1) buckets: no doublings, no per-window reduction, no Horner.  The entries
// of a bucket are split into `gsub` groups (by window index) only to give the accumulation
// enough independent lists.
// All arithmetic is exact modular integer work; results are group elements, so
// any evaluation order gives the same canonical encoding.
#pragma once
#include "ge.cuh"
#include "ge4.cuh"
#include "fe16.cuh"
#include "sc.cuh"

namespace bpg {

struct MsmCfg {
  int c;              // window width in bits
  int W;              // windows per scalar = ceil(255 / c)
  uint32_t nb;        // buckets per window = 2^(c-1)
  int nsets;          // independent sums in this launch
  uint32_t n_terms;   // scalars in this launch
  uint32_t n_points;  // implicit indexing: term t -> point t % n_points, set t / n_points
  uint32_t gsub;      // bucket groups per set: window w accumulates into group w % gsub (plain tables: gsub = W)
  uint32_t narr;      // nsets * gsub bucket arrays of nb buckets
  uint32_t B;         // narr * nb
  uint32_t big_cap;     // capacity of the big-bucket list
  uint32_t win_stride;  // 0: plain table.  >0: table holds 2^(c w) P_i at index w*win_stride + i
  sc_bias bias;
};

constexpr uint32_t ENTRY_NEG = 0x80000000u;
constexpr uint32_t ACC_SEG = 64;     // entries per work item of k_accum
constexpr uint32_t SIZE_BINS = 128;  // size classes 0..ACC_SEG of the accumulation schedule
constexpr uint32_t BIG_SEG = 2048;   // entries of an over-long bucket handled by one block of k_accum_big

// ---------------------------------------------------------------------------
// digits -> histogram
// ---------------------------------------------------------------------------
// The recoded scalar is parked in shared memory (limb-major, conflict-free) so that a window's
// digit is two LDS and a funnel shift for a run-time window width, instead of a predicated
// scan over the nine limbs held in registers.
constexpr int SORT_THREADS = 256;
__device__ __forceinline__ void digits_park(uint32_t (*sh)[SORT_THREADS], const sc_recoded& r) {
#pragma unroll
  for (int i = 0; i < 9; i++) sh[i][threadIdx.x] = r.v[i];
}
__device__ __forceinline__ int digit_at(const uint32_t (*sh)[SORT_THREADS], int w, int c) {
  int bit = c * w;
  int limb = bit >> 5, s = bit & 31;
  uint32_t lo = sh[limb][threadIdx.x], hi = sh[limb + 1][threadIdx.x];  // limb <= 7: c (W - 1) <= 254
  uint32_t raw = __funnelshift_r(lo, hi, s) & ((1u << c) - 1u);
  return (int)raw - (1 << (c - 1));
}

// Warp-aggregated: lanes whose digit lands in the same bucket (structured scalars: bit vectors,
// repeated values, the short top window) issue ONE atomic for the group.
__global__ void __launch_bounds__(SORT_THREADS) k_hist(const uint32_t* __restrict__ scalars,
                                                       const uint8_t* __restrict__ set_ids, MsmCfg cfg,
                                                       uint32_t* __restrict__ counts, uint32_t t_begin,
                                                       uint32_t t_end /*this launch: terms [t_begin, t_end)*/) {
  __shared__ uint32_t sh[9][SORT_THREADS];
  uint32_t t = t_begin + blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < t_end;
  const uint32_t lane = threadIdx.x & 31;
  sc k = sc_zero();
  if (valid) sc_load(k, scalars + (size_t)t * 8);
  digits_park(sh, sc_recode(k.v, cfg.bias));  // each thread reads back only its own column: no barrier
  uint32_t set = (valid && cfg.nsets > 1) ? (set_ids ? set_ids[t] : t / cfg.n_points) : 0;
  uint32_t base = set * cfg.gsub * cfg.nb;
  uint32_t g = 0;
  for (int w = 0; w < cfg.W; w++) {
    int d = valid ? digit_at(sh, w, cfg.c) : 0;
    uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
    uint32_t b = base + g * cfg.nb + mag - 1;
    uint32_t peers = __match_any_sync(0xffffffffu, d != 0 ? b : 0xffffffffu - lane);
    if (d != 0 && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&counts[b], (uint32_t)__popc(peers));
    g = g + 1 == cfg.gsub ? 0 : g + 1;
  }
}

// ---------------------------------------------------------------------------
// exclusive scan of counts[B] -> offsets[B+1]; zeroes counts for reuse as cursors
// ---------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;  // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* smem /*[32+1]*/) {
  // returns exclusive prefix of v across the block; *total = block sum
  uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= (uint32_t)o) x += y;
  }
  if (lane == 31) smem[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t nw = blockDim.x >> 5;
    uint32_t s = lane < nw ? smem[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= (uint32_t)o) s += y;
    }
    if (lane < nw) smem[lane] = s;  // inclusive warp totals
    if (lane == nw - 1) smem[32] = s;
  }
  __syncthreads();
  uint32_t warp_base = wid ? smem[wid - 1] : 0;
  *total = smem[32];
  return warp_base + x - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t* __restrict__ counts, uint32_t B,
                                                             uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t smem[33];
  uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    uint32_t idx = base + i;
    s += idx < B ? counts[idx] : 0;
  }
  uint32_t total;
  block_exclusive_scan(s, &total, smem);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of tile_sums[ntiles] in place; writes grand total to offsets[B]
__global__ void __launch_bounds__(1024) k_scan_spine(uint32_t* __restrict__ tile_sums, uint32_t ntiles,
                                                     uint32_t* __restrict__ offsets, uint32_t B) {
  __shared__ uint32_t smem[33];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t start = 0; start < ntiles; start += blockDim.x) {
    uint32_t i = start + threadIdx.x;
    uint32_t v = i < ntiles ? tile_sums[i] : 0;
    uint32_t total;
    uint32_t ex = block_exclusive_scan(v, &total, smem);
    uint32_t carry = carry_s;
    if (i < ntiles) tile_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[B] = carry_s;
}

// Also the size histogram of the accumulation schedule (the list lengths pass through here):
// bins[r] += buckets whose last segment has r entries, bins[ACC_SEG] += full segments.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(uint32_t* __restrict__ counts, uint32_t B,
                                                             const uint32_t* __restrict__ tile_sums,
                                                             uint32_t* __restrict__ offsets,
                                                             uint32_t* __restrict__ bins /*[SIZE_BINS], zeroed*/) {
  __shared__ uint32_t smem[33];
  __shared__ uint32_t sh[SIZE_BINS];
  for (uint32_t i = threadIdx.x; i < SIZE_BINS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    uint32_t idx = base + i;
    v[i] = idx < B ? counts[idx] : 0;
    s += v[i];
  }
  uint32_t total;
  uint32_t ex = block_exclusive_scan(s, &total, smem) + tile_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    uint32_t idx = base + i;
    if (idx < B) {
      offsets[idx] = ex;
      counts[idx] = 0;
      uint32_t len = v[i];
      if (len <= BIG_SEG) {
        uint32_t full = len / ACC_SEG, rem = len % ACC_SEG;
        if (full) atomicAdd(&sh[ACC_SEG], full);
        if (rem || !full) atomicAdd(&sh[rem], 1u);
      }
    }
    ex += v[i];
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < SIZE_BINS; i += blockDim.x)
    if (sh[i]) atomicAdd(&bins[i], sh[i]);
}

// ---------------------------------------------------------------------------
// scatter entries into bucket order (order inside a bucket is irrelevant)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS) k_scatter(const uint32_t* __restrict__ scalars,
                                                          const uint8_t* __restrict__ set_ids,
                                                          const uint32_t* __restrict__ point_ids, MsmCfg cfg,
                                                          const uint32_t* __restrict__ offsets,
                                                          uint32_t* __restrict__ cursors, uint32_t* __restrict__ entries) {
  __shared__ uint32_t sh[9][SORT_THREADS];
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < cfg.n_terms;
  const uint32_t lane = threadIdx.x & 31;
  sc k = sc_zero();
  if (valid) sc_load(k, scalars + (size_t)t * 8);
  digits_park(sh, sc_recode(k.v, cfg.bias));
  uint32_t set = (valid && cfg.nsets > 1) ? (set_ids ? set_ids[t] : t / cfg.n_points) : 0;
  uint32_t pid = valid ? (point_ids ? point_ids[t] : t % cfg.n_points) : 0;
  uint32_t base = set * cfg.gsub * cfg.nb;
  uint32_t g = 0;
  // Windows go through in batches of four: the four cursor atomics (one per group of lanes that
  // share a bucket; the group's lanes take consecutive slots) and the four offset loads are in
  // flight together before any entry is written.
  for (int w0 = 0; w0 < cfg.W; w0 += 4) {
    uint32_t b[4], peers[4], first[4], off[4];
    int d[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int w = w0 + j;
      d[j] = (valid && w < cfg.W) ? digit_at(sh, min(w, cfg.W - 1), cfg.c) : 0;
      uint32_t mag = d[j] < 0 ? (uint32_t)(-d[j]) : (uint32_t)d[j];
      b[j] = base + g * cfg.nb + mag - 1;
      g = g + 1 == cfg.gsub ? 0 : g + 1;
      peers[j] = __match_any_sync(0xffffffffu, d[j] != 0 ? b[j] : 0xffffffffu - lane);
      first[j] = 0;
      off[j] = 0;
      if (d[j] != 0) {
        if (lane == (uint32_t)(__ffs(peers[j]) - 1)) first[j] = atomicAdd(&cursors[b[j]], (uint32_t)__popc(peers[j]));
        off[j] = __ldg(offsets + b[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint32_t f = __shfl_sync(0xffffffffu, first[j], __ffs(peers[j]) - 1);
      if (d[j] != 0) {
        uint32_t pos = off[j] + f + (uint32_t)__popc(peers[j] & ((1u << lane) - 1u));
        entries[pos] = (pid + (uint32_t)(w0 + j) * cfg.win_stride) | (d[j] < 0 ? ENTRY_NEG : 0u);
      }
    }
  }
}

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

// sum of one point per quad over the whole block -> quad 0 of warp 0.
// sm: [warps][32] words.  Every thread of the block must call it.
__device__ __forceinline__ ge4 block_sum_quads(ge4 p, uint32_t (*sm)[32]) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int off = 16; off >= 4; off >>= 1) {
    ge4 o;
#pragma unroll
    for (int i = 0; i < 8; i++) o.c.v[i] = __shfl_down_sync(BPG_FULL_MASK, p.c.v[i], off);
    p = ge4_add(p, o);
  }
  if (nw == 1) return p;
  if (lane < 4) ge4_store(sm[wid], p);
  __syncthreads();
  if (wid == 0) {
    int quad = lane >> 2;
    ge4 t = ge4_identity();
    // up to 32 warps: each quad folds warps quad, quad+8, ...
    for (int k = 0; k < (nw + 7) / 8; k++) {
      int w = quad + 8 * k;
      ge4 o = w < nw ? ge4_load(sm[w]) : ge4_identity();
      t = ge4_add(t, o);
    }
#pragma unroll
    for (int off = 16; off >= 4; off >>= 1) {
      ge4 o;
#pragma unroll
      for (int i = 0; i < 8; i++) o.c.v[i] = __shfl_down_sync(BPG_FULL_MASK, t.c.v[i], off);
      t = ge4_add(t, o);
    }
    p = t;
  }
  __syncthreads();
  return p;
}

// Accumulation schedule.  The work item of k_accum is a SEGMENT: at most ACC_SEG consecutive
// entries of one bucket.  Items are ordered by decreasing length, so the 32 items of a warp have
// (almost) the same trip count and the longest start first.  A bucket of one segment is
// finished by its thread; a longer one (structured scalars, or the short top window whose few
// occupied buckets are long) leaves per-segment partial sums that k_accum_fix adds; beyond
// BIG_SEG entries the block-cooperative k_accum_big takes over.
struct AccSched {
  uint32_t* bins;        // [SIZE_BINS] class counts (k_scan_apply)
  uint32_t* cursors;     // [SIZE_BINS] items handed out per class (zeroed)
  uint32_t* n_items;     // total work items
  uint2* items;          // (bucket, segment)
  uint32_t* seg_slot;    // [B] first partial-sum slot of a multi-segment bucket
  uint32_t* part_count;  // partial-sum slots handed out
  uint32_t* multi_count; // multi-segment buckets
  uint32_t* multi_list;  // their ids
};
__device__ __forceinline__ uint32_t acc_nseg(uint32_t len) { return len == 0 ? 1u : (len + ACC_SEG - 1) / ACC_SEG; }

__global__ void __launch_bounds__(256) k_size_scatter(const uint32_t* __restrict__ offsets, MsmCfg cfg, AccSched sc,
                                                      uint32_t* __restrict__ big_count,
                                                      uint32_t* __restrict__ big_list) {
  // class start positions, longest class first: every block scans the 128 class counts itself
  // (no separate single-block launch); block 0 publishes the total
  __shared__ uint32_t cnt[SIZE_BINS];
  __shared__ uint32_t base[SIZE_BINS];
  __shared__ uint32_t start[SIZE_BINS];
  __shared__ uint32_t smem[33];
  {
    uint32_t i = threadIdx.x;
    uint32_t v = i < SIZE_BINS ? sc.bins[SIZE_BINS - 1 - i] : 0;  // reversed
    uint32_t total;
    uint32_t ex = block_exclusive_scan(v, &total, smem);
    if (i < SIZE_BINS) {
      start[SIZE_BINS - 1 - i] = ex;
      cnt[i] = 0;
    }
    if (blockIdx.x == 0 && i == 0) *sc.n_items = total;
  }
  __syncthreads();
  // block-private histogram first: one global atomic per (block, occupied size class)
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t len = 0, full = 0, rem = 0, rank_full = 0, rank_rem = 0;
  bool small = false;
  if (b < cfg.B) {
    len = offsets[b + 1] - offsets[b];
    if (len > BIG_SEG) {
      // over-long bucket: hand it to k_accum_big in segments of BIG_SEG entries
      uint32_t nseg = (len + BIG_SEG - 1) / BIG_SEG;
      uint32_t slot = atomicAdd(big_count, nseg);
      for (uint32_t j = 0; j < nseg && slot + j < cfg.big_cap; j++) {
        big_list[3 * (size_t)(slot + j)] = b;
        big_list[3 * (size_t)(slot + j) + 1] = j;
        big_list[3 * (size_t)(slot + j) + 2] = nseg;
      }
    } else {
      small = true;
      full = len / ACC_SEG;
      rem = len % ACC_SEG;
      if (full) rank_full = atomicAdd(&cnt[ACC_SEG], full);
      if (rem || !full) rank_rem = atomicAdd(&cnt[rem], 1u);
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < SIZE_BINS; i += blockDim.x)
    if (cnt[i]) base[i] = start[i] + atomicAdd(&sc.cursors[i], cnt[i]);
  __syncthreads();
  if (small) {
    for (uint32_t j = 0; j < full; j++) sc.items[base[ACC_SEG] + rank_full + j] = make_uint2(b, j);
    if (rem || !full) sc.items[base[rem] + rank_rem] = make_uint2(b, full);
    if (acc_nseg(len) > 1) {
      sc.seg_slot[b] = atomicAdd(sc.part_count, acc_nseg(len));
      sc.multi_list[atomicAdd(sc.multi_count, 1u)] = b;
    }
  }
}

BPG_DEF_CONST(K_DINV, 0xcdc9f843u, 0x25e0f276u, 0x4279542eu, 0x0b5dd698u, 0xcdb9cf66u, 0x2b162114u, 0x14d5ce43u,
              0x40907ed2u)  // 1/d

// the point (+-) of a Niels entry as an extended point with Z = 2: one multiplication
__device__ __forceinline__ ge_ext ge_from_niels(const ge_niels& q, bool neg) {
  ge_ext r;
  fe x2 = fe_sub(q.ypx, q.ymx);   // 2x
  fe t = fe_mul(q.t2d, fe_const(BPG_K(K_DINV)));  // 2xy
  r.X = fe_canon(fe_cneg(x2, neg));
  r.Y = fe_canon(fe_add(q.ypx, q.ymx));  // 2y
  r.Z = fe_zero();
  r.Z.v[0] = 2;
  r.T = fe_canon(fe_cneg(t, neg));
  return r;
}

// Software pipeline: the Niels entry of step k+1 (and the entry word of step k+2) are loaded
// before step k multiplies.  Measured at 2^20 points (13.6 M additions): 0.899 ms, against 0.947 ms
// with only the entry word prefetched (112 registers), the same 0.948 ms when that form is
// compiled for 5 blocks/SM (96 registers: occupancy is not the limiter), and 0.925 ms with two
// entries in flight (150 registers).
__global__ void __launch_bounds__(ACC_THREADS, 1) k_accum(const uint32_t* __restrict__ table,
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
    ge_load_niels(q, table + (size_t)(e & ~ENTRY_NEG) * 24);
    uint32_t e_next = beg + 1 < end ? __ldg(entries + beg + 1) : 0;
    acc = ge_from_niels(q, (e & ENTRY_NEG) != 0);
    ge_niels qn;
    if (beg + 1 < end) ge_load_niels(qn, table + (size_t)(e_next & ~ENTRY_NEG) * 24);
    for (uint32_t i = beg + 1; i < end; i++) {
      e = e_next;
      q = qn;
      e_next = i + 1 < end ? __ldg(entries + i + 1) : e;
      ge_load_niels(qn, table + (size_t)(e_next & ~ENTRY_NEG) * 24);  // next point (or a harmless re-read)
      acc = ge_madd(acc, q, (e & ENTRY_NEG) != 0);
    }
  }
  if (b_end - b_beg <= ACC_SEG) ge_store_cached(bucket_sums + (size_t)b * 32, acc);
  else ge_store_ext(seg_part + (size_t)(sc.seg_slot[b] + it.y) * 32, acc);
}

// multi-segment buckets: one quad adds the (at most BIG_SEG / ACC_SEG) partial sums
constexpr int FIX_THREADS = 128;
__global__ void __launch_bounds__(FIX_THREADS) k_accum_fix(const uint32_t* __restrict__ offsets, AccSched sc,
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
__global__ void __launch_bounds__(BIG_THREADS) k_accum_big(const uint32_t* __restrict__ table,
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
      ge_load_niels(q, table + (size_t)(e & ~ENTRY_NEG) * 24);
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
__global__ void __launch_bounds__(BIG_THREADS) k_accum_big_fin(MsmCfg cfg, uint32_t* __restrict__ bucket_sums,
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
__global__ void __launch_bounds__(MERGE_THREADS) k_merge(const uint32_t* __restrict__ bucket_sums, MsmCfg cfg,
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

// ---------------------------------------------------------------------------
// bucket reduction  T = sum_{j<n} (j+1) X_j  for `narr` independent arrays.
//
// Invariant carried between levels: T = sum_q A_q + sum_q q Y_q over pairs (A_q, Y_q).
//  * leaf pass (serial, one QUAD per chunk of LC buckets, ge4.cuh):
//      A_q = sum_k (k+1) X_{LC q + k},   Y_q = LC sum_k X_{LC q + k}
//  * binary tree step (pairs 2q', 2q'+1 -> q'):
//      A' = A_0 + A_1 + Y_1,             Y' = 2 (Y_0 + Y_1)
// The tree keeps the dependent chain at lg n steps of two additions; the serial leaf pass keeps
// the total work near 5 multiplication levels per bucket.  k_reduce_leaf: a block reduces
// 64 LC buckets to one pair; k_reduce_pairs: a block reduces up to 256 pairs to one.
// Missing items are the identity.  Every lane of a warp runs the same instruction stream
// (the quad arithmetic shuffles warp-wide); idle quads compute on clamped addresses and do not store.
// ---------------------------------------------------------------------------
constexpr int RT_THREADS = 256;
constexpr int RT_QUADS = RT_THREADS / 4;

__device__ __forceinline__ ge4 ge4_identity_cached() {
  int q = threadIdx.x & 3;
  ge4 r;
  r.c = fe_zero();
  r.c.v[0] = q == 3 ? 0u : (q == 2 ? 2u : 1u);  // (Y-X, Y+X, 2Z, 2dT) = (1, 1, 2, 0)
  return r;
}

// in-block binary tree over `m` pairs held in shared memory (ext layout), m <= blockDim/4, any m >= 1.
// Result in sa[0], sy[0].
__device__ __forceinline__ void rt_block_tree(uint32_t (*sa)[32], uint32_t (*sy)[32], uint32_t m) {
  uint32_t quad = threadIdx.x >> 2, warp = threadIdx.x >> 5;
  while (m > 1) {
    uint32_t half = (m + 1) >> 1;
    bool warp_live = warp * 8 < half;  // warp-uniform
    ge4 A, Y;
    if (warp_live) {
      bool live = quad < half;
      uint32_t q = live ? quad : 0;
      bool have1 = 2 * q + 1 < m;
      uint32_t i0 = 2 * q, i1 = have1 ? 2 * q + 1 : 2 * q;
      ge4 a0 = ge4_load(sa[i0]), a1 = ge4_load(sa[i1]), y0 = ge4_load(sy[i0]), y1 = ge4_load(sy[i1]);
      const ge4 id = ge4_identity();
      a1.c = fe_sel(have1, a1.c, id.c);
      y1.c = fe_sel(have1, y1.c, id.c);
      ge4 y1c = ge4_to_cached(y1);
      A = ge4_add_cached(ge4_add(a0, a1), y1c);
      Y = ge4_dbl(ge4_add_cached(y0, y1c));
    }
    __syncthreads();
    if (warp_live && quad < half) {
      ge4_store(sa[quad], A);
      ge4_store(sy[quad], Y);
    }
    __syncthreads();
    m = half;
  }
}

template <int LC>
__global__ void __launch_bounds__(RT_THREADS) k_reduce_leaf(const uint32_t* __restrict__ in /*[narr][n] cached*/,
                                                             uint32_t n, uint32_t tiles,
                                                             uint32_t* __restrict__ out_a, uint32_t* __restrict__ out_y) {
  __shared__ uint32_t sa[RT_QUADS][32], sy[RT_QUADS][32];
  uint32_t arr = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  uint32_t quad = threadIdx.x >> 2;
  uint32_t first = (tile * RT_QUADS + quad) * LC;
  int valid = first >= n ? 0 : (int)min((uint32_t)LC, n - first);
  const uint32_t* src = in + ((size_t)arr * n + min(first, n - 1)) * 32;
  ge4 run = ge4_identity(), acc = ge4_identity();
  const ge4 idc = ge4_identity_cached();
#pragma unroll 4
  for (int k = LC - 1; k >= 0; k--) {
    bool have = k < valid;
    ge4 x = ge4_load(src + (size_t)(have ? k : 0) * 32);
    x.c = fe_sel(have, x.c, idc.c);
    run = ge4_add_cached(run, x);
    acc = ge4_add(acc, run);
  }
#pragma unroll
  for (int i = 1; i < LC; i <<= 1) run = ge4_dbl(run);
  ge4_store(sa[quad], acc);
  ge4_store(sy[quad], run);
  __syncthreads();
  rt_block_tree(sa, sy, RT_QUADS);
  if (threadIdx.x < 32) {
    size_t o = ((size_t)arr * tiles + tile) * 32;
    out_a[o + threadIdx.x] = sa[0][threadIdx.x];
    out_y[o + threadIdx.x] = sy[0][threadIdx.x];
  }
}

// Large bucket arrays (>= 2^17): the leaf pass is throughput-bound, so ONE THREAD owns a chunk
// (8 + 9 multiplications per bucket instead of five quad levels) and writes its pair; the
// binary tree over the pairs is k_reduce_pairs.
constexpr int RL_THREADS = 128;
// p + q with q in the cached layout (Y-X, Y+X, 2Z, 2dT): 8 multiplications
__device__ __forceinline__ ge_ext ge_add_cached(const ge_ext& p, const fe& ymx, const fe& ypx, const fe& z2,
                                                const fe& t2d) {
  fe A = fe_mul(fe_sub(p.Y, p.X), ymx);
  fe B = fe_mul(fe_add_nc(p.Y, p.X), ypx);
  fe C = fe_mul(p.T, t2d);
  fe D = fe_mul(p.Z, z2);
  fe E = fe_sub(B, A), H = fe_add_nc(B, A), F = fe_sub(D, C), G = fe_add(D, C);
  ge_ext r;
  r.X = fe_mul(E, F);
  r.Y = fe_mul(G, H);
  r.Z = fe_mul(F, G);
  r.T = fe_mul(E, H);
  return r;
}
template <int LC>
__global__ void __launch_bounds__(RL_THREADS) k_reduce_leaf_thread(const uint32_t* __restrict__ in /*[narr][n] cached*/,
                                                                   uint32_t n, uint32_t chunks /*per array*/,
                                                                   uint32_t narr, uint32_t* __restrict__ out_a,
                                                                   uint32_t* __restrict__ out_y) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= narr * chunks) return;
  uint32_t arr = t / chunks, q = t % chunks;
  uint32_t first = q * LC;
  int valid = (int)min((uint32_t)LC, n - first);  // chunks = ceil(n / LC): first < n
  const uint32_t* src = in + ((size_t)arr * n + first) * 32;
  ge_ext run = ge_identity(), acc = ge_identity();
  for (int k = valid - 1; k >= 0; k--) {
    fe ymx, ypx, z2, t2d;
    fe_load(ymx, src + (size_t)k * 32);
    fe_load(ypx, src + (size_t)k * 32 + 8);
    fe_load(z2, src + (size_t)k * 32 + 16);
    fe_load(t2d, src + (size_t)k * 32 + 24);
    run = ge_add_cached(run, ymx, ypx, z2, t2d);
    acc = ge_add(acc, run);
  }
#pragma unroll
  for (int i = 1; i < LC; i <<= 1) run = ge_dbl(run);
  ge_store_ext(out_a + (size_t)t * 32, acc);
  ge_store_ext(out_y + (size_t)t * 32, run);
}

// up to RP_THREADS/2 = 64 pairs per block -> one pair (the final launch has tiles == 1 and writes T to out_a)
constexpr int RP_THREADS = 128;
constexpr uint32_t RP_PAIRS = RP_THREADS / 2;
__global__ void __launch_bounds__(RP_THREADS) k_reduce_pairs(const uint32_t* __restrict__ in_a,
                                                              const uint32_t* __restrict__ in_y, uint32_t n,
                                                              uint32_t tiles, uint32_t* __restrict__ out_a,
                                                              uint32_t* __restrict__ out_y) {
  __shared__ uint32_t sa[RP_PAIRS][32], sy[RP_PAIRS][32];
  uint32_t arr = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  uint32_t first = tile * RP_PAIRS;
  uint32_t m = min(RP_PAIRS, n - first);
  const uint32_t* ga = in_a + ((size_t)arr * n + first) * 32;
  const uint32_t* gy = in_y + ((size_t)arr * n + first) * 32;
  for (uint32_t w = threadIdx.x; w < m * 32; w += blockDim.x) {
    sa[w >> 5][w & 31] = ga[w];
    sy[w >> 5][w & 31] = gy[w];
  }
  __syncthreads();
  rt_block_tree(sa, sy, m);
  if (threadIdx.x < 32) {
    size_t o = ((size_t)arr * tiles + tile) * 32;
    out_a[o + threadIdx.x] = sa[0][threadIdx.x];
    out_y[o + threadIdx.x] = sy[0][threadIdx.x];
  }
}

// The last levels in ONE block: up to RPB_PAIRS pairs of an array -> its total.  A tree level costs
// its depth (about ten dependent field products), not its width, so finishing 256 pairs here
// takes 8 levels where two k_reduce_pairs launches took 6 + 2 (+ a launch gap and a round trip
// through global memory).  Dynamic shared memory: 2 x RPB_PAIRS x 128 bytes.
constexpr int RPB_THREADS = 512;  // 128 registers per thread stay available to the quad arithmetic
constexpr uint32_t RPB_PAIRS = 256;
constexpr size_t RPB_SMEM = 2 * (size_t)RPB_PAIRS * 128;
__global__ void __launch_bounds__(RPB_THREADS) k_reduce_pairs_final(const uint32_t* __restrict__ in_a,
                                                                     const uint32_t* __restrict__ in_y, uint32_t n,
                                                                     uint32_t* __restrict__ out_a) {
  extern __shared__ __align__(16) uint32_t rpb_smem[];
  uint32_t(*sa)[32] = reinterpret_cast<uint32_t(*)[32]>(rpb_smem);
  uint32_t(*sy)[32] = reinterpret_cast<uint32_t(*)[32]>(rpb_smem + RPB_PAIRS * 32);
  uint32_t arr = blockIdx.x;
  const uint4* ga = reinterpret_cast<const uint4*>(in_a + (size_t)arr * n * 32);
  const uint4* gy = reinterpret_cast<const uint4*>(in_y + (size_t)arr * n * 32);
  uint4* da = reinterpret_cast<uint4*>(rpb_smem);
  uint4* dy = reinterpret_cast<uint4*>(rpb_smem + RPB_PAIRS * 32);
  for (uint32_t w = threadIdx.x; w < n * 8; w += blockDim.x) {
    da[w] = ga[w];
    dy[w] = gy[w];
  }
  __syncthreads();
  rt_block_tree(sa, sy, n);
  if (threadIdx.x < 32) out_a[(size_t)arr * 32 + threadIdx.x] = sa[0][threadIdx.x];
}

// ---------------------------------------------------------------------------
// A few ad-hoc terms (the proof points of a verification, an `msm_iter` over a handful of points;
// reference src/r1cs/verifier.rs:516-547, src/inner_product_proof.rs:359-371): nothing is
// precomputed for them, so the cost is the 252-doubling chain of a scalar multiplication, and
// the sort / bucket / tree pipeline (fifteen dependent launches) only adds to it.  Here ONE QUAD
// per term walks the chain -- signed 4-bit windows, eight cached multiples of the point in
// shared memory, 4 doublings + 1 addition per window at two multiplication levels each -- and a
// tree over the block's quads adds the terms of each set.  One launch (+ one to add the blocks'
// sums when there are more than 32 terms).
// ---------------------------------------------------------------------------
constexpr int COMB_WINDOWS = 64;  // signed 4-bit windows of a 256-bit scalar (also the fixed-base comb)
constexpr int SMALL_THREADS = 128;
constexpr int SMALL_QUADS = SMALL_THREADS / 4;
constexpr uint32_t SMALL_MAX_TERMS = 1024;  // 32 blocks: what k_msm_small_fin adds in one pass
constexpr int SMALL_MAX_SETS = 4;

// sum over the block's quads of `mine` for the quads whose `member` is set; result in every quad that
// reads slot 0 afterwards (sm: [SMALL_QUADS][32] words)
__device__ __forceinline__ ge4 small_block_sum(ge4 mine, bool member, uint32_t (*sm)[32]) {
  uint32_t quad = threadIdx.x >> 2;
  ge4 v;
  v.c = fe_sel(member, mine.c, ge4_identity().c);
  __syncthreads();
  ge4_store(sm[quad], v);
  __syncthreads();
  for (uint32_t m = SMALL_QUADS; m > 1; m >>= 1) {
    uint32_t half = m >> 1;
    uint32_t q = quad < half ? quad : 0;
    ge4 a = ge4_load(sm[2 * q]), b = ge4_load(sm[2 * q + 1]);
    ge4 r = ge4_add(a, b);
    __syncthreads();
    if (quad < half) ge4_store(sm[quad], r);
    __syncthreads();
  }
  return ge4_load(sm[0]);
}

__global__ void __launch_bounds__(SMALL_THREADS) k_msm_small(const uint32_t* __restrict__ table /*affine Niels*/,
                                                              const uint32_t* __restrict__ scalars,
                                                              const uint8_t* __restrict__ set_ids,
                                                              const uint32_t* __restrict__ point_ids, uint32_t n_terms,
                                                              uint32_t n_points, int nsets, sc_bias bias4,
                                                              uint32_t* __restrict__ out /*[gridDim.x][nsets][32] ext*/) {
  __shared__ __align__(16) uint32_t mult[SMALL_QUADS][8][32];  // cached multiples 1..8 of each quad's point
  __shared__ __align__(16) uint32_t red[SMALL_QUADS][32];
  const uint32_t quad = threadIdx.x >> 2;
  const int q = threadIdx.x & 3;
  uint32_t t = blockIdx.x * SMALL_QUADS + quad;
  const bool live = t < n_terms;
  if (!live) t = 0;  // idle quads shadow term 0 (every lane takes part in the shuffles) and add nothing
  const uint32_t pid = point_ids ? point_ids[t] : t % n_points;
  const uint32_t set = nsets > 1 ? (set_ids ? set_ids[t] : t / n_points) : 0;
  sc k;
  sc_load(k, scalars + (size_t)t * 8);
  const sc_recoded rec = sc_recode(k.v, bias4);
  // the point, one coordinate per lane
  ge_niels nq;
  ge_load_niels(nq, table + (size_t)pid * 24);
  ge_ext pe = ge_from_niels(nq, false);
  ge4 P;
  P.c = q == 0 ? pe.X : (q == 1 ? pe.Y : (q == 2 ? pe.Z : pe.T));
  const ge4 Pc = ge4_to_cached(P);
  ge4 run = P;
  ge4_store(mult[quad][0], Pc);
#pragma unroll 1
  for (int d = 1; d < 8; d++) {
    run = ge4_add_cached(run, Pc);
    ge4_store(mult[quad][d], ge4_to_cached(run));
  }
  __syncwarp();
  const ge4 idc = ge4_identity_cached();
  ge4 acc = ge4_identity();
#pragma unroll 1
  for (int j = COMB_WINDOWS - 1; j >= 0; j--) {
    if (j != COMB_WINDOWS - 1) {
      acc = ge4_dbl(acc);
      acc = ge4_dbl(acc);
      acc = ge4_dbl(acc);
      acc = ge4_dbl(acc);
    }
    int d = sc_digit(rec, j, 4);
    int mag = d < 0 ? -d : d;
    ge4 m = ge4_load(mult[quad][mag ? mag - 1 : 0]);
    // -(cached): Y-X <-> Y+X, 2dT -> -2dT
    fe other = fe_quad_get(m.c, q ^ 1);
    fe neg = q < 2 ? other : (q == 3 ? fe_neg(m.c) : m.c);
    m.c = fe_sel(d < 0, neg, m.c);
    m.c = fe_sel(mag != 0, m.c, idc.c);
    acc = ge4_add_cached(acc, m);
  }
  for (int s = 0; s < nsets; s++) {
    ge4 tot = small_block_sum(acc, live && set == (uint32_t)s, red);
    if (threadIdx.x < 4) ge4_store(out + ((size_t)blockIdx.x * nsets + s) * 32, tot);
  }
}
// out[s] = sum_b parts[b][s], b < nblocks <= SMALL_QUADS
__global__ void __launch_bounds__(SMALL_THREADS) k_msm_small_fin(const uint32_t* __restrict__ parts, uint32_t nblocks,
                                                                  int nsets, uint32_t* __restrict__ out) {
  __shared__ __align__(16) uint32_t red[SMALL_QUADS][32];
  const uint32_t quad = threadIdx.x >> 2;
  for (int s = 0; s < nsets; s++) {
    bool have = quad < nblocks;
    ge4 v = ge4_load(parts + ((size_t)(have ? quad : 0) * nsets + s) * 32);
    ge4 tot = small_block_sum(v, have, red);
    if (threadIdx.x < 4) ge4_store(out + (size_t)s * 32, tot);
  }
}

// plain tables, one warp per set: sum_w 2^(c w) S_w (Horner, top window first)
__global__ void __launch_bounds__(32) k_horner(const uint32_t* __restrict__ window_sums, MsmCfg cfg,
                                                uint32_t* __restrict__ out_ext) {
  uint32_t set = blockIdx.x;
  const uint32_t* src = window_sums + (size_t)set * cfg.W * 32;
  // every quad runs the same chain (redundantly): the cost is the chain, not the lanes
  ge4 acc = ge4_load(src + (size_t)(cfg.W - 1) * 32);
  for (int w = cfg.W - 2; w >= 0; w--) {
    for (int i = 0; i < cfg.c; i++) acc = ge4_dbl(acc);
    acc = ge4_add(acc, ge4_load(src + (size_t)w * 32));
  }
  if (threadIdx.x < 4) ge4_store(out_ext + (size_t)set * 32, acc);
}

// out[set] = identity (X, Y, Z, T) = (0, 1, 1, 0)
__global__ void k_set_identity(uint32_t* __restrict__ out_ext) {
  out_ext[(size_t)blockIdx.x * 32 + threadIdx.x] = (threadIdx.x == 8 || threadIdx.x == 16) ? 1u : 0u;
}

// point ids of up to four consecutive ranges [off_i, off_i + len_i) of one table
struct SegIds {
  uint32_t off[4], len[4];
  int n;
};
__global__ void __launch_bounds__(256) k_seg_point_ids(SegIds sg, uint32_t total, uint32_t* __restrict__ ids) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  uint32_t r = t, id = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (i < sg.n) {
      if (r < sg.len[i]) {
        id = sg.off[i] + r;
        r = 0xffffffffu;
      } else if (r != 0xffffffffu) {
        r -= sg.len[i];
      }
    }
  }
  ids[t] = id;
}

// ---------------------------------------------------------------------------
// finishing: sum `nparts` partial sums per set (one per rank), encode
// ---------------------------------------------------------------------------
// parts layout: [part][set][32 words]
// One WARP per set: the sum of the parts is computed redundantly by its lanes, the encoding (one
// inverse square root, 252 dependent squarings) runs on the sixteen-lane field layer of fe16.cuh
// in its whole-warp form (the half-warps split every product).
constexpr int ENC_THREADS = 32;
__device__ __forceinline__ void store_s_bytes(uint8_t* out, const fe& s, uint32_t k) {
  uint32_t w = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) w = (k >> 1) == (uint32_t)i ? s.v[i] : w;
  w = (k & 1u) ? (w >> 16) : w;
  out[2 * k] = (uint8_t)w;
  out[2 * k + 1] = (uint8_t)(w >> 8);
}
__device__ __forceinline__ grp16 warp_group(uint32_t* sm_of_warp) {
  grp16 g;
  g.sm = sm_of_warp;
  g.k = threadIdx.x & 15u;
  g.half = (threadIdx.x >> 4) & 1u;
  g.par = 0;
  return g;
}
__global__ void __launch_bounds__(ENC_THREADS) k_sum_encode(const uint32_t* __restrict__ parts, int nparts, int nsets,
                                                             uint8_t* __restrict__ out_bytes /*nsets*32*/,
                                                             uint32_t* __restrict__ out_ext /*nsets*32 words, may be null*/) {
  __shared__ __align__(16) uint32_t sm[G16_WORDS];
  const uint32_t set = blockIdx.x;  // grid = nsets
  grp16 g = warp_group(sm);
  ge_ext acc;
  ge_load_ext(acc, parts + (size_t)set * 32);
  for (int p = 1; p < nparts; p++) {
    ge_ext o;
    ge_load_ext(o, parts + ((size_t)p * nsets + set) * 32);
    acc = ge_add(acc, o);
  }
  if (out_ext && threadIdx.x == 0) ge_store_ext(out_ext + (size_t)set * 32, acc);
  if (out_bytes) {
    fe s = ge_encode16<true>(g, acc);
    if (threadIdx.x < 16) store_s_bytes(out_bytes + (size_t)set * 32, s, g.k);
  }
}

// Accept-iff-identity (Verifier::verify, reference src/r1cs/verifier.rs:549): no encoding, hence no
// inverse square root.  A ristretto255 element equals the identity iff X = 0 or Y = 0 (RFC 9496
// §4.3.3: X1 Y2 == Y1 X2 or Y1 Y2 == X1 X2 against (0 : 1 : 1 : 0)).  out: 32 zero bytes (the
// identity's encoding) or 0x01 0x00.. (odd, so not a canonical encoding of anything).
__global__ void __launch_bounds__(32) k_sum_is_identity(const uint32_t* __restrict__ parts, int nparts,
                                                        uint8_t* __restrict__ out_bytes) {
  if (threadIdx.x != 0) return;
  ge_ext acc;
  ge_load_ext(acc, parts);
  for (int p = 1; p < nparts; p++) {
    ge_ext o;
    ge_load_ext(o, parts + (size_t)p * 32);
    acc = ge_add(acc, o);
  }
  bool id = fe_is_zero(acc.X) | fe_is_zero(acc.Y);
  uint32_t* w = reinterpret_cast<uint32_t*>(out_bytes);
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = 0;
  if (!id) w[0] = 1;
}

// ---------------------------------------------------------------------------
// Sharded MSM, the exchange step fused with the combine: ONE kernel per rank
//   1. stores this rank's partial sums (n_sets x 128 B) into slot [rank] of EVERY rank's exchange
//      buffer over NVLink (peer-mapped pointers, plain stores), fences system-wide and raises its
//      flag in every rank's flag array;
//   2. waits until all ranks' flags show this step's sequence number;
//   3. adds the `world` partials per set and encodes.
// The payload is 128 B per rank and set, so the cost is latency: this replaces an NCCL all-gather
// plus a separate combine launch.  Buffers are double-buffered by step parity: a rank can be at
// most one step ahead of the slowest one (it needs that rank's flag to finish a step).
// ---------------------------------------------------------------------------
// extended point from words written by a peer: volatile loads (never served from a stale L1 line)
__device__ __forceinline__ ge_ext ge_load_ext_volatile(const uint32_t* p) {
  const volatile uint32_t* src = p;
  ge_ext q;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    q.X.v[i] = src[i];
    q.Y.v[i] = src[8 + i];
    q.Z.v[i] = src[16 + i];
    q.T.v[i] = src[24 + i];
  }
  return q;
}
struct PeerPtrs {
  uint32_t* parts[8];  // rank p's parts buffer:  [2][world][max_sets][32] words
  uint32_t* flags[8];  // rank p's flags:          [2][world]
};
constexpr int XCH_THREADS = 256;
__global__ void __launch_bounds__(XCH_THREADS) k_exchange_sum_encode(const uint32_t* __restrict__ local_part, PeerPtrs peers,
                                                                      int world, int rank, int nsets, int max_sets,
                                                                      uint32_t seq, uint8_t* __restrict__ out_bytes,
                                                                      uint32_t* __restrict__ out_ext,
                                                                      uint32_t* __restrict__ status /*0 ok, 1 timeout*/) {
  const uint32_t slot = seq & 1u;
  const size_t slot_words = (size_t)world * max_sets * 32;
  // 1. push
  for (int p = 0; p < world; p++) {
    uint32_t* dst = peers.parts[p] + slot * slot_words + (size_t)rank * max_sets * 32;
    for (int w = threadIdx.x; w < nsets * 32; w += blockDim.x) dst[w] = local_part[w];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < (uint32_t)world) {
    volatile uint32_t* f = peers.flags[threadIdx.x] + slot * world + rank;
    *f = seq;
  }
  // 2. wait for every rank's flag (bounded: a dead peer must not hang the GPU)
  __shared__ uint32_t timed_out;
  if (threadIdx.x == 0) timed_out = 0;
  __syncthreads();
  if (threadIdx.x < (uint32_t)world) {
    volatile uint32_t* f = peers.flags[rank] + slot * world + threadIdx.x;
    uint32_t spins = 0;
    while (*f != seq) {
      __nanosleep(64);
      if (++spins > (1u << 24)) {  // > 1 s
        timed_out = 1;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  if (timed_out) {
    if (threadIdx.x == 0) *status = 1;
    return;
  }
  // 3. combine: one warp per set (the sum computed by each of its lanes, the encoding on the
  // whole-warp form of fe16.cuh); partials read past the L1 (they were written by peers)
  __shared__ __align__(16) uint32_t sm16[(XCH_THREADS / 32) * G16_WORDS];
  grp16 g = warp_group(sm16 + (threadIdx.x >> 5) * G16_WORDS);
  const uint32_t* base = peers.parts[rank] + slot * slot_words;
  for (int first = 0; first < nsets; first += XCH_THREADS / 32) {  // block-uniform trip count
    int set = first + (int)(threadIdx.x >> 5);
    if (set >= nsets) continue;  // whole warps drop out: the exchanges are warp-wide
    ge_ext acc = ge_load_ext_volatile(base + (size_t)set * 32);
    for (int p = 1; p < world; p++) acc = ge_add(acc, ge_load_ext_volatile(base + ((size_t)p * max_sets + set) * 32));
    if (out_ext && (threadIdx.x & 31) == 0) ge_store_ext(out_ext + (size_t)set * 32, acc);
    if (out_bytes) {
      fe s = ge_encode16<true>(g, acc);
      if ((threadIdx.x & 31) < 16) store_s_bytes(out_bytes + (size_t)set * 32, s, g.k);
    }
  }
}

// ---------------------------------------------------------------------------
// table construction: compressed ristretto -> affine Niels
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_decode_to_niels(const uint8_t* __restrict__ comp, uint32_t n,
                                                          uint32_t* __restrict__ table,
                                                          uint32_t* __restrict__ bad_count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t buf[32];
  const uint4* src = reinterpret_cast<const uint4*>(comp + (size_t)i * 32);
  uint4 a = src[0], b = src[1];
  uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int k = 0; k < 8; k++) {
    buf[4 * k] = (uint8_t)w[k];
    buf[4 * k + 1] = (uint8_t)(w[k] >> 8);
    buf[4 * k + 2] = (uint8_t)(w[k] >> 16);
    buf[4 * k + 3] = (uint8_t)(w[k] >> 24);
  }
  ge_ext p;
  bool ok = ge_decode(p, buf);
  ge_niels q;
  if (ok) {
    q = ge_affine_to_niels(p.X, p.Y);
  } else {
    q = ge_niels_identity();
    atomicAdd(bad_count, 1u);
  }
  ge_store_niels(table + (size_t)i * 24, q);
}

// extended -> compressed, one thread per point
__global__ void __launch_bounds__(128) k_encode(const uint32_t* __restrict__ ext, uint32_t n,
                                                 uint8_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge_ext p;
  ge_load_ext(p, ext + (size_t)i * 32);
  ge_encode(out + (size_t)i * 32, p);
}

// ---------------------------------------------------------------------------
// K-FIXED: fixed-base comb.  tab[j][d] = (d+1) * 16^j * P  (j < 64, d < 8), affine
// Niels, so k*P is 64 mixed additions and no doublings.  Serves the two-term
// Pedersen commitments `v*B + v_blinding*B_blinding` (reference
// src/generators.rs:41-43; prover.rs:325,627-631,687) and synthetic point sets.
// ---------------------------------------------------------------------------
constexpr int COMB_ENTRIES = COMB_WINDOWS * 8;

__global__ void __launch_bounds__(COMB_WINDOWS) k_comb_build(const uint8_t* __restrict__ base32,
                                                              uint32_t* __restrict__ table,
                                                              uint32_t* __restrict__ bad_count) {
  int j = threadIdx.x;
  uint8_t buf[32];
  for (int i = 0; i < 32; i++) buf[i] = base32[i];
  ge_ext p;
  if (!ge_decode(p, buf)) {
    if (j == 0) atomicAdd(bad_count, 1u);
    p = ge_identity();
  }
  for (int i = 0; i < 4 * j; i++) p = ge_dbl(p);
  ge_ext m = p;
  for (int d = 0; d < 8; d++) {
    ge_store_niels(table + (size_t)(j * 8 + d) * 24, ge_to_niels(m));
    m = ge_add(m, p);
  }
}

// out[i] = sum_t scalars[t*n + i] * base_t  for `nbases` comb tables laid out back to back
__global__ void __launch_bounds__(128) k_comb_mul(const uint32_t* __restrict__ tables, int nbases,
                                                   const uint32_t* __restrict__ scalars, uint32_t n, sc_bias bias4,
                                                   uint8_t* __restrict__ out_bytes,
                                                   uint32_t* __restrict__ out_ext) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge_ext acc = ge_identity();
  for (int t = 0; t < nbases; t++) {
    sc k;
    sc_load(k, scalars + ((size_t)t * n + i) * 8);
    sc_recoded r = sc_recode(k.v, bias4);
    const uint32_t* tab = tables + (size_t)t * COMB_ENTRIES * 24;
    for (int j = 0; j < COMB_WINDOWS; j++) {
      int d = sc_digit(r, j, 4);
      if (d != 0) {
        int mag = d < 0 ? -d : d;
        ge_niels q;
        ge_load_niels(q, tab + (size_t)(j * 8 + mag - 1) * 24);
        acc = ge_madd(acc, q, d < 0);
      }
    }
  }
  if (out_ext) ge_store_ext(out_ext + (size_t)i * 32, acc);
  if (out_bytes) ge_encode(out_bytes + (size_t)i * 32, acc);
}

// Few outputs (the five T_i of a proof, a V_j): one WARP per output.  The nbases*64 table lookups
// are spread over the lanes (a handful of mixed additions each), then a shuffle tree of five full
// additions; 128 dependent additions become ~4 + 5.
__global__ void __launch_bounds__(128) k_comb_mul_warp(const uint32_t* __restrict__ tables, int nbases,
                                                        const uint32_t* __restrict__ scalars, uint32_t n, sc_bias bias4,
                                                        uint8_t* __restrict__ out_bytes, uint32_t* __restrict__ out_ext) {
  uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t lane = threadIdx.x & 31;
  bool live = i < n;
  uint32_t ii = live ? i : n - 1;  // idle warps shadow the last output (whole warps, shuffles stay uniform)
  ge_ext acc = ge_identity();
  for (int t = 0; t < nbases; t++) {
    sc k;
    sc_load(k, scalars + ((size_t)t * n + ii) * 8);
    sc_recoded r = sc_recode(k.v, bias4);
    const uint32_t* tab = tables + (size_t)t * COMB_ENTRIES * 24;
    for (int j = lane; j < COMB_WINDOWS; j += 32) {
      int d = sc_digit(r, j, 4);
      if (d != 0) {
        int mag = d < 0 ? -d : d;
        ge_niels q;
        ge_load_niels(q, tab + (size_t)(j * 8 + mag - 1) * 24);
        acc = ge_madd(acc, q, d < 0);
      }
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    ge_ext o;
#pragma unroll
    for (int w = 0; w < 8; w++) {
      o.X.v[w] = __shfl_down_sync(0xffffffffu, acc.X.v[w], off);
      o.Y.v[w] = __shfl_down_sync(0xffffffffu, acc.Y.v[w], off);
      o.Z.v[w] = __shfl_down_sync(0xffffffffu, acc.Z.v[w], off);
      o.T.v[w] = __shfl_down_sync(0xffffffffu, acc.T.v[w], off);
    }
    acc = ge_add(acc, o);
  }
  if (live && lane == 0 && out_ext) ge_store_ext(out_ext + (size_t)i * 32, acc);
  if (out_bytes) {
    // the total sits in lane 0: hand it to every lane, encode on the whole warp (fe16.cuh)
    __shared__ __align__(16) uint32_t sm[(128 / 32) * G16_WORDS];
#pragma unroll
    for (int w = 0; w < 8; w++) {
      acc.X.v[w] = __shfl_sync(0xffffffffu, acc.X.v[w], 0);
      acc.Y.v[w] = __shfl_sync(0xffffffffu, acc.Y.v[w], 0);
      acc.Z.v[w] = __shfl_sync(0xffffffffu, acc.Z.v[w], 0);
      acc.T.v[w] = __shfl_sync(0xffffffffu, acc.T.v[w], 0);
    }
    grp16 g = warp_group(sm + (threadIdx.x >> 5) * G16_WORDS);
    fe s = ge_encode16<true>(g, acc);
    if (live && lane < 16) store_s_bytes(out_bytes + (size_t)i * 32, s, g.k);
  }
}

// ---------------------------------------------------------------------------
// Generator chains (reference src/generators.rs:107-125, 210-235): point i of a
// chain = element derivation (RFC 9496 §4.3.4) of the i-th 64-byte block of the chain's XOF
// stream.  The stream is squeezed on the host (sequential, ~1 GB/s); the two Elligator maps, the
// addition and the encoding (three inverse-square-root chains, ~900 field products per point)
// run here, one thread per point.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_from_uniform(const uint8_t* __restrict__ in /*n*64*/, uint32_t n,
                                                       uint8_t* __restrict__ out /*n*32*/) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge_ext p = ge_add(ge_elligator_map(fe_from_bytes_255(in + (size_t)i * 64)),
                    ge_elligator_map(fe_from_bytes_255(in + (size_t)i * 64 + 32)));
  ge_encode(out + (size_t)i * 32, p);
}

// ---------------------------------------------------------------------------
// windowed tables: out[w][i] = 2^(c w) * P_i in affine Niels, w < W.
// One thread per point walks the doubling chain, parks the extended multiples and
// the running product of their Z in scratch, inverts once (Montgomery's trick) and
// converts every multiple back to affine.  One-time cost at table upload; it removes
// all doublings from every later MSM over the table.
// ---------------------------------------------------------------------------
BPG_DEF_CONST(K_INV2, 0xfffffff7u, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu,
              0x3fffffffu)  // (p+1)/2

__global__ void __launch_bounds__(128) k_window_chain(const uint32_t* __restrict__ niels_in, uint32_t n_total,
                                                       uint32_t first, uint32_t count, int c, int W,
                                                       uint32_t* __restrict__ ext_scratch /*[W-1][count][32]*/,
                                                       uint32_t* __restrict__ zp_scratch /*[W-1][count][8]*/,
                                                       uint32_t* __restrict__ out /*[W][n_total][24]*/) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  uint32_t i = first + t;
  ge_niels q;
  ge_load_niels(q, niels_in + (size_t)i * 24);
  ge_store_niels(out + (size_t)i * 24, q);  // window 0
  ge_ext p;
  p.X = fe_sub(q.ypx, q.ymx);     // 2x
  p.Y = fe_add(q.ypx, q.ymx);     // 2y
  p.Z = fe_zero();
  p.Z.v[0] = 2;
  p.T = fe_mul(fe_mul(p.X, p.Y), fe_const(BPG_K(K_INV2)));  // XY/Z
  p.X = fe_mul(p.X, fe_one());    // tighten
  p.Y = fe_mul(p.Y, fe_one());
  fe zp = fe_one();
  for (int w = 1; w < W; w++) {
    for (int k = 0; k < c; k++) p = ge_dbl(p);
    zp = fe_mul(zp, p.Z);
    ge_store_ext(ext_scratch + ((size_t)(w - 1) * count + t) * 32, p);
    fe_store(zp_scratch + ((size_t)(w - 1) * count + t) * 8, zp);
  }
  fe inv = fe_invert(zp);
  for (int w = W - 1; w >= 1; w--) {
    ge_ext e;
    ge_load_ext(e, ext_scratch + ((size_t)(w - 1) * count + t) * 32);
    fe zi;
    if (w >= 2) {
      fe prev;
      fe_load(prev, zp_scratch + ((size_t)(w - 2) * count + t) * 8);
      zi = fe_mul(inv, prev);
    } else {
      zi = inv;
    }
    inv = fe_mul(inv, e.Z);
    fe x = fe_mul(e.X, zi), y = fe_mul(e.Y, zi);
    ge_store_niels(out + ((size_t)w * n_total + i) * 24, ge_affine_to_niels(x, y));
  }
}

}  // namespace bpg
