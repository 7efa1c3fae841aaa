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
static __global__ void __launch_bounds__(SORT_THREADS) k_hist(const uint32_t* __restrict__ scalars,
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

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t* __restrict__ counts, uint32_t B,
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
static __global__ void __launch_bounds__(1024) k_scan_spine(uint32_t* __restrict__ tile_sums, uint32_t ntiles,
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
static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(uint32_t* __restrict__ counts, uint32_t B,
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
static __global__ void __launch_bounds__(SORT_THREADS) k_scatter(const uint32_t* __restrict__ scalars,
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
// Two-pass radix sort of the (bucket, entry) pairs -- the default sort.  The counting sort above pays one L2
// atomic per entry in the histogram and one more in the scatter (13.6 M each at 2^20 points x 13 windows:
// ncu showed both kernels waiting on them, issue-active 14 %).  Here every per-entry atomic is a
// SHARED-memory atomic:
//   k_rs_hist     bucket id >> lb = partition; per-block histogram of the partitions, one global add per
//                 (block, occupied partition)
//   k_rs_scan     exclusive scan of the <= 8192 partition sizes
//   k_rs_scatter  per-block histogram again, ONE global reservation per (block, partition), then the pairs
//                 (bucket, entry) go to their partition's region at shared-memory ranks
//   k_rs_finish   one block per partition (2^lb buckets, a few thousand pairs): bucket histogram and cursors
//                 in shared memory -> offsets[], the size classes of the accumulation schedule, the entries in
//                 bucket order.  A partition of any size works (structured scalars: the block just loops).
// Order inside a bucket is irrelevant (the sum is a group element), so no stability is needed.
// ---------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_TERMS = 16;           // at most this many terms per thread of k_rs_hist / k_rs_scatter (tiles of 4096 terms)
constexpr uint32_t RS_MAX_PARTS = 8192;
constexpr uint32_t RS_MAX_LB = 12;     // at most 4096 buckets per partition
constexpr uint32_t RS_TARGET_PARTS = 2048;            // partitions aimed for (tuned at 2^20 points x 13 windows, tools/r2_sort_tune.sh)
constexpr size_t RS_STAGE_PAIRS = 12288;              // pairs of a tile staged in shared memory (96 KB)
constexpr size_t RS_SCATTER_SMEM = 200 * 1024;        // dynamic shared memory k_rs_scatter may be given

__device__ __forceinline__ uint32_t rs_bucket(const MsmCfg& cfg, uint32_t base, uint32_t g, int d) {
  uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
  return base + g * cfg.nb + mag - 1;
}

static __global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint32_t* __restrict__ scalars,
                                                               const uint8_t* __restrict__ set_ids, MsmCfg cfg, uint32_t lb,
                                                               uint32_t P, uint32_t* __restrict__ part_count,
                                                               uint32_t t_begin, uint32_t t_end, int tpt /*terms per thread*/) {
  extern __shared__ uint32_t rs_sm[];  // [P] histogram
  __shared__ uint32_t sh[9][SORT_THREADS];
  for (uint32_t p = threadIdx.x; p < P; p += RS_THREADS) rs_sm[p] = 0;
  __syncthreads();
  for (int k = 0; k < tpt; k++) {
    uint32_t t = t_begin + (blockIdx.x * tpt + k) * RS_THREADS + threadIdx.x;
    if (t >= t_end) break;  // later k are out of range as well
    sc s;
    sc_load(s, scalars + (size_t)t * 8);
    digits_park(sh, sc_recode(s.v, cfg.bias));
    uint32_t set = cfg.nsets > 1 ? (set_ids ? set_ids[t] : t / cfg.n_points) : 0;
    uint32_t base = set * cfg.gsub * cfg.nb;
    uint32_t g = 0;
    for (int w = 0; w < cfg.W; w++) {
      int d = digit_at(sh, w, cfg.c);
      if (d != 0) atomicAdd(&rs_sm[rs_bucket(cfg, base, g, d) >> lb], 1u);
      g = g + 1 == cfg.gsub ? 0 : g + 1;
    }
  }
  __syncthreads();
  for (uint32_t p = threadIdx.x; p < P; p += RS_THREADS)
    if (rs_sm[p]) atomicAdd(&part_count[p], rs_sm[p]);
}

// single block: part_base[0..P] = exclusive scan of part_count[0..P)
static __global__ void __launch_bounds__(1024) k_rs_scan(const uint32_t* __restrict__ part_count, uint32_t P,
                                                         uint32_t* __restrict__ part_base) {
  __shared__ uint32_t smem[33];
  const uint32_t per = (P + 1023) / 1024;  // <= 8
  uint32_t v[8];
  uint32_t s = 0;
  for (uint32_t k = 0; k < per; k++) {
    uint32_t i = threadIdx.x * per + k;
    v[k] = i < P ? part_count[i] : 0;
    s += v[k];
  }
  uint32_t total;
  uint32_t ex = block_exclusive_scan(s, &total, smem);
  for (uint32_t k = 0; k < per; k++) {
    uint32_t i = threadIdx.x * per + k;
    if (i < P) part_base[i] = ex;
    ex += v[k];
  }
  if (threadIdx.x == 0) part_base[P] = total;
}

// Tile of `tpt` x 256 terms: its pairs are counting-sorted by partition IN SHARED MEMORY, then every partition's run
// goes out in one piece (a run of r pairs is r x 8 contiguous bytes: sector-sized writes instead of one
// read-modify-write of a sector per pair).  Dynamic shared memory: pairs[cap] | hist[P] | soff[P] | gbase[P].
static __global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint32_t* __restrict__ scalars,
                                                                  const uint8_t* __restrict__ set_ids,
                                                                  const uint32_t* __restrict__ point_ids, MsmCfg cfg,
                                                                  uint32_t lb, uint32_t P,
                                                                  const uint32_t* __restrict__ part_base,
                                                                  uint32_t* __restrict__ part_cursor,
                                                                  uint2* __restrict__ pairs, int tpt, uint32_t cap) {
  extern __shared__ __align__(16) uint32_t rs_sm[];
  uint2* stage = reinterpret_cast<uint2*>(rs_sm);
  uint32_t* hist = rs_sm + 2 * (size_t)cap;
  uint32_t* soff = hist + P;
  uint32_t* gbase = soff + P;
  __shared__ uint32_t sh[9][SORT_THREADS];
  __shared__ uint32_t smem[33];
  __shared__ uint32_t carry_s;
  for (uint32_t p = threadIdx.x; p < P; p += RS_THREADS) hist[p] = 0;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int pass = 0; pass < 2; pass++) {
    for (int k = 0; k < tpt; k++) {
      uint32_t t = (blockIdx.x * tpt + k) * RS_THREADS + threadIdx.x;
      if (t >= cfg.n_terms) break;
      sc s;
      sc_load(s, scalars + (size_t)t * 8);
      digits_park(sh, sc_recode(s.v, cfg.bias));
      uint32_t set = cfg.nsets > 1 ? (set_ids ? set_ids[t] : t / cfg.n_points) : 0;
      uint32_t pid = point_ids ? point_ids[t] : t % cfg.n_points;
      uint32_t base = set * cfg.gsub * cfg.nb;
      uint32_t g = 0;
      for (int w = 0; w < cfg.W; w++) {
        int d = digit_at(sh, w, cfg.c);
        if (d != 0) {
          uint32_t b = rs_bucket(cfg, base, g, d);
          uint32_t r = atomicAdd(&hist[b >> lb], 1u);
          if (pass == 1)
            stage[soff[b >> lb] + r] = make_uint2(b, (pid + (uint32_t)w * cfg.win_stride) | (d < 0 ? ENTRY_NEG : 0u));
        }
        g = g + 1 == cfg.gsub ? 0 : g + 1;
      }
    }
    __syncthreads();
    if (pass == 0) {
      // exclusive scan of hist over the partitions -> soff; one global reservation per occupied partition
      for (uint32_t start = 0; start < P; start += RS_THREADS) {
        uint32_t p = start + threadIdx.x;
        uint32_t c = p < P ? hist[p] : 0;
        uint32_t total;
        uint32_t ex = block_exclusive_scan(c, &total, smem);
        uint32_t carry = carry_s;
        if (p < P) {
          soff[p] = carry + ex;
          gbase[p] = c ? part_base[p] + atomicAdd(&part_cursor[p], c) : 0;
          hist[p] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
      }
    }
  }
  const uint32_t npairs = carry_s;
  for (uint32_t i = threadIdx.x; i < npairs; i += RS_THREADS) {
    uint2 pr = stage[i];
    uint32_t p = pr.x >> lb;
    pairs[gbase[p] + (i - soff[p])] = pr;
  }
}

// one block per partition.  lens pass -> offsets + size classes -> placement pass.  Both passes keep four loads in
// flight per thread.  The placement happens IN SHARED MEMORY (the partition's entries in bucket order, `cap` of
// them; dynamic shared memory: 2^lb words + cap words) and goes out as one contiguous copy: 13.6 M scattered
// 4-byte stores were what bounded both the old scatter kernel and the first form of this one (one L2 sector
// transaction per entry: 358 us at 2^20 points x 13 windows).  Entries beyond `cap` (one huge partition:
// structured scalars) are stored directly.
constexpr uint32_t RS_FINISH_CAP = 8192;
constexpr uint32_t RS_FINISH_THREADS = 512;
static __global__ void __launch_bounds__(512) k_rs_finish(const uint2* __restrict__ pairs,
                                                                 const uint32_t* __restrict__ part_base, uint32_t B,
                                                                 uint32_t lb, uint32_t P, uint32_t* __restrict__ offsets,
                                                                 uint32_t* __restrict__ entries,
                                                                 uint32_t* __restrict__ bins /*[SIZE_BINS], zeroed*/,
                                                                 uint32_t cap) {
  extern __shared__ __align__(16) uint32_t rs_sm[];
  uint32_t* cnt = rs_sm;  // [2^lb] bucket counts, then cursors
  uint32_t* placed = rs_sm + (1u << lb);  // [cap]
  __shared__ uint32_t shb[SIZE_BINS];
  __shared__ uint32_t smem[33];
  const uint32_t p = blockIdx.x;
  const uint32_t beg = part_base[p], end = part_base[p + 1];
  const uint32_t first = p << lb;
  const uint32_t nbk = min(1u << lb, B - first);
  const uint32_t count = end - beg;
  for (uint32_t j = threadIdx.x; j < (1u << lb); j += blockDim.x) cnt[j] = 0;
  for (uint32_t j = threadIdx.x; j < SIZE_BINS; j += blockDim.x) shb[j] = 0;
  __syncthreads();
  for (uint32_t i0 = 0; i0 < count; i0 += 4 * blockDim.x) {
    uint32_t bk[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      uint32_t i = i0 + u * blockDim.x + threadIdx.x;
      if (i < count) bk[u] = pairs[beg + i].x;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      uint32_t i = i0 + u * blockDim.x + threadIdx.x;
      if (i < count) atomicAdd(&cnt[bk[u] - first], 1u);
    }
  }
  __syncthreads();
  // exclusive scan of cnt[0..2^lb): each thread owns a run of consecutive buckets
  const uint32_t per = ((1u << lb) + blockDim.x - 1) / blockDim.x;  // <= 16
  uint32_t v[16];
  uint32_t s = 0;
  for (uint32_t k = 0; k < per; k++) {
    uint32_t j = threadIdx.x * per + k;
    v[k] = j < nbk ? cnt[j] : 0;
    s += v[k];
  }
  uint32_t total;
  uint32_t ex = block_exclusive_scan(s, &total, smem);
  for (uint32_t k = 0; k < per; k++) {
    uint32_t j = threadIdx.x * per + k;
    if (j < nbk) {
      offsets[first + j] = beg + ex;
      cnt[j] = ex;  // becomes the bucket's cursor
      uint32_t len = v[k];
      if (len <= BIG_SEG) {
        uint32_t full = len / ACC_SEG, rem = len % ACC_SEG;
        if (full) atomicAdd(&shb[ACC_SEG], full);
        if (rem || !full) atomicAdd(&shb[rem], 1u);
      }
    }
    ex += v[k];
  }
  if (p == P - 1 && threadIdx.x == 0) offsets[B] = end;
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < SIZE_BINS; j += blockDim.x)
    if (shb[j]) atomicAdd(&bins[j], shb[j]);
  for (uint32_t i0 = 0; i0 < count; i0 += 4 * blockDim.x) {
    uint2 pr[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      uint32_t i = i0 + u * blockDim.x + threadIdx.x;
      if (i < count) pr[u] = pairs[beg + i];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      uint32_t i = i0 + u * blockDim.x + threadIdx.x;
      if (i < count) {
        uint32_t pos = atomicAdd(&cnt[pr[u].x - first], 1u);
        if (pos < cap) placed[pos] = pr[u].y;
        else entries[beg + pos] = pr[u].y;
      }
    }
  }
  __syncthreads();
  const uint32_t m = min(count, cap);
  for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) entries[beg + i] = placed[i];
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

static __global__ void __launch_bounds__(256) k_size_scatter(const uint32_t* __restrict__ offsets, MsmCfg cfg, AccSched sc,
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

}  // namespace bpg
