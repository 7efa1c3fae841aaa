// Pippenger pipeline, bucket reduction, the few-term path and Horner (see msm_sort_kernels.cuh).
#pragma once
#include "msm_sort_kernels.cuh"

namespace bpg {

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
static __global__ void __launch_bounds__(RP_THREADS) k_reduce_pairs(const uint32_t* __restrict__ in_a,
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
static __global__ void __launch_bounds__(RPB_THREADS) k_reduce_pairs_final(const uint32_t* __restrict__ in_a,
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

static __global__ void __launch_bounds__(SMALL_THREADS) k_msm_small(const uint32_t* __restrict__ table /*affine Niels*/,
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
  ge_load_niels(nq, table + (size_t)pid * NIELS_WORDS);
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
static __global__ void __launch_bounds__(SMALL_THREADS) k_msm_small_fin(const uint32_t* __restrict__ parts, uint32_t nblocks,
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
static __global__ void __launch_bounds__(32) k_horner(const uint32_t* __restrict__ window_sums, MsmCfg cfg,
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
static __global__ void k_set_identity(uint32_t* __restrict__ out_ext) {
  out_ext[(size_t)blockIdx.x * 32 + threadIdx.x] = (threadIdx.x == 8 || threadIdx.x == 16) ? 1u : 0u;
}

}  // namespace bpg

