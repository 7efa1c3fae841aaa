// A few-term MSM over per-proof combs (the verifier's proof points, reference src/r1cs/verifier.rs:516-547 and
// src/inner_product_proof.rs:352-366): sum_k s_k P_k with the 64 x 8 cached multiples of every P_k built ahead
// (comb_from_points), so that what remains once the scalars exist is 64 additions per term and a tree -- the
// 253 doublings of a variable-base multiplication happened while the verifier was still replaying its transcript.
#pragma once
#include "comb_kernels.cuh"

namespace bpg {

struct CombMsm {
  const uint32_t* comb;     // [n][64][8][32] cached
  const uint32_t* scalars;  // [n][8] canonical
  uint32_t n, wsplit;
  sc_bias bias4;
  uint32_t* parts;          // [gridDim.x][32] ext
  uint32_t* ticket;         // zero before and after the launch
};
static __global__ void __launch_bounds__(CB_THREADS) k_comb_msm(CombMsm M, uint32_t* __restrict__ out_ext) {
  __shared__ __align__(16) uint32_t pts[CB_THREADS][32];
  __shared__ __align__(16) uint32_t sm[CB_THREADS / 32][32];
  __shared__ uint32_t s_last;
  const uint32_t u = blockIdx.x * CB_THREADS + threadIdx.x;
  const uint32_t k = u / M.wsplit, slice = u % M.wsplit;
  ge_ext acc = ge_identity();
  if (k < M.n) {
    sc v;
    sc_load(v, M.scalars + (size_t)k * 8);
    const sc_recoded r = sc_recode(v.v, M.bias4);
    const int per = COMB_WINDOWS / (int)M.wsplit;
    acc = comb_windows<false>(M.comb + (size_t)k * COMB_ENTRIES * COMB_CACHED_WORDS, r, (int)slice * per, (int)(slice + 1) * per);
  }
  ge4 tot = comb_block_sum(acc, pts, sm);
  if (gridDim.x == 1) {
    if (threadIdx.x < 4) ge4_store(out_ext, tot);
    return;
  }
  if (threadIdx.x < 4) ge4_store(M.parts + (size_t)blockIdx.x * 32, tot);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(M.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  // the last block to finish adds the blocks' sums (read around L1: other SMs wrote them)
  __threadfence();
  const uint32_t g = threadIdx.x >> 2, q = threadIdx.x & 3;
  ge4 t = ge4_identity();
  for (uint32_t base = 0; base < gridDim.x; base += CB_THREADS / 4) {  // block-uniform trip count
    const uint32_t i = base + g;
    ge4 o = ge4_identity();
    if (i < gridDim.x) {
      const uint4* src = reinterpret_cast<const uint4*>(M.parts + (size_t)i * 32 + 8 * q);
      const uint4 lo = __ldcg(src), hi = __ldcg(src + 1);
      o.c.v[0] = lo.x, o.c.v[1] = lo.y, o.c.v[2] = lo.z, o.c.v[3] = lo.w;
      o.c.v[4] = hi.x, o.c.v[5] = hi.y, o.c.v[6] = hi.z, o.c.v[7] = hi.w;
    }
    t = cb_add4(t, o);
  }
  tot = comb_tree_quads(t, sm);
  if (threadIdx.x < 4) ge4_store(out_ext, tot);
  if (threadIdx.x == 0) *M.ticket = 0;
}

// ---------------------------------------------------------------------------
// Indexed terms over the combs of a resident table, several output sets: the prover's A_I, A_O, S commitments
// (reference src/r1cs/prover.rs:465-494, 532-565) while the circuit is small.  A bucket-method MSM costs a dozen
// dependent launches whatever its size (0.36 ms for 5 x 1024 + 3 terms); 64 comb additions per term and two
// launches cost less up to a few thousand multipliers.  Set s owns the term `single[s]` and the range
// [lo[s], hi[s]) of the term arrays; grid.y = set, so a block's accumulators reduce to one partial sum.
// ---------------------------------------------------------------------------
struct CombTerms {
  const uint32_t* comb;       // affine combs of the table
  const uint32_t* scalars;    // [n_terms][8] canonical
  const uint32_t* point_ids;  // [n_terms]
  uint32_t single[4], lo[4], hi[4];
  uint32_t wsplit;
  sc_bias bias4;
};
static __global__ void __launch_bounds__(CB_THREADS) k_comb_terms(CombTerms M, uint32_t* __restrict__ parts /*[sets][gridDim.x][32]*/) {
  __shared__ __align__(16) uint32_t pts[CB_THREADS][32];
  __shared__ __align__(16) uint32_t sm[CB_THREADS / 32][32];
  const uint32_t set = blockIdx.y;
  const uint32_t units = 1 + M.hi[set] - M.lo[set];
  uint32_t* out = parts + ((size_t)set * gridDim.x + blockIdx.x) * 32;
  if (blockIdx.x * CB_THREADS >= units * M.wsplit) {  // block-uniform: nothing of this set left for this block
    if (threadIdx.x < 4) ge4_store(out, ge4_identity());
    return;
  }
  const uint32_t u = blockIdx.x * CB_THREADS + threadIdx.x;
  const uint32_t k = u / M.wsplit, slice = u % M.wsplit;
  ge_ext acc = ge_identity();
  if (k < units) {
    const uint32_t t = k == 0 ? M.single[set] : M.lo[set] + k - 1;
    sc v;
    sc_load(v, M.scalars + (size_t)t * 8);
    const sc_recoded r = sc_recode(v.v, M.bias4);
    const int per = COMB_WINDOWS / (int)M.wsplit;
    acc = comb_windows<true>(M.comb + (size_t)M.point_ids[t] * COMB_ENTRIES * COMB_AFFINE_WORDS, r, (int)slice * per,
                             (int)(slice + 1) * per);
  }
  ge4 tot = comb_block_sum(acc, pts, sm);
  if (threadIdx.x < 4) ge4_store(out, tot);
}
// the blocks' partial sums of every set -> the set's encoding (one block per set)
static __global__ void __launch_bounds__(CBQ_THREADS) k_parts_encode(const uint32_t* __restrict__ parts /*[sets][nparts][32]*/,
                                                              uint32_t nparts, uint8_t* __restrict__ out_bytes /*or null*/,
                                                              uint32_t* __restrict__ out_ext /*or null*/) {
  __shared__ __align__(16) uint32_t sm[CBQ_THREADS / 32][32];
  __shared__ __align__(16) uint32_t pt0[32];
  __shared__ __align__(16) uint32_t g16[G16_WORDS];
  const uint32_t set = blockIdx.x, g = threadIdx.x >> 2;
  ge4 acc = ge4_identity();
  for (uint32_t base = 0; base < nparts; base += CBQ_THREADS / 4) {  // block-uniform trip count
    const uint32_t i = base + g;
    ge4 o = i < nparts ? ge4_load(parts + ((size_t)set * nparts + i) * 32) : ge4_identity();
    acc = cb_add4(acc, o);
  }
  comb_tree_encode(acc, sm, pt0, g16, set, out_bytes, out_ext);
}

}  // namespace bpg
