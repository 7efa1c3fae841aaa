// Comb (fixed-base, doubling-free) multiscalar multiplication for the inner-product argument, sm_100a.
//
// `InnerProductProof::create` (reference src/inner_product_proof.rs:49-193) works on vectors that halve every
// round: its L/R MSMs have 2m + 1 terms in a round of length m, and it folds the generators with 2m
// two-term MSMs (:226-227).  A bucket-method MSM costs a fixed dozen dependent launches whatever its size,
// and the no-fold form of ipp_kernels.cuh costs 2n terms in every round.  This file removes both:
//
//  * A COMB of a point P holds (d+1) 16^j P for the 64 signed 4-bit windows j and d < 8, so that k P is 64
//    additions, no doublings, no buckets, no sort.  Resident generator tables carry one (affine Niels, 96 B
//    per entry, built once: k_table_comb_build).
//  * Rounds over at most a few thousand generators run as ONE accumulation kernel over (term, window) pairs
//    plus ONE finishing kernel (cross terms, the Q term, the tree over the blocks' partial sums, the two
//    encodings): k_comb_round / k_comb_final.  Their scalars a_partner * w(i) are formed on the fly, and so is
//    the fold of a, b and the generator weights that precedes the round (CombRound::fold: the reference's
//    fold_witness, :224-227, inside the kernel that forms L and R).
//  * For long vectors the first rounds keep the bucket method over the original generators (no-fold form);
//    when the vectors have shrunk to m0 entries the folded generators G'_p = sum_{i = p mod m0} w(i) G_i are
//    MATERIALISED once from the generator combs (k_comb_materialize: the reference's accumulated folds,
//    :125-146, as 2 m0 small MSMs), given combs of their own (k_comb_chain, k_comb_multiples: projective
//    "cached" entries, no inversion) and the remaining rounds run on m0 points.
// All of it computes the same group elements as the reference's round, hence the same encodings.
#pragma once
#include "ge.cuh"
#include "ge4.cuh"
#include "fe16.cuh"
#include "sc.cuh"

namespace bpg {

constexpr int COMB_AFFINE_WORDS = 24;  // (y+x, y-x, 2dxy), Z = 1
constexpr int COMB_CACHED_WORDS = 32;  // (Y-X, Y+X, 2Z, 2dT)

// Point operations of the latency-bound comb kernels.  A translation unit that defines BPG_GE_OUTLINE gets ONE
// out-of-line copy of each (products inline inside it): a lone warp runs hot looped code 1.3-1.7x faster than
// straight-line code it meets once, and an addition with inline products 20 % faster than one that calls each
// product (profiles/r2_lone_warp_latency.json), so the kernels below call these from every site.
#if defined(BPG_GE_OUTLINE) && defined(__CUDA_ARCH__)
#define BPG_CB_FN static __device__ __noinline__
#else
#define BPG_CB_FN __device__ __forceinline__
#endif
BPG_CB_FN ge_ext cb_madd(ge_ext p, ge_niels q, bool neg) { return ge_madd(p, q, neg); }
BPG_CB_FN ge_ext cb_add_cached(ge_ext p, fe ymx, fe ypx, fe z2, fe t2d) { return ge_add_cached(p, ymx, ypx, z2, t2d); }
BPG_CB_FN ge4 cb_add4(ge4 p, ge4 q) { return ge4_add(p, q); }
BPG_CB_FN ge4 cb_add4_cached(ge4 p, ge4 qc) { return ge4_add_cached(p, qc); }

// acc +- entry
template <bool AFFINE>
__device__ __forceinline__ ge_ext comb_add(const ge_ext& acc, const uint32_t* __restrict__ e, bool neg) {
  if (AFFINE) {
    ge_niels q;
    ge_load_niels(q, e);
    return ge_madd(acc, q, neg);
  } else {
    fe ymx, ypx, z2, t2d;
    fe_load(ymx, e);
    fe_load(ypx, e + 8);
    fe_load(z2, e + 16);
    fe_load(t2d, e + 24);
    fe nt = fe_neg(t2d);
    fe a = fe_sel(neg, ypx, ymx), b = fe_sel(neg, ymx, ypx), t = fe_sel(neg, nt, t2d);
    return ge_add_cached(acc, a, b, z2, t);
  }
}
// sum over windows [j0, j1) of digit_j * 16^j * P from P's comb.  Software-pipelined: the entry of window j + 1 is
// loaded before the addition of window j multiplies; a zero digit adds the neutral entry (the formulas are
// complete), so the loop has no data-dependent branch.
template <bool AFFINE>
struct comb_entry {
  fe c[AFFINE ? 3 : 4];
  bool neg;
};
template <bool AFFINE>
__device__ __forceinline__ comb_entry<AFFINE> comb_fetch(const uint32_t* __restrict__ comb_of_point, const sc_recoded& r, int j) {
  constexpr int WORDS = AFFINE ? COMB_AFFINE_WORDS : COMB_CACHED_WORDS;
  const int d = sc_digit(r, j, 4);
  const int mag = d < 0 ? -d : d;
  const uint32_t* e = comb_of_point + (size_t)(j * 8 + (mag ? mag - 1 : 0)) * WORDS;
  comb_entry<AFFINE> o;
  o.neg = d < 0;
#pragma unroll
  for (int k = 0; k < (AFFINE ? 3 : 4); k++) fe_load(o.c[k], e + 8 * k);
  if (mag == 0) {  // neutral element: affine (y+x, y-x, 2dxy) = (1, 1, 0); cached (Y-X, Y+X, 2Z, 2dT) = (1, 1, 2, 0)
    o.c[0] = fe_one();
    o.c[1] = fe_one();
    o.c[2] = fe_zero();
    if (!AFFINE) {
      o.c[2].v[0] = 2;
      o.c[3] = fe_zero();
    }
  }
  return o;
}
template <bool AFFINE>
__device__ __forceinline__ ge_ext comb_apply(const ge_ext& acc, const comb_entry<AFFINE>& q) {
  if (AFFINE) {
    ge_niels n;
    n.ypx = q.c[0];
    n.ymx = q.c[1];
    n.t2d = q.c[2];
    return cb_madd(acc, n, q.neg);
  } else {
    fe nt = fe_neg(q.c[3]);
    fe a = fe_sel(q.neg, q.c[1], q.c[0]), b = fe_sel(q.neg, q.c[0], q.c[1]), t = fe_sel(q.neg, nt, q.c[3]);
    return cb_add_cached(acc, a, b, q.c[2], t);
  }
}
template <bool AFFINE>
__device__ __forceinline__ ge_ext comb_windows(const uint32_t* __restrict__ comb_of_point, const sc_recoded& r, int j0, int j1,
                                               ge_ext acc = ge_identity()) {
  comb_entry<AFFINE> cur = comb_fetch<AFFINE>(comb_of_point, r, j0);
  for (int j = j0; j < j1; j++) {
    comb_entry<AFFINE> nxt = comb_fetch<AFFINE>(comb_of_point, r, j + 1 < j1 ? j + 1 : j);
    acc = comb_apply<AFFINE>(acc, cur);
    cur = nxt;
  }
  return acc;
}

// every thread of the block holds one extended point; the block's sum lands in quad 0 of warp 0 (all threads call)
constexpr int CB_THREADS = 128;
// sum of one point per quad over the whole block -> quad 0 of warp 0 (block_sum_quads of ge4.cuh through cb_add4,
// as loops: every level runs the same code).  Every thread of the block must call it.
__device__ __forceinline__ ge4 comb_tree_quads(ge4 p, uint32_t (*sm)[32]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll 1
  for (int off = 16; off >= 4; off >>= 1) {
    ge4 o;
#pragma unroll
    for (int i = 0; i < 8; i++) o.c.v[i] = __shfl_down_sync(BPG_FULL_MASK, p.c.v[i], off);
    p = cb_add4(p, o);
  }
  if (nw == 1) return p;
  if (lane < 4) ge4_store(sm[wid], p);
  __syncthreads();
  if (wid == 0) {
    const int quad = lane >> 2;
    ge4 t = quad < nw ? ge4_load(sm[quad]) : ge4_identity();
#pragma unroll 1
    for (int k = 1; k < (nw + 7) / 8; k++) {  // more than eight warps: quad q also folds warps q + 8, q + 16, ...
      const int w = quad + 8 * k;
      t = cb_add4(t, w < nw ? ge4_load(sm[w]) : ge4_identity());
    }
#pragma unroll 1
    for (int off = 16; off >= 4; off >>= 1) {
      ge4 o;
#pragma unroll
      for (int i = 0; i < 8; i++) o.c.v[i] = __shfl_down_sync(BPG_FULL_MASK, t.c.v[i], off);
      t = cb_add4(t, o);
    }
    p = t;
  }
  __syncthreads();
  return p;
}
__device__ __forceinline__ ge4 comb_block_sum(const ge_ext& mine, uint32_t (*pts)[32] /*[CB_THREADS][32]*/,
                                              uint32_t (*sm)[32] /*[CB_THREADS/32][32]*/) {
  ge_store_ext(pts[threadIdx.x], mine);
  __syncthreads();
  int g = threadIdx.x >> 2;
  ge4 t = ge4_load(pts[4 * g]);
#pragma unroll 1
  for (int k = 1; k < 4; k++) t = cb_add4(t, ge4_load(pts[4 * g + k]));
  return comb_tree_quads(t, sm);
}

// ---------------------------------------------------------------------------
// a round's accumulation.  grid.y = output set s = 2 lane + side (side 0: L, 1: R), grid.x covers the n terms
// of that set -- the G_i with bit h of i set (L) / clear (R) and the H_i with bit h clear (L) / set (R), exactly
// the pairing of inner_product_proof.rs:90-114, 159-172 -- times WSPLIT window slices.  A block is
// set-homogeneous, so its 128 accumulators reduce to ONE partial sum: parts[s][blockIdx.x].
// ---------------------------------------------------------------------------
struct IppPair {
  uint32_t v[16];  // u | u^-1, canonical words
};
struct CombRound {
  const uint32_t* comb;   // combs of the generators
  uint32_t g_id, h_id;    // comb index of G_0 and of H_0
  const uint32_t *a, *b;  // [lanes][stride] current vectors (normal form)
  const uint32_t *wG, *wH;  // [n] weights, Montgomery form (null: all ones)
  uint32_t stride;        // lane stride of a, b in scalars
  uint32_t n;             // generators per vector (weights length)
  uint32_t m;             // current vector length (power of two <= n)
  uint32_t wsplit;        // threads per term: each takes 64 / wsplit windows (power of two <= 64)
  sc_bias bias4;
  // The previous round's fold, taken on the fly (single-prover path): with `fold` set, a, b, wG, wH are the vectors
  // BEFORE the fold by the challenge `up` (a, b of length 2 m), every term forms its folded entries itself
  // (fold_witness, inner_product_proof.rs:224-227, 239-242, for the two entries it needs), and the folded state is
  // left in a2, b2, wG2, wH2 -- each entry by the one unit that computed it anyway.  The round is then ONE
  // accumulation kernel that folds a, b and the generator weights and forms L and R, plus the finish.
  uint32_t fold;
  IppPair up;
  uint32_t *a2, *b2, *wG2, *wH2;
};
template <bool AFFINE>
__global__ void __launch_bounds__(CB_THREADS) k_comb_round(CombRound R, uint32_t* __restrict__ parts /*[sets][gridDim.x][32]*/) {
  constexpr int WORDS = AFFINE ? COMB_AFFINE_WORDS : COMB_CACHED_WORDS;
  __shared__ __align__(16) uint32_t pts[CB_THREADS][32];
  __shared__ __align__(16) uint32_t sm[CB_THREADS / 32][32];
  const uint32_t set = blockIdx.y, lane_id = set >> 1, side = set & 1;
  const uint32_t u = blockIdx.x * CB_THREADS + threadIdx.x;
  const uint32_t k = u / R.wsplit, slice = u % R.wsplit;
  ge_ext acc = ge_identity();
  if (k < R.n) {
    const uint32_t half = R.n >> 1, h = R.m >> 1;
    const bool is_h = k >= half;
    const uint32_t kk = is_h ? k - half : k;
    // the kk-th index whose bit h is set / clear
    const bool want_bit = (side == 0) != is_h;  // L: G_hi, H_lo;  R: G_lo, H_hi
    uint32_t i = ((kk / h) * 2 * h) + (kk % h) + (want_bit ? h : 0);
    const uint32_t partner = (i & (R.m - 1)) ^ h;
    const uint32_t* vec = (is_h ? R.b : R.a) + (size_t)lane_id * R.stride * 8;
    const uint32_t* w = is_h ? R.wH : R.wG;
    sc v;
    if (R.fold) {
      sc u, ui;
#pragma unroll
      for (int t = 0; t < 8; t++) {
        u.v[t] = R.up.v[t];
        ui.v[t] = R.up.v[8 + t];
      }
      u = sc_to_mont(u);
      ui = sc_to_mont(ui);
      // a' = u a_lo + u^-1 a_hi,  b' = u^-1 b_lo + u b_hi  (entry `partner` of the folded vector)
      sc x0, x1;
      sc_load(x0, vec + (size_t)partner * 8);
      sc_load(x1, vec + (size_t)(partner + R.m) * 8);
      v = sc_add(sc_montmul(x0, is_h ? ui : u), sc_montmul(x1, is_h ? u : ui));
      // weight of generator i: times u for the half that is folded onto (G: upper, H: lower), u^-1 for the other
      const bool hi = (i & R.m) != 0;
      sc ww;
      sc_load(ww, w + (size_t)i * 8);
      ww = sc_montmul(ww, (hi != is_h) ? u : ui);
      if (slice == 0) {
        if (i < R.m) sc_store((is_h ? R.b2 : R.a2) + ((size_t)lane_id * R.stride + partner) * 8, v);
        if (lane_id == 0) sc_store((is_h ? R.wH2 : R.wG2) + (size_t)i * 8, ww);
      }
      v = sc_montmul(v, ww);
    } else {
      sc_load(v, vec + (size_t)partner * 8);
      if (w) {
        sc ww;
        sc_load(ww, w + (size_t)i * 8);
        v = sc_montmul(v, ww);
      }
    }
    const sc_recoded r = sc_recode(v.v, R.bias4);
    const int per = COMB_WINDOWS / (int)R.wsplit;
    const uint32_t id = (is_h ? R.h_id : R.g_id) + i;
    acc = comb_windows<AFFINE>(R.comb + (size_t)id * COMB_ENTRIES * WORDS, r, (int)slice * per, (int)(slice + 1) * per);
  }
  ge4 tot = comb_block_sum(acc, pts, sm);
  if (threadIdx.x < 4) ge4_store(parts + ((size_t)set * gridDim.x + blockIdx.x) * 32, tot);
}

// ---------------------------------------------------------------------------
// a round's finish, one block per output set: the cross term c_side (sum of the fold kernel's partials, or
// the caller's share of it), c * q_mul on Q's comb (one window per thread), the tree over the accumulation
// partials, the encoding.
// ---------------------------------------------------------------------------
struct CombFinal {
  const uint32_t* parts;       // [sets][nparts][32] ext
  uint32_t nparts;
  const uint32_t* cross;       // [ncross][16]: (c_L | c_R) partials, Montgomery-scaled by R^-1 (k_ipp_cross form); null if external
  uint32_t ncross;
  const uint32_t* c_ext;       // [lanes][2][8] canonical cross terms supplied by the caller (shares path); null otherwise
  const uint32_t *va, *vb;     // cross == null and c_ext == null: the current vectors (normal form), c_L = <a_lo, b_hi>,
  uint32_t vh;                 //   c_R = <a_hi, b_lo> over halves of length vh are formed here (inner_product_proof.rs:156-157)
  const uint32_t* q_mul;       // null or the scalar with Q = q_mul * (point of q_comb)
  const uint32_t* q_comb;      // comb of Q's base point
  sc_bias bias4;
};
// ---------------------------------------------------------------------------
// The finish in QUAD form: four adjacent lanes own one point, one coordinate each (ge4.cuh).  Adding a comb entry
// is two multiplication levels, the partial sums are loaded straight into quads and the block tree needs no
// regrouping through shared memory: the finish is one dependent chain on a single warp, and its length is what
// it costs (tools/lat_bench.cu: 1620 cycles per quad addition of a cached entry, 2350 per full quad addition,
// against 3530 / 4540 for a thread's mixed / full addition).  The ACCUMULATION stays thread-per-unit: at a few
// thousand terms it is bound by issue slots, and a thread's mixed addition costs 1.75 quad additions of work
// for 4 lanes' worth of registers (measured: the quad form of the accumulation was 20 % slower at every size).
// ---------------------------------------------------------------------------
// lane q of the quad gets its operand of the comb entry of window j in "cached" layout
// (lane0 Y-X, lane1 Y+X, lane2 2Z, lane3 2dT), negated for a negative digit, neutral for a zero digit
template <bool AFFINE>
__device__ __forceinline__ ge4 comb_fetch4(const uint32_t* __restrict__ comb_of_point, const sc_recoded& r, int j) {
  constexpr int WORDS = AFFINE ? COMB_AFFINE_WORDS : COMB_CACHED_WORDS;
  const int q = threadIdx.x & 3;
  const int d = sc_digit(r, j, 4);
  const int mag = d < 0 ? -d : d;
  const bool neg = d < 0;
  const uint32_t* e = comb_of_point + (size_t)(j * 8 + (mag ? mag - 1 : 0)) * WORDS;
  int comp;
  if (AFFINE) comp = q == 0 ? (neg ? 0 : 1) : (q == 1 ? (neg ? 1 : 0) : 2);  // words: y+x | y-x | 2dxy
  else comp = q == 0 ? (neg ? 1 : 0) : (q == 1 ? (neg ? 0 : 1) : q);         // words: Y-X | Y+X | 2Z | 2dT
  ge4 o;
  fe_load(o.c, e + 8 * comp);
  if (q == 3 && neg) o.c = fe_neg(o.c);
  fe two = fe_zero();
  two.v[0] = 2;
  if (AFFINE && q == 2) o.c = two;
  if (mag == 0) o.c = q < 2 ? fe_one() : (q == 2 ? two : fe_zero());
  return o;
}
template <bool AFFINE>
__device__ __forceinline__ ge4 comb_windows4(const uint32_t* __restrict__ comb_of_point, const sc_recoded& r, int j0, int j1, ge4 acc) {
  ge4 cur = comb_fetch4<AFFINE>(comb_of_point, r, j0);
  for (int j = j0; j < j1; j++) {
    ge4 nxt = comb_fetch4<AFFINE>(comb_of_point, r, j + 1 < j1 ? j + 1 : j);
    acc = cb_add4_cached(acc, cur);
    cur = nxt;
  }
  return acc;
}

constexpr int CBQ_THREADS = 256;  // 64 quads
// one point per quad over the block -> their sum, encoded by warp 0 (whole-warp sixteen-lane form) into
// out_bytes[set] when given (and, extended, into out_ext[set] when given).  Every thread of the block must call it.
__device__ __forceinline__ void comb_tree_encode(ge4 acc, uint32_t (*sm)[32], uint32_t* pt0 /*[32] shared*/,
                                                 uint32_t* g16 /*[G16_WORDS] shared*/, uint32_t set,
                                                 uint8_t* __restrict__ out_bytes, uint32_t* __restrict__ out_ext) {
  ge4 tot4 = comb_tree_quads(acc, sm);
  if (threadIdx.x < 4) ge4_store(pt0, tot4);
  __syncthreads();
  if (threadIdx.x < 32) {
    ge_ext tot;
    ge_load_ext(tot, pt0);
    if (out_ext && threadIdx.x == 0) ge_store_ext(out_ext + (size_t)set * 32, tot);
    if (!out_bytes) return;  // block-uniform: the caller wants the extended sum only
    grp16 gg;
    gg.sm = g16;
    gg.k = threadIdx.x & 15u;
    gg.half = (threadIdx.x >> 4) & 1u;
    gg.par = 0;
    fe s = ge_encode16<true>(gg, tot);
    if (threadIdx.x < 16) {
      uint32_t w = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) w = (gg.k >> 1) == (uint32_t)i ? s.v[i] : w;
      w = (gg.k & 1u) ? (w >> 16) : w;
      out_bytes[(size_t)set * 32 + 2 * gg.k] = (uint8_t)w;
      out_bytes[(size_t)set * 32 + 2 * gg.k + 1] = (uint8_t)(w >> 8);
    }
  }
}

template <bool Q_AFFINE>
__global__ void __launch_bounds__(CBQ_THREADS) k_comb_final(CombFinal F, uint8_t* __restrict__ out_bytes /*[sets][32]*/,
                                                             uint32_t* __restrict__ out_ext /*[sets][32] or null*/) {
  __shared__ __align__(16) uint32_t sm[CBQ_THREADS / 32][32];
  __shared__ __align__(16) uint32_t pt0[32];
  __shared__ __align__(16) uint32_t g16[G16_WORDS];
  __shared__ uint32_t csum[CBQ_THREADS / 32][8];
  const uint32_t set = blockIdx.x, lane_id = set >> 1, side = set & 1;
  // 1. the cross term of this side
  sc c = sc_zero();
  if (F.c_ext) {
    sc_load(c, F.c_ext + ((size_t)lane_id * 2 + side) * 8);
    if (F.q_mul) {
      sc q;
      sc_load(q, F.q_mul);
      c = sc_montmul(c, sc_to_mont(q));
    }
  } else {
    if (F.cross) {
      for (uint32_t i = threadIdx.x; i < F.ncross; i += CBQ_THREADS) {
        sc x;
        sc_load(x, F.cross + (size_t)i * 16 + 8 * side);
        c = sc_add(c, x);
      }
    } else {
      // side 0: sum_p a[p] b[p + h];  side 1: sum_p a[p + h] b[p]   (each product carries R^-1, like the partials)
      for (uint32_t p = threadIdx.x; p < F.vh; p += CBQ_THREADS) {
        sc x, y;
        sc_load(x, F.va + (size_t)(p + (side ? F.vh : 0)) * 8);
        sc_load(y, F.vb + (size_t)(p + (side ? 0 : F.vh)) * 8);
        c = sc_add(c, sc_montmul(x, y));
      }
    }
    // warp tree by shuffles, then the eight warp sums
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      sc o;
#pragma unroll
      for (int w = 0; w < 8; w++) o.v[w] = __shfl_down_sync(BPG_FULL_MASK, c.v[w], off);
      c = sc_add(c, o);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int w = 0; w < 8; w++) csum[threadIdx.x >> 5][w] = c.v[w];
    }
    __syncthreads();
    c = sc_zero();
#pragma unroll
    for (int k = 0; k < CBQ_THREADS / 32; k++) {
      sc x;
#pragma unroll
      for (int w = 0; w < 8; w++) x.v[w] = csum[k][w];
      c = sc_add(c, x);
    }
    sc f = sc_const(BPG_K(K_RR));  // the partial products carry R^-1
    if (F.q_mul) {
      sc q;
      sc_load(q, F.q_mul);
      f = sc_montmul(sc_to_mont(q), f);
    }
    c = sc_montmul(c, f);
  }
  // 2. c * Q: quad g contributes window g; every quad also folds in its share of the partial sums
  const uint32_t g = threadIdx.x >> 2;
  const sc_recoded r = sc_recode(c.v, F.bias4);
  ge4 acc = comb_windows4<Q_AFFINE>(F.q_comb, r, (int)g, (int)g + 1, ge4_identity());
  for (uint32_t base = 0; base < F.nparts; base += CBQ_THREADS / 4) {  // warp-uniform trip count (whole-warp shuffles inside)
    const uint32_t i = base + g;
    ge4 o = i < F.nparts ? ge4_load(F.parts + ((size_t)set * F.nparts + i) * 32) : ge4_identity();
    acc = cb_add4(acc, o);
  }
  comb_tree_encode(acc, sm, pt0, g16, set, out_bytes, out_ext);
}

// ---------------------------------------------------------------------------
// fold_witness for a, b (inner_product_proof.rs:224-225, 239-240) fused with the NEXT round's cross terms
// (:156-157): thread j < h' folds the four entries j, j + h' of both vectors (h' = half of the folded length)
// and adds its two products to the block's partial sums; the weights pick up u^(+-1) as in k_ipp_fold.
// grid.y = lane; cross terms only where `partials` is given (shares: they come from the fabric).
// ---------------------------------------------------------------------------
constexpr int IFC_THREADS = 256;
static __global__ void __launch_bounds__(IFC_THREADS) k_ipp_fold_cross(uint32_t* __restrict__ a, uint32_t* __restrict__ b,
                                                                uint32_t* __restrict__ wG, uint32_t* __restrict__ wH,
                                                                uint32_t n /*weights*/, uint32_t m /*length before the fold*/,
                                                                uint32_t stride, IppPair up,
                                                                uint32_t* __restrict__ partials /*[gridDim.x][16] or null*/) {
  __shared__ uint32_t sm[IFC_THREADS / 2][16];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane_id = blockIdx.y;
  a += (size_t)lane_id * stride * 8;
  b += (size_t)lane_id * stride * 8;
  sc u, ui;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    u.v[w] = up.v[w];
    ui.v[w] = up.v[8 + w];
  }
  u = sc_to_mont(u);
  ui = sc_to_mont(ui);
  const uint32_t mn = m >> 1, hn = mn >> 1;  // folded length and its half
  if (lane_id == 0 && i < n && wG) {
    const bool hi = (i & mn) != 0;
    sc g, hh;
    sc_load(g, wG + (size_t)i * 8);
    sc_load(hh, wH + (size_t)i * 8);
    sc_store(wG + (size_t)i * 8, sc_montmul(g, hi ? u : ui));
    sc_store(wH + (size_t)i * 8, sc_montmul(hh, hi ? ui : u));
  }
  sc cl = sc_zero(), cr = sc_zero();
  const uint32_t cnt = hn ? hn : 1;  // folded length 1: a single entry, no cross terms
  if (i < cnt) {
    sc x0, x1, y0, y1;
    sc_load(x0, a + (size_t)i * 8);
    sc_load(x1, a + (size_t)(i + mn) * 8);
    sc_load(y0, b + (size_t)i * 8);
    sc_load(y1, b + (size_t)(i + mn) * 8);
    sc alo = sc_add(sc_montmul(x0, u), sc_montmul(x1, ui));
    sc blo = sc_add(sc_montmul(y0, ui), sc_montmul(y1, u));
    if (hn) {
      sc x2, x3, y2, y3;
      sc_load(x2, a + (size_t)(i + hn) * 8);
      sc_load(x3, a + (size_t)(i + hn + mn) * 8);
      sc_load(y2, b + (size_t)(i + hn) * 8);
      sc_load(y3, b + (size_t)(i + hn + mn) * 8);
      sc ahi = sc_add(sc_montmul(x2, u), sc_montmul(x3, ui));
      sc bhi = sc_add(sc_montmul(y2, ui), sc_montmul(y3, u));
      sc_store(a + (size_t)(i + hn) * 8, ahi);
      sc_store(b + (size_t)(i + hn) * 8, bhi);
      cl = sc_montmul(alo, bhi);
      cr = sc_montmul(ahi, blo);
    }
    sc_store(a + (size_t)i * 8, alo);
    sc_store(b + (size_t)i * 8, blo);
  }
  if (!partials) return;
  // block sums (two accumulators at once)
  for (int half = IFC_THREADS / 2; half >= 1; half >>= 1) {
    if (threadIdx.x >= (uint32_t)half && threadIdx.x < (uint32_t)(2 * half)) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        sm[threadIdx.x - half][k] = cl.v[k];
        sm[threadIdx.x - half][8 + k] = cr.v[k];
      }
    }
    __syncthreads();
    if (threadIdx.x < (uint32_t)half) {
      sc ox, oy;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        ox.v[k] = sm[threadIdx.x][k];
        oy.v[k] = sm[threadIdx.x][8 + k];
      }
      cl = sc_add(cl, ox);
      cr = sc_add(cr, oy);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && lane_id == 0) {
    sc_store(partials + (size_t)blockIdx.x * 16, cl);
    sc_store(partials + (size_t)blockIdx.x * 16 + 8, cr);
  }
}

}  // namespace bpg
