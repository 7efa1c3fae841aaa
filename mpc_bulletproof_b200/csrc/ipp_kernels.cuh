// Inner-product-argument rounds, sm_100a.
//
// Stands behind `InnerProductProof::create` / `fold_witness`
// (reference src/inner_product_proof.rs:49-193, 202-248).
//
// The reference folds the generator vectors every round with 2h two-term scalar
// multiplications (G' = u^-1 G_lo + u G_hi, H' = u H_lo + u^-1 H_hi, :226-227) and,
// in round 0, multiplies all 2n generators by their factors (:125-134).  Here the
// generators are never touched: after j rounds the folded generator at position p is
//     G^(j)_p = sum_{i = p mod m} wG_j(i) * G_i,   wG_j(i) = g_i * prod_{r<j} u_r^{+-1}
// (the sign chosen by bit k-1-r of i, exactly the `s` vector of :300-307), so the
// round's cross terms
//     L = <a_lo, G_hi> + <b_hi, H_lo> + c_L Q,   R = <a_hi, G_lo> + <b_lo, H_hi> + c_R Q
// are ONE two-output MSM over the original 2n generators with scalars
// a_{p^h} * wG(i) and b_{p^h} * wH(i): every generator contributes to exactly one
// of L, R.  Same group elements as the reference's, hence the same encodings;
// the table stays in affine-Niels form with its window multiples precomputed, and
// only scalar vectors (a, b, wG, wH) are updated between rounds.
#pragma once
#include "sc.cuh"

namespace bpg {

// wG/wH are kept in Montgomery form so that montmul(normal, mont) lands in normal form.
static __global__ void __launch_bounds__(256) k_ipp_init_weights(const uint32_t* __restrict__ g_factors,
                                                           const uint32_t* __restrict__ h_factors, uint32_t n,
                                                           uint32_t* __restrict__ wG, uint32_t* __restrict__ wH) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sc one_m = sc_const(BPG_K(K_R1));
  sc g = one_m, h = one_m;
  if (g_factors) {
    sc_load(g, g_factors + (size_t)i * 8);
    g = sc_to_mont(g);
  }
  if (h_factors) {
    sc_load(h, h_factors + (size_t)i * 8);
    h = sc_to_mont(h);
  }
  sc_store(wG + (size_t)i * 8, g);
  sc_store(wH + (size_t)i * 8, h);
}

// block-wide sum of scalars mod l (two accumulators at once)
__device__ __forceinline__ void block_sum2(sc& x, sc& y, uint32_t (*sm)[16]) {
  // sm: [blockDim.x/2][16]
  for (int half = blockDim.x / 2; half >= 1; half >>= 1) {
    if (threadIdx.x >= (uint32_t)half && threadIdx.x < (uint32_t)(2 * half)) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        sm[threadIdx.x - half][k] = x.v[k];
        sm[threadIdx.x - half][8 + k] = y.v[k];
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
      x = sc_add(x, ox);
      y = sc_add(y, oy);
    }
    __syncthreads();
  }
}

// partial cross terms: c_L = <a_lo, b_hi>, c_R = <a_hi, b_lo> (:87-88, :156-157); Montgomery-scaled
constexpr int IPP_THREADS = 256;
static __global__ void __launch_bounds__(IPP_THREADS) k_ipp_cross(const uint32_t* __restrict__ a,
                                                            const uint32_t* __restrict__ b, uint32_t h,
                                                            uint32_t* __restrict__ partials /*[grid][16]*/) {
  __shared__ uint32_t sm[IPP_THREADS / 2][16];
  sc cl = sc_zero(), cr = sc_zero();
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < h; p += gridDim.x * blockDim.x) {
    sc alo, ahi, blo, bhi;
    sc_load(alo, a + (size_t)p * 8);
    sc_load(ahi, a + (size_t)(p + h) * 8);
    sc_load(blo, b + (size_t)p * 8);
    sc_load(bhi, b + (size_t)(p + h) * 8);
    cl = sc_add(cl, sc_montmul(alo, bhi));
    cr = sc_add(cr, sc_montmul(ahi, blo));
  }
  block_sum2(cl, cr, sm);
  if (threadIdx.x == 0) {
    sc_store(partials + (size_t)blockIdx.x * 16, cl);
    sc_store(partials + (size_t)blockIdx.x * 16 + 8, cr);
  }
}

// single block: finish the cross terms and append them as the Q terms of the round's MSM
static __global__ void __launch_bounds__(IPP_THREADS) k_ipp_cross_finish(const uint32_t* __restrict__ partials,
                                                                   uint32_t nparts, uint32_t n,
                                                                   const uint32_t* __restrict__ q_mul /*null or scalar*/,
                                                                   uint32_t* __restrict__ scalars,
                                                                   uint8_t* __restrict__ set_ids,
                                                                   uint32_t* __restrict__ q_side /*null, or c_L | c_R go here*/) {
  __shared__ uint32_t sm[IPP_THREADS / 2][16];
  sc cl = sc_zero(), cr = sc_zero();
  for (uint32_t i = threadIdx.x; i < nparts; i += blockDim.x) {
    sc x, y;
    sc_load(x, partials + (size_t)i * 16);
    sc_load(y, partials + (size_t)i * 16 + 8);
    cl = sc_add(cl, x);
    cr = sc_add(cr, y);
  }
  block_sum2(cl, cr, sm);
  if (threadIdx.x == 0) {
    // cl, cr carry a factor R^-1; multiply by R^2 (and by q_mul when Q = q_mul * base point)
    sc f = sc_const(BPG_K(K_RR));
    if (q_mul) {
      sc q;
      sc_load(q, q_mul);
      f = sc_montmul(sc_to_mont(q), f);  // q * R^2
    }
    sc vl = sc_montmul(cl, f), vr = sc_montmul(cr, f);
    if (q_side) {
      // Q is not in the table: c_L Q, c_R Q are formed beside the MSM (fixed-base comb of Q)
      sc_store(q_side, vl);
      sc_store(q_side + 8, vr);
      vl = sc_zero();
      vr = sc_zero();
    }
    sc_store(scalars + (size_t)(2 * n) * 8, vl);      // c_L * Q -> L
    sc_store(scalars + (size_t)(2 * n + 1) * 8, vr);  // c_R * Q -> R
    set_ids[2 * n] = 0;
    set_ids[2 * n + 1] = 1;
  }
}

// the round's 2n generator scalars and their output set (0 = L, 1 = R)
static __global__ void __launch_bounds__(256) k_ipp_round_scalars(const uint32_t* __restrict__ a,
                                                            const uint32_t* __restrict__ b,
                                                            const uint32_t* __restrict__ wG,
                                                            const uint32_t* __restrict__ wH, uint32_t n, uint32_t m,
                                                            uint32_t* __restrict__ scalars,
                                                            uint8_t* __restrict__ set_ids) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // lane = blockIdx.y: a party's share vectors and their MAC vectors (r1cs_mpc) are further (a, b) pairs that
  // meet the same generator weights; lane k's sums are output sets 2k (L) and 2k + 1 (R)
  const uint32_t lane = blockIdx.y;
  const size_t T = 2 * (size_t)n + 2;
  a += (size_t)lane * n * 8;
  b += (size_t)lane * n * 8;
  scalars += lane * T * 8;
  set_ids += lane * T;
  uint32_t h = m >> 1;
  uint32_t p = i & (m - 1);
  bool hi = (p & h) != 0;
  uint32_t partner = p ^ h;
  sc av, bv, g, hh;
  sc_load(av, a + (size_t)partner * 8);
  sc_load(bv, b + (size_t)partner * 8);
  sc_load(g, wG + (size_t)i * 8);
  sc_load(hh, wH + (size_t)i * 8);
  sc_store(scalars + (size_t)i * 8, sc_montmul(av, g));         // G_i: a_lo with G_hi -> L, a_hi with G_lo -> R
  sc_store(scalars + (size_t)(n + i) * 8, sc_montmul(bv, hh));  // H_i: b_hi with H_lo -> L, b_lo with H_hi -> R
  set_ids[i] = (uint8_t)(2 * lane + (hi ? 0 : 1));
  set_ids[n + i] = (uint8_t)(2 * lane + (hi ? 1 : 0));
}

// Cross terms supplied by the caller (r1cs_mpc: c_L, c_R are products of SHARED vectors, so they come out of
// the fabric's multiplication protocol, reference src/r1cs_mpc/mpc_inner_product.rs:104-105, 172-173):
// c[lane][0..1] canonical scalars -> the two Q terms of every lane, times q_mul when Q = q_mul * base point.
static __global__ void k_ipp_q_terms_ext(const uint32_t* __restrict__ c /*[lanes][2][8]*/, const uint32_t* __restrict__ q_mul,
                                         uint32_t n, uint32_t lanes, uint32_t* __restrict__ scalars,
                                         uint8_t* __restrict__ set_ids) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * lanes) return;
  uint32_t lane = t >> 1, side = t & 1;
  const size_t T = 2 * (size_t)n + 2;
  sc v;
  sc_load(v, c + (size_t)t * 8);
  if (q_mul) {
    sc q;
    sc_load(q, q_mul);
    v = sc_montmul(v, sc_to_mont(q));
  }
  sc_store(scalars + (lane * T + 2 * (size_t)n + side) * 8, v);
  set_ids[lane * T + 2 * (size_t)n + side] = (uint8_t)(2 * lane + side);
}

// the round's challenge and its inverse (canonical words), passed by value
struct ScPair {
  uint32_t v[16];
};

// fold_witness (:202-248) for a, b; the generator fold becomes a weight update
static __global__ void __launch_bounds__(256) k_ipp_fold(uint32_t* __restrict__ a, uint32_t* __restrict__ b,
                                                   uint32_t* __restrict__ wG, uint32_t* __restrict__ wH, uint32_t n,
                                                   uint32_t m, ScPair u_pair /*u, u_inv: kernel arguments, no copy*/) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t lane = blockIdx.y;  // further (a, b) pairs folding with the same public challenge (r1cs_mpc lanes)
  a += (size_t)lane * n * 8;
  b += (size_t)lane * n * 8;
  sc u, ui;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    u.v[w] = u_pair.v[w];
    ui.v[w] = u_pair.v[8 + w];
  }
  u = sc_to_mont(u);
  ui = sc_to_mont(ui);
  uint32_t h = m >> 1;
  uint32_t p = i & (m - 1);
  bool hi = (p & h) != 0;
  if (lane == 0) {
    sc g, hh;
    sc_load(g, wG + (size_t)i * 8);
    sc_load(hh, wH + (size_t)i * 8);
    sc_store(wG + (size_t)i * 8, sc_montmul(g, hi ? u : ui));   // G' = u^-1 G_lo + u G_hi
    sc_store(wH + (size_t)i * 8, sc_montmul(hh, hi ? ui : u));  // H' = u H_lo + u^-1 H_hi
  }
  if (i < h) {
    sc alo, ahi, blo, bhi;
    sc_load(alo, a + (size_t)i * 8);
    sc_load(ahi, a + (size_t)(i + h) * 8);
    sc_load(blo, b + (size_t)i * 8);
    sc_load(bhi, b + (size_t)(i + h) * 8);
    sc_store(a + (size_t)i * 8, sc_add(sc_montmul(alo, u), sc_montmul(ahi, ui)));  // a_lo*u + u^-1*a_hi
    sc_store(b + (size_t)i * 8, sc_add(sc_montmul(blo, ui), sc_montmul(bhi, u)));  // b_lo*u^-1 + u*b_hi
  }
}

// point ids of the round MSM's 2n+2 terms: G_i, H_i, Q, Q
static __global__ void k_ipp_point_ids(uint32_t* out, uint32_t n, uint32_t g_base, uint32_t h_base, uint32_t q_id) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  out += (size_t)blockIdx.y * (2 * (size_t)n + 2);  // one copy per lane
  if (i < n) {
    out[i] = g_base + i;
    out[n + i] = h_base + i;
  }
  if (i < 2) out[2 * n + i] = q_id;
}

}  // namespace bpg
