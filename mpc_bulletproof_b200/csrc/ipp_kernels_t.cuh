// Inner-product-argument round kernels, generic over the scalar field (policy SP: montmul, add,
// one_m, rr).  Same round structure as ipp_kernels.cuh (reference src/inner_product_proof.rs:49-193,
// 202-248: generators are never folded; each round is one two-output MSM over the original 2n
// generators with scalars a_{p^h} wG(i), b_{p^h} wH(i)); instantiated for the Stark-curve scalars
// (stark_sc.cuh).  a, b are kept in normal form, the weights wG, wH in Montgomery form, so every
// product montmul(normal, Montgomery) lands in normal form.
#pragma once
#include "stark_sc.cuh"

namespace bpg {

template <class SP>
__global__ void __launch_bounds__(256) k_ipp_init_weights_t(const uint32_t* __restrict__ g_factors,
                                                             const uint32_t* __restrict__ h_factors, uint32_t n,
                                                             uint32_t* __restrict__ wG, uint32_t* __restrict__ wH) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sc g = SP::one_m(), h = SP::one_m();
  if (g_factors) {
    sc_load(g, g_factors + (size_t)i * 8);
    g = SP::montmul(g, SP::rr());
  }
  if (h_factors) {
    sc_load(h, h_factors + (size_t)i * 8);
    h = SP::montmul(h, SP::rr());
  }
  sc_store(wG + (size_t)i * 8, g);
  sc_store(wH + (size_t)i * 8, h);
}

template <class SP>
__device__ __forceinline__ void block_sum2_t(sc& x, sc& y, uint32_t (*sm)[16]) {
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
      x = SP::add(x, ox);
      y = SP::add(y, oy);
    }
    __syncthreads();
  }
}

// c_L = <a_lo, b_hi>, c_R = <a_hi, b_lo> (:87-88, :156-157); partial sums carry a factor R^-1
template <class SP>
__global__ void __launch_bounds__(256) k_ipp_cross_t(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                      uint32_t h, uint32_t* __restrict__ partials /*[grid][16]*/) {
  __shared__ uint32_t sm[128][16];
  sc cl = sc_zero(), cr = sc_zero();
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < h; p += gridDim.x * blockDim.x) {
    sc alo, ahi, blo, bhi;
    sc_load(alo, a + (size_t)p * 8);
    sc_load(ahi, a + (size_t)(p + h) * 8);
    sc_load(blo, b + (size_t)p * 8);
    sc_load(bhi, b + (size_t)(p + h) * 8);
    cl = SP::add(cl, SP::montmul(alo, bhi));
    cr = SP::add(cr, SP::montmul(ahi, blo));
  }
  block_sum2_t<SP>(cl, cr, sm);
  if (threadIdx.x == 0) {
    sc_store(partials + (size_t)blockIdx.x * 16, cl);
    sc_store(partials + (size_t)blockIdx.x * 16 + 8, cr);
  }
}

// single block: finish the cross terms and append them as the Q terms (indices 2n, 2n+1) of the round's MSM
template <class SP>
__global__ void __launch_bounds__(256) k_ipp_cross_finish_t(const uint32_t* __restrict__ partials, uint32_t nparts,
                                                             uint32_t n, uint32_t* __restrict__ scalars,
                                                             uint8_t* __restrict__ set_ids) {
  __shared__ uint32_t sm[128][16];
  sc cl = sc_zero(), cr = sc_zero();
  for (uint32_t i = threadIdx.x; i < nparts; i += blockDim.x) {
    sc x, y;
    sc_load(x, partials + (size_t)i * 16);
    sc_load(y, partials + (size_t)i * 16 + 8);
    cl = SP::add(cl, x);
    cr = SP::add(cr, y);
  }
  block_sum2_t<SP>(cl, cr, sm);
  if (threadIdx.x == 0) {
    sc f = SP::rr();  // undo the R^-1 of the normal x normal products
    sc_store(scalars + (size_t)(2 * n) * 8, SP::montmul(cl, f));
    sc_store(scalars + (size_t)(2 * n + 1) * 8, SP::montmul(cr, f));
    set_ids[2 * n] = 0;
    set_ids[2 * n + 1] = 1;
  }
}

// the round's 2n generator scalars and their output set (0 = L, 1 = R)
template <class SP>
__global__ void __launch_bounds__(256) k_ipp_round_scalars_t(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                              const uint32_t* __restrict__ wG,
                                                              const uint32_t* __restrict__ wH, uint32_t n, uint32_t m,
                                                              uint32_t* __restrict__ scalars,
                                                              uint8_t* __restrict__ set_ids) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t h = m >> 1;
  uint32_t p = i & (m - 1);
  bool hi = (p & h) != 0;
  uint32_t partner = p ^ h;
  sc av, bv, g, hh;
  sc_load(av, a + (size_t)partner * 8);
  sc_load(bv, b + (size_t)partner * 8);
  sc_load(g, wG + (size_t)i * 8);
  sc_load(hh, wH + (size_t)i * 8);
  sc_store(scalars + (size_t)i * 8, SP::montmul(av, g));
  sc_store(scalars + (size_t)(n + i) * 8, SP::montmul(bv, hh));
  set_ids[i] = hi ? 0 : 1;
  set_ids[n + i] = hi ? 1 : 0;
}

// fold_witness (:202-248) for a, b; the generator fold becomes a weight update
template <class SP>
__global__ void __launch_bounds__(256) k_ipp_fold_t(uint32_t* __restrict__ a, uint32_t* __restrict__ b,
                                                     uint32_t* __restrict__ wG, uint32_t* __restrict__ wH, uint32_t n,
                                                     uint32_t m, ScPair u_pair /*u, u_inv: normal form, by value*/) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sc u, ui;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    u.v[w] = u_pair.v[w];
    ui.v[w] = u_pair.v[8 + w];
  }
  u = SP::montmul(u, SP::rr());
  ui = SP::montmul(ui, SP::rr());
  uint32_t h = m >> 1;
  uint32_t p = i & (m - 1);
  bool hi = (p & h) != 0;
  sc g, hh;
  sc_load(g, wG + (size_t)i * 8);
  sc_load(hh, wH + (size_t)i * 8);
  sc_store(wG + (size_t)i * 8, SP::montmul(g, hi ? u : ui));   // G' = u^-1 G_lo + u G_hi
  sc_store(wH + (size_t)i * 8, SP::montmul(hh, hi ? ui : u));  // H' = u H_lo + u^-1 H_hi
  if (i < h) {
    sc alo, ahi, blo, bhi;
    sc_load(alo, a + (size_t)i * 8);
    sc_load(ahi, a + (size_t)(i + h) * 8);
    sc_load(blo, b + (size_t)i * 8);
    sc_load(bhi, b + (size_t)(i + h) * 8);
    sc_store(a + (size_t)i * 8, SP::add(SP::montmul(alo, u), SP::montmul(ahi, ui)));  // a_lo*u + u^-1*a_hi
    sc_store(b + (size_t)i * 8, SP::add(SP::montmul(blo, ui), SP::montmul(bhi, u)));  // b_lo*u^-1 + u*b_hi
  }
}

}  // namespace bpg
