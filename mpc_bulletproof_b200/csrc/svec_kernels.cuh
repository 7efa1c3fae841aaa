// Scalar-vector kernels: the O(n) scalar preparation around the R1CS MSMs, on the device
// so that n-vectors never cross PCIe more than once.
//
//   prover  (reference src/r1cs/prover.rs:589-619, 650-697; src/util.rs:152-181):
//     l/r polynomial coefficients, the six t_i inner products, l(x), r(x) with padding,
//     G/H factors — from a_L, a_R, a_O, s_L, s_R and the flattened weights wL, wR, wO.
//   verifier (reference src/r1cs/verifier.rs:468-501; src/inner_product_proof.rs:283-307):
//     y^-i, the s vector, delta, g_scalars, h_scalars.
//
// All vectors handled here are in Montgomery form (R = 2^256), which is also the host
// mirror's in-memory form (4x64 little-endian limbs == 8x32), so uploads are raw copies;
// values that feed an MSM or leave the device are converted to canonical form last.
#pragma once
#include "sc.cuh"

namespace bpg {

struct PowTable {  // base^(2^k), k < 32, Montgomery form
  uint32_t v[32][8];
};

__device__ __forceinline__ sc sc_pow(const PowTable& t, uint32_t e) {
  sc r = sc_const(BPG_K(K_R1));
  for (int k = 0; k < 32; k++) {
    if ((e >> k) == 0) break;
    if ((e >> k) & 1u) {
      sc b;
#pragma unroll
      for (int j = 0; j < 8; j++) b.v[j] = t.v[k][j];
      r = sc_montmul(r, b);
    }
  }
  return r;
}
__device__ __forceinline__ sc sc_param(const uint32_t* p) {
  sc r;
#pragma unroll
  for (int j = 0; j < 8; j++) r.v[j] = p[j];
  return r;
}

constexpr int SV_THREADS = 256;

// block sum of K scalars per thread -> thread 0 holds the totals
template <int K>
__device__ __forceinline__ void block_sum_k(sc (&x)[K], uint32_t (*sm)[8 * K]) {
  for (int half = blockDim.x / 2; half >= 1; half >>= 1) {
    if (threadIdx.x >= (uint32_t)half && threadIdx.x < (uint32_t)(2 * half)) {
#pragma unroll
      for (int k = 0; k < K; k++)
#pragma unroll
        for (int j = 0; j < 8; j++) sm[threadIdx.x - half][8 * k + j] = x[k].v[j];
    }
    __syncthreads();
    if (threadIdx.x < (uint32_t)half) {
#pragma unroll
      for (int k = 0; k < K; k++) {
        sc o;
#pragma unroll
        for (int j = 0; j < 8; j++) o.v[j] = sm[threadIdx.x][8 * k + j];
        x[k] = sc_add(x[k], o);
      }
    }
    __syncthreads();
  }
}

// The two blinding vectors of a phase (s_L, s_R; prover.rs:460-461, 523-528), generated where
// they are consumed: consecutive 64-byte blocks of the SplitMix64 stream keyed by one draw of
// the prover's PRNG are s_L[0], s_R[0], s_L[1], ...; each block is reduced mod l as
// lo*R + hi*R^2 (Montgomery form of lo + hi 2^256).  Same definition as oracle/protocol.py
// Blindings.vector_pair.
__device__ __forceinline__ unsigned long long splitmix_word(unsigned long long key, unsigned long long counter) {
  unsigned long long z = key + 0x9E3779B97F4A7C15ULL * (counter + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ sc sc_from_stream_block(unsigned long long key, unsigned long long block) {
  sc lo, hi;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    unsigned long long a = splitmix_word(key, 8 * block + k), b = splitmix_word(key, 8 * block + 4 + k);
    lo.v[2 * k] = (uint32_t)a;
    lo.v[2 * k + 1] = (uint32_t)(a >> 32);
    hi.v[2 * k] = (uint32_t)b;
    hi.v[2 * k + 1] = (uint32_t)(b >> 32);
  }
  sc rr = sc_const(BPG_K(K_RR));
  // montmul accepts any 256-bit left operand against rr < l: (x * R^2)/R = x R
  sc lo_m = sc_montmul(lo, rr);
  sc hi_m = sc_montmul(sc_montmul(hi, rr), rr);  // hi R -> hi R^2 = Montgomery form of hi*R
  return sc_add(lo_m, hi_m);
}
static __global__ void __launch_bounds__(256) k_blind_vectors(unsigned long long key, uint32_t n, uint32_t* __restrict__ sL,
                                                        uint32_t* __restrict__ sR /*[n][8] mont*/) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sc_store(sL + (size_t)i * 8, sc_from_stream_block(key, 2ull * i));
  sc_store(sR + (size_t)i * 8, sc_from_stream_block(key, 2ull * i + 1));
}

// Production form of the blinding vectors: the 64-byte blocks are ChaCha20 blocks (RFC 8439 2.3:
// 20 rounds, 32-bit block counter, 96-bit nonce "bpg sLsR v01") under a 256-bit key that the host draws
// from the transcript-bound RNG (reference src/r1cs/prover.rs:435-445: transcript state, v_blinding
// witnesses and external entropy); block 2j -> s_L[j], block 2j+1 -> s_R[j], each reduced mod l.
static __global__ void __launch_bounds__(256) k_blind_vectors_chacha(ChaKey key, uint32_t n, uint32_t* __restrict__ sL,
                                                               uint32_t* __restrict__ sR /*[n][8] mont*/) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sc_store(sL + (size_t)i * 8, sc_from_chacha_block(key, 2u * i));
  sc_store(sR + (size_t)i * 8, sc_from_chacha_block(key, 2u * i + 1u));
}

// ---- prover: terms of the (A_I, A_O, S) launch over gens[first .. first+cnt) -------------
// term layout (5 cnt + 3): [i_b o_b s_b | a_L | a_R | a_O | s_L | s_R]
static __global__ void __launch_bounds__(256) k_aios_terms(const uint32_t* __restrict__ aL, const uint32_t* __restrict__ aR,
                                                     const uint32_t* __restrict__ aO, const uint32_t* __restrict__ sL,
                                                     const uint32_t* __restrict__ sR, uint32_t first, uint32_t cnt,
                                                     const uint32_t* __restrict__ blind3 /*3 canonical scalars*/,
                                                     uint32_t g_base, uint32_t h_base, uint32_t bb_id,
                                                     uint32_t* __restrict__ scalars, uint32_t* __restrict__ point_ids,
                                                     uint8_t* __restrict__ set_ids) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3) {
    sc b;
    sc_load(b, blind3 + (size_t)i * 8);
    sc_store(scalars + (size_t)i * 8, b);
    point_ids[i] = bb_id;
    set_ids[i] = (uint8_t)i;
  }
  if (i >= cnt) return;
  const uint32_t* src[5] = {aL, aR, aO, sL, sR};
  const uint32_t base_id[5] = {g_base, h_base, g_base, g_base, h_base};
  const uint8_t set[5] = {0, 0, 1, 2, 2};
#pragma unroll
  for (int k = 0; k < 5; k++) {
    sc v;
    sc_load(v, src[k] + (size_t)(first + i) * 8);
    size_t t = 3 + (size_t)k * cnt + i;
    sc_store(scalars + t * 8, sc_from_mont(v));
    point_ids[t] = base_id[k] + first + i;
    set_ids[t] = set[k];
  }
}

// ---- prover: t_1..t_6 (util.rs:152-170 on the polynomials of prover.rs:596-617) ---------
struct PolyParams {
  PowTable y, y_inv;
};
static __global__ void __launch_bounds__(SV_THREADS) k_poly_t(const uint32_t* __restrict__ aL, const uint32_t* __restrict__ aR,
                                                        const uint32_t* __restrict__ aO, const uint32_t* __restrict__ sL,
                                                        const uint32_t* __restrict__ sR, const uint32_t* __restrict__ wL,
                                                        const uint32_t* __restrict__ wR, const uint32_t* __restrict__ wO,
                                                        uint32_t n, PolyParams pp,
                                                        uint32_t* __restrict__ partials /*[grid][48]*/) {
  __shared__ uint32_t sm[SV_THREADS / 2][48];
  sc t[6];
#pragma unroll
  for (int k = 0; k < 6; k++) t[k] = sc_zero();
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    sc al, ar, ao, sl, sr, wl, wr, wo;
    sc_load(al, aL + (size_t)i * 8); sc_load(ar, aR + (size_t)i * 8); sc_load(ao, aO + (size_t)i * 8);
    sc_load(sl, sL + (size_t)i * 8); sc_load(sr, sR + (size_t)i * 8);
    sc_load(wl, wL + (size_t)i * 8); sc_load(wr, wR + (size_t)i * 8); sc_load(wo, wO + (size_t)i * 8);
    sc yi = sc_pow(pp.y, i), yni = sc_pow(pp.y_inv, i);
    sc l1 = sc_add(al, sc_montmul(yni, wr));
    sc r0 = sc_sub(wo, yi);
    sc r1 = sc_add(sc_montmul(yi, ar), wl);
    sc r3 = sc_montmul(yi, sr);
    t[0] = sc_add(t[0], sc_montmul(l1, r0));
    t[1] = sc_add(t[1], sc_add(sc_montmul(l1, r1), sc_montmul(ao, r0)));
    t[2] = sc_add(t[2], sc_add(sc_montmul(ao, r1), sc_montmul(sl, r0)));
    t[3] = sc_add(t[3], sc_add(sc_montmul(l1, r3), sc_montmul(sl, r1)));
    t[4] = sc_add(t[4], sc_montmul(ao, r3));
    t[5] = sc_add(t[5], sc_montmul(sl, r3));
  }
  block_sum_k<6>(t, sm);
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < 6; k++) sc_store(partials + (size_t)blockIdx.x * 48 + 8 * k, t[k]);
}
// single block: sum the partials, leave canonical scalars
template <int K>
__global__ void __launch_bounds__(SV_THREADS) k_sum_partials(const uint32_t* __restrict__ partials, uint32_t nparts,
                                                              uint32_t* __restrict__ out /*[K][8] canonical*/) {
  __shared__ uint32_t sm[SV_THREADS / 2][8 * K];
  sc t[K];
#pragma unroll
  for (int k = 0; k < K; k++) t[k] = sc_zero();
  for (uint32_t i = threadIdx.x; i < nparts; i += blockDim.x)
#pragma unroll
    for (int k = 0; k < K; k++) {
      sc o;
      sc_load(o, partials + (size_t)i * 8 * K + 8 * k);
      t[k] = sc_add(t[k], o);
    }
  block_sum_k<K>(t, sm);
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < K; k++) sc_store(out + 8 * k, sc_from_mont(t[k]));
}

// ---- prover: l(x), r(x), padding, G/H factors (prover.rs:650-697) -> canonical --------------
struct EvalParams {
  PowTable y, y_inv;
  uint32_t x[8], u[8];  // Montgomery
  uint32_t n, n1, N;
};
static __global__ void __launch_bounds__(256) k_lr_eval(const uint32_t* __restrict__ aL, const uint32_t* __restrict__ aR,
                                                  const uint32_t* __restrict__ aO, const uint32_t* __restrict__ sL,
                                                  const uint32_t* __restrict__ sR, const uint32_t* __restrict__ wL,
                                                  const uint32_t* __restrict__ wR, const uint32_t* __restrict__ wO,
                                                  EvalParams ep, uint32_t* __restrict__ l_vec,
                                                  uint32_t* __restrict__ r_vec, uint32_t* __restrict__ g_fac,
                                                  uint32_t* __restrict__ h_fac) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ep.N) return;
  sc x = sc_param(ep.x), u = sc_param(ep.u), one = sc_const(BPG_K(K_R1));
  sc yi = sc_pow(ep.y, i), yni = sc_pow(ep.y_inv, i);
  sc l = sc_zero(), r;
  if (i < ep.n) {
    sc al, ar, ao, sl, sr, wl, wr, wo;
    sc_load(al, aL + (size_t)i * 8); sc_load(ar, aR + (size_t)i * 8); sc_load(ao, aO + (size_t)i * 8);
    sc_load(sl, sL + (size_t)i * 8); sc_load(sr, sR + (size_t)i * 8);
    sc_load(wl, wL + (size_t)i * 8); sc_load(wr, wR + (size_t)i * 8); sc_load(wo, wO + (size_t)i * 8);
    sc l1 = sc_add(al, sc_montmul(yni, wr));
    sc r0 = sc_sub(wo, yi);
    sc r1 = sc_add(sc_montmul(yi, ar), wl);
    sc r3 = sc_montmul(yi, sr);
    // l = x (l1 + x (l2 + x l3)),  r = r0 + x (r1 + x (x r3))      (util.rs:172-181, l0 = r2 = 0)
    l = sc_montmul(x, sc_add(l1, sc_montmul(x, sc_add(ao, sc_montmul(x, sl)))));
    r = sc_add(r0, sc_montmul(x, sc_add(r1, sc_montmul(x, sc_montmul(x, r3)))));
  } else {
    r = sc_neg(yi);  // prover.rs:661-672
  }
  sc gf = i < ep.n1 ? one : u;
  sc_store(l_vec + (size_t)i * 8, sc_from_mont(l));
  sc_store(r_vec + (size_t)i * 8, sc_from_mont(r));
  sc_store(g_fac + (size_t)i * 8, sc_from_mont(gf));
  sc_store(h_fac + (size_t)i * 8, sc_from_mont(sc_montmul(yni, gf)));
}

// ---- verifier: g_scalars, h_scalars, delta (verifier.rs:468-501) ------------------------
struct VerifyParams {
  PowTable y_inv;
  uint32_t u_sq[32][8];  // u_j^2 in creation order, Montgomery
  uint32_t allinv[8], x[8], a[8], b[8], u[8];
  uint32_t c0[8], c1[8];  // B scalar = c0 + c1 * delta
  uint32_t lg_n, n, n1, N;
};
static __global__ void __launch_bounds__(SV_THREADS) k_verify_scalars(const uint32_t* __restrict__ wL,
                                                                const uint32_t* __restrict__ wR,
                                                                const uint32_t* __restrict__ wO, VerifyParams vp,
                                                                uint32_t* __restrict__ g_out, uint32_t* __restrict__ h_out,
                                                                uint32_t* __restrict__ partials /*[grid][8]*/) {
  __shared__ uint32_t sm[SV_THREADS / 2][8];
  sc delta[1] = {sc_zero()};
  sc x = sc_param(vp.x), a = sc_param(vp.a), b = sc_param(vp.b), u = sc_param(vp.u), one = sc_const(BPG_K(K_R1));
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < vp.N; i += gridDim.x * blockDim.x) {
    // s_i = allinv * prod_{bits of i} u^2, s_{N-1-i} the same over the clear bits (inner_product_proof.rs:300-307)
    sc s = sc_param(vp.allinv), srev = s;
    for (uint32_t bit = 0; bit < vp.lg_n; bit++) {
      // every bit multiplies exactly one of the two: ONE product per bit for the whole warp, operand by select
      // (a branch would make a warp with mixed bits run both sides)
      const sc usq = sc_param(vp.u_sq[(vp.lg_n - 1) - bit]);
      const bool set = (i >> bit) & 1u;
      sc t = sc_montmul(sc_sel(set, s, srev), usq);
      s = sc_sel(set, t, s);
      srev = sc_sel(set, srev, t);
    }
    sc yni = sc_pow(vp.y_inv, i);
    sc wl = sc_zero(), wr = sc_zero(), wo = sc_zero();
    if (i < vp.n) {
      sc_load(wl, wL + (size_t)i * 8);
      sc_load(wr, wR + (size_t)i * 8);
      sc_load(wo, wO + (size_t)i * 8);
    }
    sc yneg_wr = sc_montmul(wr, yni);
    delta[0] = sc_add(delta[0], sc_montmul(yneg_wr, wl));
    sc U = i < vp.n1 ? one : u;
    sc g = sc_montmul(U, sc_sub(sc_montmul(x, yneg_wr), sc_montmul(a, s)));
    sc inner = sc_sub(sc_add(sc_montmul(x, wl), wo), sc_montmul(b, srev));
    sc h = sc_montmul(U, sc_sub(sc_montmul(yni, inner), one));
    sc_store(g_out + (size_t)i * 8, sc_from_mont(g));
    sc_store(h_out + (size_t)i * 8, sc_from_mont(h));
  }
  block_sum_k<1>(delta, sm);
  if (threadIdx.x == 0) sc_store(partials + (size_t)blockIdx.x * 8, delta[0]);
}
// single block: delta -> the scalar of B = c0 + c1*delta (verifier.rs:527-529), canonical
static __global__ void __launch_bounds__(SV_THREADS) k_verify_finish(const uint32_t* __restrict__ partials, uint32_t nparts,
                                                               VerifyParams vp, uint32_t* __restrict__ b_scalar_out) {
  __shared__ uint32_t sm[SV_THREADS / 2][8];
  sc d[1] = {sc_zero()};
  for (uint32_t i = threadIdx.x; i < nparts; i += blockDim.x) {
    sc o;
    sc_load(o, partials + (size_t)i * 8);
    d[0] = sc_add(d[0], o);
  }
  block_sum_k<1>(d, sm);
  if (threadIdx.x == 0) {
    sc r = sc_add(sc_param(vp.c0), sc_montmul(sc_param(vp.c1), d[0]));
    sc_store(b_scalar_out, sc_from_mont(r));
  }
}

// batch verification: out[i] = sum_j rho_j * slots[idx_j][i] over the (2 + 2N) generator scalars of the chosen proofs
static __global__ void __launch_bounds__(256) k_vbatch_combine(const uint32_t* __restrict__ slots, size_t stride_words,
                                                         const uint32_t* __restrict__ idx, uint32_t cnt,
                                                         const uint32_t* __restrict__ rho /*[cnt][8] canonical*/,
                                                         uint32_t total, uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  sc acc = sc_zero();
  for (uint32_t j = 0; j < cnt; j++) {
    sc x, r;
    sc_load(x, slots + (size_t)idx[j] * stride_words + (size_t)i * 8);
    sc_load(r, rho + (size_t)j * 8);
    acc = sc_add(acc, sc_montmul(x, sc_to_mont(r)));  // canonical * Montgomery -> canonical
  }
  sc_store(out + (size_t)i * 8, acc);
}

// ---- InnerProductProof::verify scalars (inner_product_proof.rs:283-307, 335-351) -----------
//   g_i = a * s_i * G_factors[i],  h_i = b * s_{N-1-i} * H_factors[i]   (1/s_i = s_{N-1-i})
struct IppVerifyParams {
  uint32_t u_sq[32][8];  // u_j^2 in creation order, Montgomery
  uint32_t allinv[8], a[8], b[8];
  uint32_t lg_n, N;
};
static __global__ void __launch_bounds__(256) k_ipp_verify_scalars(const uint32_t* __restrict__ Gf /*canonical or null*/,
                                                             const uint32_t* __restrict__ Hf, IppVerifyParams vp,
                                                             uint32_t* __restrict__ g_out, uint32_t* __restrict__ h_out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= vp.N) return;
  sc s = sc_param(vp.allinv), srev = s;
  for (uint32_t bit = 0; bit < vp.lg_n; bit++) {  // one product per bit, operand by select (see k_verify_scalars)
    const sc usq = sc_param(vp.u_sq[(vp.lg_n - 1) - bit]);
    const bool set = (i >> bit) & 1u;
    sc t = sc_montmul(sc_sel(set, s, srev), usq);
    s = sc_sel(set, t, s);
    srev = sc_sel(set, srev, t);
  }
  sc g = sc_montmul(sc_param(vp.a), s), h = sc_montmul(sc_param(vp.b), srev);
  // a canonical factor f times a Montgomery value x R: montmul(xR, f) = x f, already canonical
  if (Gf) {
    sc f;
    sc_load(f, Gf + (size_t)i * 8);
    g = sc_montmul(g, f);
  } else {
    g = sc_from_mont(g);
  }
  if (Hf) {
    sc f;
    sc_load(f, Hf + (size_t)i * 8);
    h = sc_montmul(h, f);
  } else {
    h = sc_from_mont(h);
  }
  sc_store(g_out + (size_t)i * 8, g);
  sc_store(h_out + (size_t)i * 8, h);
}

// ---- flattened constraints (prover.rs:342-379, verifier.rs:323-362) ----------------------
// Constraint q (row q) contributes z^(q+1) * coeff to the weight of each variable it names:
//   wL[i], wR[i], wO[i] += ...;   wV[j] -= ...;   wc -= ... (the constant term).
// One thread per TERM forms its product and adds it into its variable's accumulator: eight
// 64-bit counters, one per 32-bit limb of the Montgomery residue, with plain integer atomics
// (exact and order-independent; 2^32 terms cannot overflow a counter).  The constant is named
// by a large share of the rows of some circuits, so its terms are summed per block first.
// k_flat_finish carries the counters, reduces mod l and applies the signs.
constexpr uint32_t FLAT_KIND_SHIFT = 28;  // term code: kind << 28 | index; kinds as in the host mirror
constexpr uint32_t FLAT_IDX_MASK = (1u << 27) - 1, FLAT_NEG = 1u << 31;
enum : uint32_t { FK_LEFT = 1, FK_RIGHT = 2, FK_OUT = 3, FK_COMMITTED = 4, FK_ONE = 5, FK_ZERO = 6 };

static __global__ void __launch_bounds__(SV_THREADS) k_flat_terms(const uint32_t* __restrict__ t_code,
                                                            const uint32_t* __restrict__ t_row,
                                                            const uint32_t* __restrict__ t_coeff /*[n_general] Montgomery*/,
                                                            uint32_t n_general, uint32_t n_terms /*general, then unit*/,
                                                            uint32_t n, uint32_t m, PowTable z,
                                                            unsigned long long* __restrict__ acc /*[3n+m+1][8]*/) {
  __shared__ uint32_t sm[SV_THREADS / 2][8];
  sc one_sum[1] = {sc_zero()};
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_terms) {
    const uint32_t code = t_code[t], kind = (code >> FLAT_KIND_SHIFT) & 7u, idx = code & FLAT_IDX_MASK;
    sc v = sc_pow(z, t_row[t] + 1);
    if (t < n_general) {
      sc c;
      sc_load(c, t_coeff + (size_t)t * 8);
      v = sc_montmul(c, v);
    } else if (code & FLAT_NEG) {  // unit terms carry no coefficient: +1, or -1 with the top bit of the code
      v = sc_sub(sc_zero(), v);
    }
    if (kind == FK_ONE) {
      one_sum[0] = v;
    } else if (kind >= FK_LEFT && kind <= FK_COMMITTED) {
      size_t key = kind == FK_COMMITTED ? (size_t)3 * n + idx : (size_t)(kind - 1) * n + idx;
      unsigned long long* a = acc + key * 8;
#pragma unroll
      for (int j = 0; j < 8; j++) atomicAdd(a + j, (unsigned long long)v.v[j]);
    }
  }
  block_sum_k<1>(one_sum, sm);
  if (threadIdx.x == 0) {
    unsigned long long* a = acc + ((size_t)3 * n + m) * 8;
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (one_sum[0].v[j]) atomicAdd(a + j, (unsigned long long)one_sum[0].v[j]);
  }
}
// counters -> Montgomery residues; keys [0,3n): +, keys [3n, 3n+m]: negated
static __global__ void __launch_bounds__(256) k_flat_finish(const unsigned long long* __restrict__ acc, uint32_t n, uint32_t m,
                                                      uint32_t* __restrict__ w3 /*[3n][8]: wL | wR | wO*/,
                                                      uint32_t* __restrict__ wv /*[m+1][8]: wV | wc*/) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t total = 3 * n + m + 1;
  if (k >= total) return;
  const unsigned long long* a = acc + (size_t)k * 8;
  sc lo;
  unsigned long long carry = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    unsigned long long x = a[j] + carry;  // a[j] < 2^64 - 2^32 in practice: no wrap for < 2^31 terms
    lo.v[j] = (uint32_t)x;
    carry = x >> 32;
  }
  // value = lo + carry * 2^256, carry < 2^33:  lo mod l  +  carry * (2^256 mod l)
  sc hi = sc_zero();
  hi.v[0] = (uint32_t)carry;
  hi.v[1] = (uint32_t)(carry >> 32);
  sc r = sc_add(sc_montmul(lo, sc_const(BPG_K(K_R1))), sc_to_mont(hi));
  if (k >= 3 * n) {
    sc_store(wv + (size_t)(k - 3 * n) * 8, sc_neg(r));
  } else {
    sc_store(w3 + (size_t)k * 8, r);
  }
}

}  // namespace bpg
