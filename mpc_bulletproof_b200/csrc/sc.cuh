// Scalars modulo the ristretto255 group order
//   l = 2^252 + 27742317777372353535851937790883648493
// in eight 32-bit limbs, Montgomery form with R = 2^256.
//
// Backs the O(n) scalar-vector work that feeds the MSMs: the a/b folds and
// cross terms of reference src/inner_product_proof.rs:87-88,224-225,463-472,
// `verification_scalars` (:254-310) and the verifier's scalar preparation
// (reference src/r1cs/verifier.rs:468-501).
#pragma once
#include "fe.cuh"

namespace bpg {

struct sc {
  uint32_t v[8];
};

// signed 4-bit windows of a 256-bit scalar: the fixed-base comb (64 x 8 multiples) and the few-term MSM
constexpr int COMB_WINDOWS = 64;
constexpr int COMB_ENTRIES = COMB_WINDOWS * 8;

#define BPG_DEF_CONST_SC(name, ...)                         \
  static __device__ __constant__ uint32_t name[8] = {__VA_ARGS__}; \
  static const uint32_t name##_h[8] = {__VA_ARGS__};
#ifndef BPG_K
#if defined(__CUDA_ARCH__)
#define BPG_K(name) name
#else
#define BPG_K(name) name##_h
#endif
#endif

BPG_DEF_CONST_SC(K_L, 0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0x00000000u, 0x00000000u,
                 0x00000000u, 0x10000000u)
// R^2 mod l and R mod l (R = 2^256)
BPG_DEF_CONST_SC(K_RR, 0x449c0f01u, 0xa40611e3u, 0x68859347u, 0xd00e1ba7u, 0x17f5be65u, 0xceec73d2u,
                 0x7c309a3du, 0x0399411bu)
BPG_DEF_CONST_SC(K_R1, 0x8d98951du, 0xd6ec3174u, 0x737dcf70u, 0xc6ef5bf4u, 0xfffffffeu, 0xffffffffu,
                 0xffffffffu, 0x0fffffffu)
#define BPG_L_NINV32 0x12547e1bu  // -l^{-1} mod 2^32

BPG_DI sc sc_const(const uint32_t* k) {
  sc o;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = k[i];
  return o;
}
BPG_DI sc sc_zero() {
  sc o;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = 0;
  return o;
}

// 256x256 -> 512 product, same even/odd column schedule as fe_mul.
BPG_DI void mul256_wide(uint32_t r[16], const uint32_t* a, const uint32_t* b) {
  uint32_t e[16], o[16];
  mul_wide(e[0], e[1], a[0], b[0]);
  mul_wide(e[2], e[3], a[2], b[0]);
  mul_wide(e[4], e[5], a[4], b[0]);
  mul_wide(e[6], e[7], a[6], b[0]);
  mul_wide(o[0], o[1], a[1], b[0]);
  mul_wide(o[2], o[3], a[3], b[0]);
  mul_wide(o[4], o[5], a[5], b[0]);
  mul_wide(o[6], o[7], a[7], b[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) {
    if (i & 1) {
      mad_wide_cc(e[i + 1], e[i + 2], a[1], b[i]);
      madc_wide_cc(e[i + 3], e[i + 4], a[3], b[i]);
      madc_wide_cc(e[i + 5], e[i + 6], a[5], b[i]);
      if (i == 1) e[i + 7] = 0;
      madc_wide_top(e[i + 7], e[i + 8], a[7], b[i]);
      mad_wide_cc(o[i - 1], o[i], a[0], b[i]);
      madc_wide_cc(o[i + 1], o[i + 2], a[2], b[i]);
      madc_wide_cc(o[i + 3], o[i + 4], a[4], b[i]);
      madc_wide_cc(o[i + 5], o[i + 6], a[6], b[i]);
      o[i + 7] = addc(0u, 0u);
    } else {
      mad_wide_cc(e[i], e[i + 1], a[0], b[i]);
      madc_wide_cc(e[i + 2], e[i + 3], a[2], b[i]);
      madc_wide_cc(e[i + 4], e[i + 5], a[4], b[i]);
      madc_wide_cc(e[i + 6], e[i + 7], a[6], b[i]);
      e[i + 8] = addc(0u, 0u);
      mad_wide_cc(o[i], o[i + 1], a[1], b[i]);
      madc_wide_cc(o[i + 2], o[i + 3], a[3], b[i]);
      madc_wide_cc(o[i + 4], o[i + 5], a[5], b[i]);
      madc_wide_top(o[i + 6], o[i + 7], a[7], b[i]);
    }
  }
  r[0] = e[0];
  r[1] = add_cc(e[1], o[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) r[k] = addc_cc(e[k], o[k - 1]);
  r[15] = addc(e[15], o[14]);
}

// x >= l ?
BPG_DI bool sc_geq_l(const uint32_t* x) {
  const uint32_t* l = BPG_K(K_L);
  sub_cc(x[0], l[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) subc_cc(x[i], l[i]);
  uint32_t bw = subc(0u, 0u);
  return bw == 0;
}

// x (< 2l) -> x mod l
BPG_DI sc sc_cond_sub_l(const uint32_t* x) {
  const uint32_t* l = BPG_K(K_L);
  uint32_t d[8];
  d[0] = sub_cc(x[0], l[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) d[i] = subc_cc(x[i], l[i]);
  uint32_t bw = subc(0u, 0u);
  sc o;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = bw ? x[i] : d[i];
  return o;
}

// Montgomery product a*b*R^-1 mod l; a, b < l.
BPG_DI sc sc_montmul_inl(const sc& a, const sc& b) {
  uint32_t t[17];
  mul256_wide(t, a.v, b.v);
  t[16] = 0;
  const uint32_t* l = BPG_K(K_L);
  // l = (l3 l2 l1 l0) + 2^252: limbs 4..6 are zero, limb 7 = 2^28.
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t m = t[i] * BPG_L_NINV32;
    // t += m * l * 2^(32 i)
    uint32_t lo, hi, c;
    // low four limbs
    mul_wide(lo, hi, m, l[0]);
    t[i] = add_cc(t[i], lo);
    c = addc(hi, 0u);  // hi <= 2^32-2: no overflow
    mul_wide(lo, hi, m, l[1]);
    lo = add_cc(lo, c);
    hi = addc(hi, 0u);
    t[i + 1] = add_cc(t[i + 1], lo);
    c = addc(hi, 0u);
    mul_wide(lo, hi, m, l[2]);
    lo = add_cc(lo, c);
    hi = addc(hi, 0u);
    t[i + 2] = add_cc(t[i + 2], lo);
    c = addc(hi, 0u);
    mul_wide(lo, hi, m, l[3]);
    lo = add_cc(lo, c);
    hi = addc(hi, 0u);
    t[i + 3] = add_cc(t[i + 3], lo);
    c = addc(hi, 0u);
    // carry c into limb i+4, plus m * 2^28 at limb i+7 (lo) / i+8 (hi)
    t[i + 4] = add_cc(t[i + 4], c);
    t[i + 5] = addc_cc(t[i + 5], 0u);
    t[i + 6] = addc_cc(t[i + 6], 0u);
    t[i + 7] = addc_cc(t[i + 7], m << 28);
    t[i + 8] = addc_cc(t[i + 8], m >> 4);
#pragma unroll
    for (int k = i + 9; k < 17; k++) t[k] = addc_cc(t[k], 0u);
  }
  // result = t[8..16] < 2l
  return sc_cond_sub_l(t + 8);
}
BPG_DI sc sc_sel(bool p, const sc& a, const sc& b) {
  sc o;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = p ? a.v[i] : b.v[i];
  return o;
}
// out of line in the latency-bound translation units (see fe.cuh, BPG_FE_OUTLINE)
#if (defined(BPG_FE_OUTLINE) || defined(BPG_GE_OUTLINE)) && defined(__CUDA_ARCH__)
static __device__ __noinline__ sc sc_montmul_call(sc a, sc b) { return sc_montmul_inl(a, b); }
__device__ __forceinline__ sc sc_montmul(const sc& a, const sc& b) { return sc_montmul_call(a, b); }
#else
BPG_DI sc sc_montmul(const sc& a, const sc& b) { return sc_montmul_inl(a, b); }
#endif

BPG_DI sc sc_add(const sc& a, const sc& b) {
  uint32_t s[8];
  s[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) s[i] = addc_cc(a.v[i], b.v[i]);
  s[7] = addc(a.v[7], b.v[7]);  // < 2^254
  return sc_cond_sub_l(s);
}
BPG_DI sc sc_sub(const sc& a, const sc& b) {
  const uint32_t* l = BPG_K(K_L);
  uint32_t d[8];
  d[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) d[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t bw = subc(0u, 0u);
  sc o;
  o.v[0] = add_cc(d[0], bw & l[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) o.v[i] = addc_cc(d[i], bw & l[i]);
  o.v[7] = addc(d[7], bw & l[7]);
  return o;
}
BPG_DI sc sc_neg(const sc& a) { return sc_sub(sc_zero(), a); }

BPG_DI sc sc_to_mont(const sc& a) { return sc_montmul(a, sc_const(BPG_K(K_RR))); }
BPG_DI sc sc_from_mont(const sc& a) {
  sc one = sc_zero();
  one.v[0] = 1;
  return sc_montmul(a, one);
}
// plain product a*b mod l of two normal-form values
BPG_DI sc sc_mul(const sc& a, const sc& b) { return sc_montmul(sc_montmul(a, b), sc_const(BPG_K(K_RR))); }

BPG_DI void sc_load(sc& o, const uint32_t* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w;
  o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
}
BPG_DI void sc_store(uint32_t* p, const sc& o) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(o.v[0], o.v[1], o.v[2], o.v[3]);
  q[1] = make_uint4(o.v[4], o.v[5], o.v[6], o.v[7]);
}

// ---- signed c-bit window digits ---------------------------------------------
// For a scalar k < 2^253 (canonical scalars are < l < 2^253) and window width c, with W = ceil(255/c)
// windows:  k = sum_w d_w 2^(c w),  d_w in [-2^(c-1), 2^(c-1)).
// Adding 2^(c-1) into every window up front turns the carry recursion into one
// 256-bit addition; window w's digit is then a bit-field minus 2^(c-1).
struct sc_recoded {
  uint32_t v[9];
};
struct sc_bias {  // sum_w 2^(c w + c - 1); built on the host once per (c, W), passed by value
  uint32_t v[9];
};
BPG_DI sc_recoded sc_recode(const uint32_t k[8], const sc_bias& bias) {
  sc_recoded r;
  r.v[0] = add_cc(k[0], bias.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) r.v[i] = addc_cc(k[i], bias.v[i]);
  r.v[8] = addc(0u, bias.v[8]);
  return r;
}
// digit of window w: value in [-2^(c-1), 2^(c-1)).  Limb selection is a predicated
// scan so that r stays in registers for a run-time c.
BPG_DI int sc_digit(const sc_recoded& r, int w, int c) {
  int bit = c * w;
  int limb = bit >> 5, sh = bit & 31;
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    lo = (i == limb) ? r.v[i] : lo;
    hi = (i == limb + 1) ? r.v[i] : hi;
  }
  uint64_t two = ((uint64_t)hi << 32) | lo;
  uint32_t raw = (uint32_t)(two >> sh) & ((1u << c) - 1u);
  return (int)raw - (1 << (c - 1));
}

// ---- ChaCha20 block -> scalar (the device side of the keyed blinding vectors, svec_kernels.cuh) ----
struct ChaKey {
  uint32_t k[8];
};
BPG_DI void chacha_qr(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  a += b; d ^= a; d = (d << 16) | (d >> 16);
  c += d; b ^= c; b = (b << 12) | (b >> 20);
  a += b; d ^= a; d = (d << 8) | (d >> 24);
  c += d; b ^= c; b = (b << 7) | (b >> 25);
}
BPG_DI sc sc_from_chacha_block(const ChaKey& key, uint32_t block) {
  uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3],
                     key.k[4],    key.k[5],    key.k[6],    key.k[7],    block,    0x20677062u, 0x52734c73u, 0x31307620u};
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = in[i];
#pragma unroll
  for (int r = 0; r < 10; r++) {
    chacha_qr(x[0], x[4], x[8], x[12]);
    chacha_qr(x[1], x[5], x[9], x[13]);
    chacha_qr(x[2], x[6], x[10], x[14]);
    chacha_qr(x[3], x[7], x[11], x[15]);
    chacha_qr(x[0], x[5], x[10], x[15]);
    chacha_qr(x[1], x[6], x[11], x[12]);
    chacha_qr(x[2], x[7], x[8], x[13]);
    chacha_qr(x[3], x[4], x[9], x[14]);
  }
  sc lo, hi;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    lo.v[i] = x[i] + in[i];
    hi.v[i] = x[8 + i] + in[8 + i];
  }
  sc rr = sc_const(BPG_K(K_RR));
  sc lo_m = sc_montmul(lo, rr);
  sc hi_m = sc_montmul(sc_montmul(hi, rr), rr);
  return sc_add(lo_m, hi_m);
}

}  // namespace bpg
