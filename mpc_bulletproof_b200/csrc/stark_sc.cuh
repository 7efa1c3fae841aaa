// Stark-curve policy, scalar layer: arithmetic modulo the group order
//   n = 0x0800000000000010ffffffffffffffffb781126dcae7b2321e66a241adc64d2f  (252 bits, prime)
// in eight 32-bit limbs, Montgomery form with R = 2^256 (mpc-stark's `Scalar` is the same field in
// 4x64 Montgomery limbs).  Backs the inner-product-argument scalar vectors of the Stark policy
// (reference src/inner_product_proof.rs:87-88, 224-225); n has no sparse structure, so the reduction
// is the plain word-by-word one.  A policy struct exposes it to the templated IPP kernels next to
// the ristretto255 scalars of sc.cuh.
#pragma once
#include "sc.cuh"

namespace bpg {

BPG_DEF_CONST_SC(KSS_N, 0xadc64d2fu, 0x1e66a241u, 0xcae7b232u, 0xb781126du, 0xffffffffu, 0xffffffffu, 0x00000010u,
                 0x08000000u)
BPG_DEF_CONST_SC(KSS_R1, 0xf4fca74fu, 0x51925a0bu, 0x6df16beeu, 0xc75ec4b4u, 0x00000008u, 0x00000000u, 0xfffffdf1u,
                 0x07ffffffu)  // R mod n
BPG_DEF_CONST_SC(KSS_RR, 0xea1c688du, 0x6021b3f1u, 0x14ce60b9u, 0x509cf64du, 0xf78bbabbu, 0xbaf0ab4cu, 0x2333766eu,
                 0x07d9e57cu)  // R^2 mod n
#define BPG_SS_NINV32 0xe8bde631u  // -n^{-1} mod 2^32

// x (< 2n) -> x mod n
BPG_DI sc scs_cond_sub_n(const uint32_t* x) {
  const uint32_t* n = BPG_K(KSS_N);
  uint32_t d[8];
  d[0] = sub_cc(x[0], n[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) d[i] = subc_cc(x[i], n[i]);
  uint32_t bw = subc(0u, 0u);
  sc o;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = bw ? x[i] : d[i];
  return o;
}

// Montgomery product a*b*R^-1 mod n; b < n, a any 256-bit value
BPG_DI sc scs_montmul(const sc& a, const sc& b) {
  uint32_t t[17];
  mul256_wide(t, a.v, b.v);
  t[16] = 0;
  const uint32_t* n = BPG_K(KSS_N);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t m = t[i] * BPG_SS_NINV32;
    unsigned long long carry = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      unsigned long long acc = (unsigned long long)m * n[j] + t[i + j] + carry;
      t[i + j] = (uint32_t)acc;
      carry = acc >> 32;
    }
#pragma unroll
    for (int k = i + 8; k < 17; k++) {
      unsigned long long acc = (unsigned long long)t[k] + carry;
      t[k] = (uint32_t)acc;
      carry = acc >> 32;
    }
  }
  return scs_cond_sub_n(t + 8);  // < 2n
}
BPG_DI sc scs_add(const sc& a, const sc& b) {
  uint32_t s[8];
  s[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) s[i] = addc_cc(a.v[i], b.v[i]);
  s[7] = addc(a.v[7], b.v[7]);  // < 2^253
  return scs_cond_sub_n(s);
}
BPG_DI sc scs_sub(const sc& a, const sc& b) {
  const uint32_t* n = BPG_K(KSS_N);
  uint32_t d[8];
  d[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) d[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t bw = subc(0u, 0u);
  sc o;
  o.v[0] = add_cc(d[0], bw & n[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) o.v[i] = addc_cc(d[i], bw & n[i]);
  o.v[7] = addc(d[7], bw & n[7]);
  return o;
}

// scalar policies for the templated IPP kernels (ipp_kernels_t.cuh)
struct ScRistretto {
  static BPG_DI sc montmul(const sc& a, const sc& b) { return sc_montmul(a, b); }
  static BPG_DI sc add(const sc& a, const sc& b) { return sc_add(a, b); }
  static BPG_DI sc one_m() { return sc_const(BPG_K(K_R1)); }
  static BPG_DI sc rr() { return sc_const(BPG_K(K_RR)); }
};
struct ScStark {
  static BPG_DI sc montmul(const sc& a, const sc& b) { return scs_montmul(a, b); }
  static BPG_DI sc add(const sc& a, const sc& b) { return scs_add(a, b); }
  static BPG_DI sc one_m() { return sc_const(BPG_K(KSS_R1)); }
  static BPG_DI sc rr() { return sc_const(BPG_K(KSS_RR)); }
};

}  // namespace bpg
