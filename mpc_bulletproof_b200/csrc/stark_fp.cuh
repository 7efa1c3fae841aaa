// The Stark-curve policy, field layer: arithmetic modulo
//   p = 2^251 + 17*2^192 + 1
// in eight 32-bit limbs, Montgomery form with R = 2^256, values fully reduced to [0, p).
//
// This is the field of `mpc_stark::algebra::stark_curve::StarkPoint`, the group the mounted
// reference computes over (reference Cargo.toml:13,21; src/generators.rs:11-16).
//
// p = 1 (mod 2^192), so -p^-1 = -1 (mod 2^32) and the Montgomery quotient digit of limb i is
// just -(t_i + carry).  For the low six limbs the digits depend only on a one-bit carry, so they
// are produced by a scan; their joint contribution to the upper limbs is U * (2^59 + 17) * 2^192
// (U the six digits as one number): one 17x multiply and one 59-bit shift instead of forty-eight
// multiply-adds.  Only the last two digits need the sequential step.
#pragma once
#include "sc.cuh"

namespace bpg {

struct fp {
  uint32_t v[8];
};

#define BPG_DEF_CONST_FP(name, ...)                         \
  static __device__ __constant__ uint32_t name[8] = {__VA_ARGS__}; \
  static const uint32_t name##_h[8] = {__VA_ARGS__};

BPG_DEF_CONST_FP(KS_P, 0x00000001u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000011u,
                 0x08000000u)
BPG_DEF_CONST_FP(KS_R1, 0xffffffe1u, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xfffffdf0u,
                 0x07ffffffu)  // R mod p = Montgomery form of 1
BPG_DEF_CONST_FP(KS_RR, 0x7e000401u, 0xfffffd73u, 0x330fffffu, 0x00000001u, 0xff6f8000u, 0xffffffffu, 0x5e008810u,
                 0x07ffd4abu)  // R^2 mod p
BPG_DEF_CONST_FP(KS_BETA, 0xb59a21cau, 0x359ddd67u, 0x7aab9006u, 0x6725f223u, 0x2a41f947u, 0xab8a1e00u, 0x1774247fu,
                 0x01393165u)  // curve constant beta, Montgomery form

BPG_DI fp fp_const(const uint32_t* k) {
  fp o;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = k[i];
  return o;
}
BPG_DI fp fp_zero() {
  fp o;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = 0;
  return o;
}
BPG_DI fp fp_one() { return fp_const(BPG_K(KS_R1)); }

// x (< 2p) -> x mod p
BPG_DI fp fp_cond_sub_p(const uint32_t* x) {
  // p = [1, 0, 0, 0, 0, 0, 17, 2^27]
  uint32_t d[8];
  d[0] = sub_cc(x[0], 1u);
  d[1] = subc_cc(x[1], 0u);
  d[2] = subc_cc(x[2], 0u);
  d[3] = subc_cc(x[3], 0u);
  d[4] = subc_cc(x[4], 0u);
  d[5] = subc_cc(x[5], 0u);
  d[6] = subc_cc(x[6], 17u);
  d[7] = subc_cc(x[7], 0x08000000u);
  uint32_t bw = subc(0u, 0u);  // all ones if x < p
  fp o;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = bw ? x[i] : d[i];
  return o;
}

BPG_DI fp fp_add(const fp& a, const fp& b) {
  uint32_t s[8];
  s[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) s[i] = addc_cc(a.v[i], b.v[i]);
  s[7] = addc(a.v[7], b.v[7]);  // < 2p < 2^253
  return fp_cond_sub_p(s);
}
BPG_DI fp fp_sub(const fp& a, const fp& b) {
  uint32_t d[8];
  d[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) d[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t bw = subc(0u, 0u);
  fp o;
  o.v[0] = add_cc(d[0], bw & 1u);
  o.v[1] = addc_cc(d[1], 0u);
  o.v[2] = addc_cc(d[2], 0u);
  o.v[3] = addc_cc(d[3], 0u);
  o.v[4] = addc_cc(d[4], 0u);
  o.v[5] = addc_cc(d[5], 0u);
  o.v[6] = addc_cc(d[6], bw & 17u);
  o.v[7] = addc(d[7], bw & 0x08000000u);
  return o;
}
BPG_DI fp fp_neg(const fp& a) { return fp_sub(fp_zero(), a); }
BPG_DI fp fp_dbl(const fp& a) { return fp_add(a, a); }

BPG_DI bool fp_is_zero(const fp& a) {
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) x |= a.v[i];
  return x == 0;
}
BPG_DI bool fp_eq(const fp& a, const fp& b) {
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) x |= a.v[i] ^ b.v[i];
  return x == 0;
}
BPG_DI fp fp_sel(bool p, const fp& a, const fp& b) {
  fp o;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = p ? a.v[i] : b.v[i];
  return o;
}

// Montgomery reduction of a 512-bit value t (t[16] scratch, must be 0 on entry): t / 2^256 mod p
BPG_DI fp fp_redc(uint32_t t[17]) {
  // Stage A: quotient digits of limbs 0..5 by a carry scan
  uint32_t m[6];
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    m[i] = 0u - (t[i] + c);
    c = (t[i] | c) != 0u ? 1u : 0u;
  }
  // Stage B: t[6..] += U*17 + c  and  t[7..] += U << 27  (U = m[0..5]; 2^251 = 2^59 * 2^192)
  uint32_t v[7];
  {
    unsigned long long acc = c;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      acc += (unsigned long long)m[k] * 17ull;
      v[k] = (uint32_t)acc;
      acc >>= 32;
    }
    v[6] = (uint32_t)acc;
  }
  t[6] = add_cc(t[6], v[0]);
#pragma unroll
  for (int k = 1; k < 7; k++) t[6 + k] = addc_cc(t[6 + k], v[k]);
  t[13] = addc_cc(t[13], 0u);
  t[14] = addc_cc(t[14], 0u);
  t[15] = addc_cc(t[15], 0u);
  t[16] = addc(t[16], 0u);
  t[7] = add_cc(t[7], m[0] << 27);
#pragma unroll
  for (int k = 1; k < 6; k++) t[7 + k] = addc_cc(t[7 + k], (m[k] << 27) | (m[k - 1] >> 5));
  t[13] = addc_cc(t[13], m[5] >> 5);
  t[14] = addc_cc(t[14], 0u);
  t[15] = addc_cc(t[15], 0u);
  t[16] = addc(t[16], 0u);
  // Stage C: limbs 6 and 7, one digit at a time:  t += m * (1 + (2^59 + 17) * 2^192) * 2^(32 i)
#pragma unroll
  for (int i = 6; i < 8; i++) {
    uint32_t x = t[i];
    uint32_t mi = 0u - x;
    uint32_t cy = x != 0u ? 1u : 0u;
    unsigned long long q01 = (unsigned long long)mi * 17ull + ((unsigned long long)(mi << 27) << 32);
    uint32_t q0 = (uint32_t)q01, q1 = (uint32_t)(q01 >> 32);
    // (mi * 17) >> 32 <= 16 and (mi << 27) can wrap the 64-bit sum: carry into q2
    uint32_t hi17 = (uint32_t)(((unsigned long long)mi * 17ull) >> 32);
    uint32_t wrap = q1 < hi17 ? 1u : 0u;
    uint32_t q2 = (mi >> 5) + wrap;
    t[i + 1] = add_cc(t[i + 1], cy);
#pragma unroll
    for (int k = i + 2; k < i + 6; k++) t[k] = addc_cc(t[k], 0u);
    t[i + 6] = addc_cc(t[i + 6], q0);
    t[i + 7] = addc_cc(t[i + 7], q1);
    t[i + 8] = addc_cc(t[i + 8], q2);
#pragma unroll
    for (int k = i + 9; k < 16; k++) t[k] = addc_cc(t[k], 0u);
    t[16] = addc(t[16], 0u);
  }
  return fp_cond_sub_p(t + 8);
}

BPG_DI fp fp_mul(const fp& a, const fp& b) {
  uint32_t t[17];
  mul256_wide(t, a.v, b.v);
  t[16] = 0;
  return fp_redc(t);
}
BPG_DI fp fp_sq(const fp& a) { return fp_mul(a, a); }

BPG_DI fp fp_to_mont(const fp& a) { return fp_mul(a, fp_const(BPG_K(KS_RR))); }
BPG_DI fp fp_from_mont(const fp& a) {
  uint32_t t[17];
#pragma unroll
  for (int i = 0; i < 8; i++) t[i] = a.v[i];
#pragma unroll
  for (int i = 8; i < 17; i++) t[i] = 0;
  return fp_redc(t);
}

BPG_DI fp fp_sqn(fp a, int n) {
  for (int i = 0; i < n; i++) a = fp_sq(a);
  return a;
}
// a^(p-2);  p - 2 = 2^251 + 2^196 + (2^192 - 1)
BPG_DI fp fp_invert(const fp& a) {
  // a^(2^192 - 1) by the doubling ladder a^(2^2k - 1) = (a^(2^k - 1))^(2^k) * a^(2^k - 1): 191 squarings, 8 products
  fp x2 = fp_mul(fp_sq(a), a);
  fp x3 = fp_mul(fp_sq(x2), a);
  fp x6 = fp_mul(fp_sqn(x3, 3), x3);
  fp x12 = fp_mul(fp_sqn(x6, 6), x6);
  fp x24 = fp_mul(fp_sqn(x12, 12), x12);
  fp x48 = fp_mul(fp_sqn(x24, 24), x24);
  fp x96 = fp_mul(fp_sqn(x48, 48), x48);
  fp low = fp_mul(fp_sqn(x96, 96), x96);
  // a^(2^251 + 2^196) = (a^(2^55 + 1))^(2^196)
  fp h = fp_mul(fp_sqn(a, 55), a);
  h = fp_sqn(h, 196);
  return fp_mul(h, low);
}

BPG_DI void fp_load(fp& o, const uint32_t* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w;
  o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
}
BPG_DI void fp_store(uint32_t* p, const fp& o) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(o.v[0], o.v[1], o.v[2], o.v[3]);
  q[1] = make_uint4(o.v[4], o.v[5], o.v[6], o.v[7]);
}

// canonical little-endian coordinate (< p) -> Montgomery form; false if >= p
BPG_DI bool fp_from_canonical(fp& o, const uint32_t w[8]) {
  // w < p  <=>  w - p borrows
  sub_cc(w[0], 1u);
  subc_cc(w[1], 0u);
  subc_cc(w[2], 0u);
  subc_cc(w[3], 0u);
  subc_cc(w[4], 0u);
  subc_cc(w[5], 0u);
  subc_cc(w[6], 17u);
  subc_cc(w[7], 0x08000000u);
  uint32_t bw = subc(0u, 0u);
  fp a;
#pragma unroll
  for (int i = 0; i < 8; i++) a.v[i] = w[i];
  o = fp_to_mont(a);
  return bw != 0u;
}

}  // namespace bpg
