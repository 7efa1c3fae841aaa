// GF(2^255 - 19) in eight saturated 32-bit limbs, sm_100a.
//
// Replaces the field layer underneath the reference's group calls
// (`StarkPoint::msm_iter` etc., reference src/inner_product_proof.rs:90-114);
// ristretto255 instantiation per SURVEY.md §0-D1.
//
// Representation: little-endian limbs v[0..8), value = sum v[i] 2^(32 i), any
// value in [0, 2^256) is a valid representative ("loose").  fe_mul / fe_sq
// return a value < 2^255 ("tight"), so that one add of two tight values or one
// doubling never leaves 256 bits; fe_add / fe_sub accept loose inputs and fold
// the carry/borrow back with 2^256 = 38 (mod p).
//
// The 8x8 product is accumulated in two interleaved columns of 64-bit
// accumulators (even- and odd-aligned) with mad.lo.cc / madc.hi.cc pairs, which
// ptxas fuses to IMAD.WIDE.U32[.X] carrying through a predicate: 64 wide
// multiply-adds for the product + 8 for the 2^256 -> 38 fold.
#pragma once
#include <stdint.h>

namespace bpg {

struct fe {
  uint32_t v[8];
};

#define BPG_DI __host__ __device__ __forceinline__

// ---- carry-chain primitives ------------------------------------------------
// Device: PTX carry-flag instructions.  Host (tests/hostsim only, never the
// product path): the same primitives with the flag modelled in a thread-local,
// so the limb schedules below can be checked against oracle/ without a GPU.
#if defined(__CUDA_ARCH__)
BPG_DI uint32_t add_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
BPG_DI uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
BPG_DI uint32_t addc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
BPG_DI uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
BPG_DI uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
BPG_DI uint32_t subc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// (lo,hi) = a*b                       -- starts a column, no carry in/out
BPG_DI void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;"
               : "=r"(lo), "=r"(hi)
               : "r"(a), "r"(b));
}
// (lo,hi) += a*b, carry out           -- first link of a chain
BPG_DI void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
               : "+r"(lo), "+r"(hi)
               : "r"(a), "r"(b));
}
// (lo,hi) += a*b + carry in, carry out -- inner link
BPG_DI void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
               : "+r"(lo), "+r"(hi)
               : "r"(a), "r"(b));
}
// lo += a*b.lo + carry in; hi = a*b.hi + carry (hi column is fresh; cannot carry out)
BPG_DI void madc_wide_top(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, 0;"
               : "+r"(lo), "=r"(hi)
               : "r"(a), "r"(b));
}
#else
namespace hostsim {
inline uint32_t& cf() {
  static thread_local uint32_t f = 0;
  return f;
}
inline uint32_t adc(uint32_t a, uint32_t b, uint32_t cin, bool set) {
  uint64_t t = (uint64_t)a + b + cin;
  if (set) cf() = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
inline uint32_t sbb(uint32_t a, uint32_t b, uint32_t bin, bool set) {
  uint64_t t = (uint64_t)a - b - bin;
  if (set) cf() = (uint32_t)((t >> 32) & 1);
  return (uint32_t)t;
}
}  // namespace hostsim
inline uint32_t add_cc(uint32_t a, uint32_t b) { return hostsim::adc(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return hostsim::adc(a, b, hostsim::cf(), true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return hostsim::adc(a, b, hostsim::cf(), false); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return hostsim::sbb(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return hostsim::sbb(a, b, hostsim::cf(), true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return hostsim::sbb(a, b, hostsim::cf(), false); }
inline void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b;
  lo = (uint32_t)t;
  hi = (uint32_t)(t >> 32);
}
inline void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b;
  lo = hostsim::adc(lo, (uint32_t)t, 0, true);
  hi = hostsim::adc(hi, (uint32_t)(t >> 32), hostsim::cf(), true);
}
inline void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b;
  lo = hostsim::adc(lo, (uint32_t)t, hostsim::cf(), true);
  hi = hostsim::adc(hi, (uint32_t)(t >> 32), hostsim::cf(), true);
}
inline void madc_wide_top(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b;
  lo = hostsim::adc(lo, (uint32_t)t, hostsim::cf(), true);
  hi = hostsim::adc(0, (uint32_t)(t >> 32), hostsim::cf(), false);
}
#endif

// ---- reduction of a 512-bit product ----------------------------------------
// r[0..16) -> tight fe.  2^256 = 38, then 2^255 = 19 twice.
BPG_DI fe fe_reduce512(uint32_t r[16]) {
  // even chain: (r0,r1)+=38 r8, (r2,r3)+=38 r10, (r4,r5)+=38 r12, (r6,r7)+=38 r14
  uint32_t c_even, c_odd = 0, t8;
  mad_wide_cc(r[0], r[1], r[8], 38u);
  madc_wide_cc(r[2], r[3], r[10], 38u);
  madc_wide_cc(r[4], r[5], r[12], 38u);
  madc_wide_cc(r[6], r[7], r[14], 38u);
  c_even = addc(0u, 0u);
  // odd chain: (r1,r2)+=38 r9, (r3,r4)+=38 r11, (r5,r6)+=38 r13, (r7,t8)+=38 r15
  mad_wide_cc(r[1], r[2], r[9], 38u);
  madc_wide_cc(r[3], r[4], r[11], 38u);
  madc_wide_cc(r[5], r[6], r[13], 38u);
  madc_wide_top(r[7], t8, r[15], 38u);
  (void)c_odd;  // t8 <= 37 + 1: the odd chain cannot carry out
  // value = r[0..8) + 2^256 (t8 + c_even); fold at bit 255:
  uint32_t top = ((t8 + c_even) << 1) | (r[7] >> 31);  // < 2^8
  r[7] &= 0x7fffffffu;
  uint32_t f = top * 19u;
  fe o;
  o.v[0] = add_cc(r[0], f);
  o.v[1] = addc_cc(r[1], 0u);
  o.v[2] = addc_cc(r[2], 0u);
  o.v[3] = addc_cc(r[3], 0u);
  o.v[4] = addc_cc(r[4], 0u);
  o.v[5] = addc_cc(r[5], 0u);
  o.v[6] = addc_cc(r[6], 0u);
  o.v[7] = addc(r[7], 0u);
  // now < 2^255 + 19*2^8; if bit 255 is set the low part is < 2^13: add 19, no carry.
  uint32_t b = o.v[7] >> 31;
  o.v[7] &= 0x7fffffffu;
  o.v[0] += 19u * b;
  return o;
}

// ---- multiplication -----------------------------------------------------------
BPG_DI fe fe_mul_inl(const fe& A, const fe& B) {
  const uint32_t* a = A.v;
  const uint32_t* b = B.v;
  uint32_t e[16];  // even-aligned columns: pair (e[2k], e[2k+1]) = columns 2k, 2k+1
  uint32_t o[16];  // odd-aligned: o[k] = column k+1
  // row 0
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
      // a_odd * b_i -> even columns i+j, j = 1,3,5,7: e[i+1 .. i+8]
      mad_wide_cc(e[i + 1], e[i + 2], a[1], b[i]);
      madc_wide_cc(e[i + 3], e[i + 4], a[3], b[i]);
      madc_wide_cc(e[i + 5], e[i + 6], a[5], b[i]);
      if (i == 1) {
        e[i + 7] = 0;
        madc_wide_top(e[i + 7], e[i + 8], a[7], b[i]);
      } else {
        // e[i+7] holds the carry word of row i-1's even chain; e[i+8] fresh
        madc_wide_top(e[i + 7], e[i + 8], a[7], b[i]);
      }
      // a_even * b_i -> odd columns i+j, j = 0,2,4,6: o[i-1 .. i+6], carry -> o[i+7]
      mad_wide_cc(o[i - 1], o[i], a[0], b[i]);
      madc_wide_cc(o[i + 1], o[i + 2], a[2], b[i]);
      madc_wide_cc(o[i + 3], o[i + 4], a[4], b[i]);
      madc_wide_cc(o[i + 5], o[i + 6], a[6], b[i]);
      o[i + 7] = addc(0u, 0u);
    } else {
      // a_even * b_i -> even columns i+j, j = 0,2,4,6: e[i .. i+7], carry -> e[i+8]
      mad_wide_cc(e[i], e[i + 1], a[0], b[i]);
      madc_wide_cc(e[i + 2], e[i + 3], a[2], b[i]);
      madc_wide_cc(e[i + 4], e[i + 5], a[4], b[i]);
      madc_wide_cc(e[i + 6], e[i + 7], a[6], b[i]);
      e[i + 8] = addc(0u, 0u);
      // a_odd * b_i -> odd columns i+j, j = 1,3,5,7: o[i .. i+7]; o[i+6] is carry word
      mad_wide_cc(o[i], o[i + 1], a[1], b[i]);
      madc_wide_cc(o[i + 2], o[i + 3], a[3], b[i]);
      madc_wide_cc(o[i + 4], o[i + 5], a[5], b[i]);
      madc_wide_top(o[i + 6], o[i + 7], a[7], b[i]);
    }
  }
  // e covers columns 0..15 (e[15] written by row 7's top), o columns 1..15 (o[0..14]).
  uint32_t r[16];
  r[0] = e[0];
  r[1] = add_cc(e[1], o[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) r[k] = addc_cc(e[k], o[k - 1]);
  r[15] = addc(e[15], o[14]);
  return fe_reduce512(r);
}

// Dedicated squaring: the 28 cross products a_i a_j (i < j) in the same even/odd column chains,
// doubled once as a 512-bit shift, plus the 8 squares: 36 + 8 wide multiply-adds instead of 64 + 8.
// Measured on B200 (tools/fe_lat.cu): 207 ns against 262 ns for fe_mul(a, a) on a lone
// warp, 119 against 165 ns per warp at full occupancy.
BPG_DI fe fe_sq_inl(const fe& A) {
  const uint32_t* a = A.v;
  uint32_t e[16], o[16];
  e[0] = e[1] = 0;
  e[15] = 0;
  // i = 0: columns 1..7
  mul_wide(e[2], e[3], a[0], a[2]);
  mul_wide(e[4], e[5], a[0], a[4]);
  mul_wide(e[6], e[7], a[0], a[6]);
  mul_wide(o[0], o[1], a[0], a[1]);
  mul_wide(o[2], o[3], a[0], a[3]);
  mul_wide(o[4], o[5], a[0], a[5]);
  mul_wide(o[6], o[7], a[0], a[7]);
  // i = 1: columns 3..8
  mad_wide_cc(o[2], o[3], a[1], a[2]);
  madc_wide_cc(o[4], o[5], a[1], a[4]);
  madc_wide_cc(o[6], o[7], a[1], a[6]);
  o[8] = addc(0u, 0u);
  e[8] = 0;
  mad_wide_cc(e[4], e[5], a[1], a[3]);
  madc_wide_cc(e[6], e[7], a[1], a[5]);
  madc_wide_top(e[8], e[9], a[1], a[7]);
  // i = 2: columns 5..9
  o[9] = 0;
  mad_wide_cc(o[4], o[5], a[2], a[3]);
  madc_wide_cc(o[6], o[7], a[2], a[5]);
  madc_wide_cc(o[8], o[9], a[2], a[7]);
  o[10] = addc(0u, 0u);
  mad_wide_cc(e[6], e[7], a[2], a[4]);
  madc_wide_cc(e[8], e[9], a[2], a[6]);
  e[10] = addc(0u, 0u);
  // i = 3: columns 7..10
  mad_wide_cc(o[6], o[7], a[3], a[4]);
  madc_wide_cc(o[8], o[9], a[3], a[6]);
  o[10] = addc(o[10], 0u);
  e[11] = 0;
  mad_wide_cc(e[8], e[9], a[3], a[5]);
  madc_wide_cc(e[10], e[11], a[3], a[7]);
  e[12] = addc(0u, 0u);
  // i = 4: columns 9..11
  o[11] = 0;
  mad_wide_cc(o[8], o[9], a[4], a[5]);
  madc_wide_cc(o[10], o[11], a[4], a[7]);
  o[12] = addc(0u, 0u);
  mad_wide_cc(e[10], e[11], a[4], a[6]);
  e[12] = addc(e[12], 0u);
  // i = 5: columns 11, 12
  mad_wide_cc(o[10], o[11], a[5], a[6]);
  o[12] = addc(o[12], 0u);
  e[13] = 0;
  mad_wide_cc(e[12], e[13], a[5], a[7]);
  e[14] = addc(0u, 0u);
  // i = 6: column 13
  o[13] = 0;
  mad_wide_cc(o[12], o[13], a[6], a[7]);
  o[14] = addc(0u, 0u);
  // cross = e + (o << 32), then doubled
  uint32_t r[16];
  r[0] = 0;
  r[1] = o[0];
  r[2] = add_cc(e[2], o[1]);
#pragma unroll
  for (int k = 3; k < 15; k++) r[k] = addc_cc(e[k], o[k - 1]);
  r[15] = addc(e[15], o[14]);
#pragma unroll
  for (int k = 15; k >= 2; k--) r[k] = (r[k] << 1) | (r[k - 1] >> 31);
  r[1] <<= 1;
  // + squares
  uint32_t lo, hi;
  mul_wide(lo, hi, a[0], a[0]);
  r[0] = lo;
  r[1] = add_cc(r[1], hi);
#pragma unroll
  for (int i = 1; i < 8; i++) {
    mul_wide(lo, hi, a[i], a[i]);
    r[2 * i] = addc_cc(r[2 * i], lo);
    r[2 * i + 1] = addc_cc(r[2 * i + 1], hi);
  }
  return fe_reduce512(r);
}

// Out-of-line products for the latency-bound kernels (tree reductions, encodings, comb rounds): a handful of
// warps run hundreds of products each, and with every product inlined (about 1.5 KB of straight-line code)
// such a kernel is bound by instruction FETCH, not by the multiplier -- the measured cold cost of a finishing
// kernel was twice its arithmetic.  A translation unit defines BPG_FE_OUTLINE before including this header to
// get calls (operands by value travel in registers: no local memory); the bucket accumulation keeps them inline.
#if defined(BPG_FE_OUTLINE) && defined(__CUDA_ARCH__)
static __device__ __noinline__ fe fe_mul_call(fe a, fe b) { return fe_mul_inl(a, b); }
static __device__ __noinline__ fe fe_sq_call(fe a) { return fe_sq_inl(a); }
__device__ __forceinline__ fe fe_mul(const fe& a, const fe& b) { return fe_mul_call(a, b); }
__device__ __forceinline__ fe fe_sq(const fe& a) { return fe_sq_call(a); }
#else
BPG_DI fe fe_mul(const fe& a, const fe& b) { return fe_mul_inl(a, b); }
BPG_DI fe fe_sq(const fe& a) { return fe_sq_inl(a); }
#endif

// ---- addition / subtraction -------------------------------------------------
// loose + loose -> loose.  Carry out of 2^256 folds as +38; a second carry can
// only happen when the low words are within 38 of 2^256 and is folded again.
BPG_DI fe fe_add(const fe& a, const fe& b) {
  fe o;
  o.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) o.v[i] = addc_cc(a.v[i], b.v[i]);
  uint32_t c = addc(0u, 0u);
  uint32_t f = c * 38u;
  o.v[0] = add_cc(o.v[0], f);
#pragma unroll
  for (int i = 1; i < 8; i++) o.v[i] = addc_cc(o.v[i], 0u);
  c = addc(0u, 0u);
  o.v[0] += c * 38u;  // low word is < 38 here when c == 1: no further carry
  return o;
}

// tight + tight (both < 2^255): cannot carry.
BPG_DI fe fe_add_nc(const fe& a, const fe& b) {
  fe o;
  o.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) o.v[i] = addc_cc(a.v[i], b.v[i]);
  o.v[7] = addc(a.v[7], b.v[7]);
  return o;
}

// loose - loose -> loose.  A borrow folds as -38 (2^256 = 38); a second borrow
// can only happen when the wrapped value is < 38 and is folded again.
BPG_DI fe fe_sub(const fe& a, const fe& b) {
  fe o;
  o.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) o.v[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t bw = subc(0u, 0u);  // 0 or 0xffffffff
  uint32_t f = bw & 38u;
  o.v[0] = sub_cc(o.v[0], f);
#pragma unroll
  for (int i = 1; i < 8; i++) o.v[i] = subc_cc(o.v[i], 0u);
  bw = subc(0u, 0u);
  o.v[0] -= bw & 38u;  // value is >= 2^256 - 38 here when bw set: no further borrow
  return o;
}

BPG_DI fe fe_neg(const fe& a) {
  fe z;
#pragma unroll
  for (int i = 0; i < 8; i++) z.v[i] = 0;
  return fe_sub(z, a);
}

BPG_DI fe fe_zero() {
  fe z;
#pragma unroll
  for (int i = 0; i < 8; i++) z.v[i] = 0;
  return z;
}
BPG_DI fe fe_one() {
  fe z = fe_zero();
  z.v[0] = 1;
  return z;
}

// ---- canonical form -----------------------------------------------------------
// loose -> the unique representative in [0, p).
BPG_DI fe fe_canon(const fe& a) {
  // fold bit 255 twice -> < 2^255, then conditionally subtract p.
  fe o = a;
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    uint32_t t = o.v[7] >> 31;
    o.v[7] &= 0x7fffffffu;
    o.v[0] = add_cc(o.v[0], t * 19u);
#pragma unroll
    for (int i = 1; i < 7; i++) o.v[i] = addc_cc(o.v[i], 0u);
    o.v[7] = addc(o.v[7], 0u);
  }
  // o < 2^255 now (second pass cannot set bit 255 again unless low was tiny; handled by >= p test)
  // q = 1 iff o >= p  <=>  o + 19 >= 2^255
  uint32_t t0 = add_cc(o.v[0], 19u);
  uint32_t t;
#pragma unroll
  for (int i = 1; i < 7; i++) t = addc_cc(o.v[i], 0u);
  t = addc(o.v[7], 0u);
  (void)t0;
  uint32_t q = t >> 31;
  o.v[0] = add_cc(o.v[0], q * 19u);
#pragma unroll
  for (int i = 1; i < 7; i++) o.v[i] = addc_cc(o.v[i], 0u);
  o.v[7] = addc(o.v[7], 0u);
  o.v[7] &= 0x7fffffffu;
  return o;
}

BPG_DI bool fe_is_zero(const fe& a) {
  fe c = fe_canon(a);
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) x |= c.v[i];
  return x == 0;
}
BPG_DI bool fe_eq(const fe& a, const fe& b) { return fe_is_zero(fe_sub(a, b)); }
BPG_DI bool fe_is_neg(const fe& a) { return fe_canon(a).v[0] & 1u; }

BPG_DI fe fe_cneg(const fe& a, bool neg) {
  fe n = fe_neg(a);
  fe o;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = neg ? n.v[i] : a.v[i];
  return o;
}
// |a|: the canonical representative of a or -a whose low bit is clear (RFC 9496 §4.1 CT_ABS)
BPG_DI fe fe_abs(const fe& a) {
  fe c = fe_canon(a);
  return fe_canon(fe_cneg(c, c.v[0] & 1u));
}

// n squarings
BPG_DI fe fe_sqn(fe a, int n) {
  for (int i = 0; i < n; i++) a = fe_sq(a);
  return a;
}

// a^(2^252 - 3) = a^((p-5)/8)
BPG_DI fe fe_pow22523(const fe& z) {
  fe t0 = fe_sq(z);                 // 2
  fe t1 = fe_sqn(t0, 2);            // 8
  t1 = fe_mul(z, t1);               // 9
  t0 = fe_mul(t0, t1);              // 11
  t0 = fe_sq(t0);                   // 22
  t0 = fe_mul(t1, t0);              // 31 = 2^5-1
  t1 = fe_sqn(t0, 5);
  t0 = fe_mul(t1, t0);              // 2^10-1
  t1 = fe_sqn(t0, 10);
  t1 = fe_mul(t1, t0);              // 2^20-1
  fe t2 = fe_sqn(t1, 20);
  t1 = fe_mul(t2, t1);              // 2^40-1
  t1 = fe_sqn(t1, 10);
  t0 = fe_mul(t1, t0);              // 2^50-1
  t1 = fe_sqn(t0, 50);
  t1 = fe_mul(t1, t0);              // 2^100-1
  t2 = fe_sqn(t1, 100);
  t1 = fe_mul(t2, t1);              // 2^200-1
  t1 = fe_sqn(t1, 50);
  t0 = fe_mul(t1, t0);              // 2^250-1
  t0 = fe_sqn(t0, 2);               // 2^252-4
  return fe_mul(t0, z);             // 2^252-3
}

// a^(p-2)
BPG_DI fe fe_invert(const fe& z) {
  // z^(2^255-21) = (z^(2^252-3))^8 * z^3
  fe t = fe_pow22523(z);
  t = fe_sqn(t, 3);
  fe z3 = fe_mul(fe_sq(z), z);
  return fe_mul(t, z3);
}

BPG_DI void fe_load(fe& o, const uint32_t* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w;
  o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
}
BPG_DI void fe_store(uint32_t* p, const fe& o) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(o.v[0], o.v[1], o.v[2], o.v[3]);
  q[1] = make_uint4(o.v[4], o.v[5], o.v[6], o.v[7]);
}

}  // namespace bpg
