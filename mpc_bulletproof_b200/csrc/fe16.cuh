// GF(2^255 - 19) spread over SIXTEEN LANES: the latency form of the field layer.
//
// Why: the Ristretto encoding that ends every MSM (reference: `CompressedRistretto` bytes of the
// `msm_iter` result fed to the transcript, src/inner_product_proof.rs:116-122, 174-180) is one
// inverse square root = 252 dependent squarings + ~40 multiplications on ONE point.  A lone thread
// pays ~207 ns per squaring (fe.cuh: 44 chained IMAD.WIDE), i.e. 80-110 us per encoding, once per
// inner-product round.  Nothing else can run meanwhile (the challenge depends on the bytes), so
// the only way to shorten it is to shorten the product itself.
//
// Representation: radix 2^16, limb k on lane k of a half-warp, LAZY (limbs above 16 bits are
// allowed).  A product is
//     T_k = sum_i a_i * b'_(k-i),   b'_j = b_j (j >= 0),  38 * b_(j+16) (j < 0)      (2^256 = 38)
// sixteen independent 32x32->64 multiply-adds per lane with NO carry chain (T_k < 2^45 fits a
// 64-bit accumulator with 19 bits to spare), operands exchanged through shared memory (a: four
// broadcast LDS.128; b': sixteen conflict-free LDS.32 at lane-rotated addresses), followed by ONE
// carry step  l_k = T_k[0:16] + T_(k-1)[16:32] + T_(k-2)[32:]  (two shuffles; the wrap-around
// terms times 38).  Steady-state bounds (tests/test_hostsim.py checks them): limb 0 < 2^22,
// limbs 1..15 < 2^18, T < 2^45.  Sums and differences take one extra carry step so that every
// product operand obeys those bounds.
//
// Host build (tests/hostsim): the same code, one host thread per lane, shuffles and the
// shared-memory exchange modelled with a barrier.
#pragma once
#include "ge.cuh"

#if defined(BPG_HOSTSIM)
#include <pthread.h>
#endif

namespace bpg {

constexpr int G16_WORDS = 112;  // per-group exchange buffer: 2 x 48 words, padded so that the two
                                // half-warps of a warp start 16 banks apart

struct grp16 {
  uint32_t* sm;   // this half-warp's exchange buffer (G16_WORDS words, 16-byte aligned)
  uint32_t k;     // lane within the group (limb index)
  uint32_t par;   // which half of the buffer the next exchange uses
  uint32_t half;  // WIDE mode only: which half-warp this lane is in (lane >> 4)
#if defined(BPG_HOSTSIM)  // tests/hostsim only
  pthread_barrier_t* bar;
  uint32_t* slots;  // 16 words for the shuffle model
#endif
};

#if defined(__CUDA_ARCH__)
// Every lane of the warp executes every call (idle groups shadow a live one), so the full mask.
__device__ __forceinline__ void g16_sync(const grp16&) { __syncwarp(); }
__device__ __forceinline__ uint32_t g16_shfl(const grp16&, uint32_t v, uint32_t src) {
  return __shfl_sync(0xffffffffu, v, (int)src, 16);
}
// WIDE mode: the value held by the same limb's lane in the other half-warp
__device__ __forceinline__ uint64_t g16_other_half(const grp16&, uint64_t v) {
  uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, 16);
  uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), 16);
  return ((uint64_t)hi << 32) | lo;
}
#elif defined(BPG_HOSTSIM)
// slots: 2 x 32 words; a half-warp of the model uses its own 16 (half = 0 when only 16 threads run)
inline void g16_sync(const grp16& g) { pthread_barrier_wait(g.bar); }
inline uint32_t g16_shfl(const grp16& g, uint32_t v, uint32_t src) {
  g.slots[g.half * 16 + g.k] = v;
  pthread_barrier_wait(g.bar);
  uint32_t r = g.slots[g.half * 16 + src];
  pthread_barrier_wait(g.bar);
  return r;
}
inline uint64_t g16_other_half(const grp16& g, uint64_t v) {
  g.slots[g.half * 16 + g.k] = (uint32_t)v;
  g.slots[32 + g.half * 16 + g.k] = (uint32_t)(v >> 32);
  pthread_barrier_wait(g.bar);
  uint64_t r = ((uint64_t)g.slots[32 + (g.half ^ 1) * 16 + g.k] << 32) | g.slots[(g.half ^ 1) * 16 + g.k];
  pthread_barrier_wait(g.bar);
  return r;
}
#else  // host pass of the product build: never called
inline void g16_sync(const grp16&) {}
inline uint32_t g16_shfl(const grp16&, uint32_t v, uint32_t) { return v; }
inline uint64_t g16_other_half(const grp16&, uint64_t v) { return v; }
#endif

struct fe16 {
  uint32_t l;  // this lane's limb
};

BPG_DI uint32_t* g16_next_buf(grp16& g) {
  uint32_t* b = g.sm + g.par * 48;
  g.par ^= 1u;
  return b;
}

// one carry step: limbs < 2^26 -> limbs < 2^17 (limb 0 < 2^16 + 38 * 2^10)
BPG_DI fe16 fe16_carry(grp16& g, fe16 a) {
  uint32_t c = g16_shfl(g, a.l >> 16, (g.k + 15u) & 15u);
  if (g.k == 0) c *= 38u;
  fe16 r;
  r.l = (a.l & 0xffffu) + c;
  return r;
}

// a * b.  Operands: limb 0 < 2^23, others < 2^19 (any product, carried sum or carried difference).
// WIDE: both half-warps work on the SAME element (same exchange buffer, same limb on lanes k and
// k + 16); each half forms eight of the sixteen column terms and the halves swap their partial
// columns, so a product is 8 multiply-adds + 10 loads per lane instead of 16 + 20.
template <bool WIDE>
BPG_DI fe16 fe16_mul(grp16& g, fe16 a, fe16 b) {
  uint32_t* buf = g16_next_buf(g);
  buf[g.k] = 38u * b.l;
  buf[16 + g.k] = b.l;
  buf[32 + g.k] = a.l;
  g16_sync(g);
  uint64_t T;
  if (WIDE) {
    const uint32_t hb = g.half * 8u;
    const uint32_t* rot = buf + 16 + g.k - hb;  // rot[-j] = b'_(k - hb - j)
    const uint32_t* av = buf + 32 + hb;
    uint64_t t0 = 0, t1 = 0;
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      uint4 x = *reinterpret_cast<const uint4*>(av + j);
      t0 += (uint64_t)x.x * rot[-j];
      t1 += (uint64_t)x.y * rot[-j - 1];
      t0 += (uint64_t)x.z * rot[-j - 2];
      t1 += (uint64_t)x.w * rot[-j - 3];
    }
    uint64_t part = t0 + t1;
    T = part + g16_other_half(g, part);
  } else {
    const uint32_t* rot = buf + 16 + g.k;  // rot[-i] = b'_(k-i)
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      uint4 x = *reinterpret_cast<const uint4*>(buf + 32 + i);
      t0 += (uint64_t)x.x * rot[-i];
      t1 += (uint64_t)x.y * rot[-i - 1];
      t2 += (uint64_t)x.z * rot[-i - 2];
      t3 += (uint64_t)x.w * rot[-i - 3];
    }
    T = (t0 + t1) + (t2 + t3);
  }
  uint32_t lo = (uint32_t)T;
  uint32_t s1 = g16_shfl(g, lo >> 16, (g.k + 15u) & 15u);
  uint32_t s2 = g16_shfl(g, (uint32_t)(T >> 32), (g.k + 14u) & 15u);
  if (g.k < 1) s1 *= 38u;
  if (g.k < 2) s2 *= 38u;
  fe16 r;
  r.l = (lo & 0xffffu) + s1 + s2;
  return r;
}
template <bool WIDE>
BPG_DI fe16 fe16_sq(grp16& g, fe16 a) { return fe16_mul<WIDE>(g, a, a); }
template <bool WIDE>
BPG_DI fe16 fe16_sqn(grp16& g, fe16 a, int n) {
  for (int i = 0; i < n; i++) a = fe16_sq<WIDE>(g, a);
  return a;
}

BPG_DI fe16 fe16_add(grp16& g, fe16 a, fe16 b) {
  fe16 r;
  r.l = a.l + b.l;
  return fe16_carry(g, r);
}
// a - b + 256 p: the limbs of 256 p (256 * 0xffed, 256 * 0xffff ..., 256 * 0x7fff) exceed any
// operand limb (< 2^23), so no lane goes negative.
BPG_DI fe16 fe16_sub(grp16& g, fe16 a, fe16 b) {
  uint32_t bias = g.k == 0 ? 256u * 0xffedu : (g.k == 15 ? 256u * 0x7fffu : 256u * 0xffffu);
  fe16 r;
  r.l = a.l + bias - b.l;
  return fe16_carry(g, r);
}
BPG_DI fe16 fe16_sel(bool c, fe16 a, fe16 b) {
  fe16 r;
  r.l = c ? a.l : b.l;
  return r;
}

// the group's copy of a replicated (lane-uniform) field element
BPG_DI fe16 fe16_from_fe(const grp16& g, const fe& a) {
  uint32_t w = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) w = (g.k >> 1) == (uint32_t)i ? a.v[i] : w;
  fe16 r;
  r.l = (g.k & 1u) ? (w >> 16) : (w & 0xffffu);
  return r;
}

// canonical value, replicated in every lane of the group
BPG_DI fe fe16_to_fe(grp16& g, fe16 a) {
  uint32_t* buf = g16_next_buf(g);
  buf[g.k] = a.l;
  g16_sync(g);
  uint64_t acc = 0;
  fe r;
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    uint4 w = *reinterpret_cast<const uint4*>(buf + i);
    acc += w.x;
    uint32_t h0 = (uint32_t)acc & 0xffffu;
    acc >>= 16;
    acc += w.y;
    r.v[i >> 1] = h0 | ((uint32_t)acc << 16);
    acc >>= 16;
    acc += w.z;
    uint32_t h2 = (uint32_t)acc & 0xffffu;
    acc >>= 16;
    acc += w.w;
    r.v[(i >> 1) + 1] = h2 | ((uint32_t)acc << 16);
    acc >>= 16;
  }
  fe top = fe_zero();
  top.v[0] = 38u * (uint32_t)acc;  // what left 2^256
  return fe_canon(fe_add(r, top));
}

// a^(2^252 - 3), the chain of fe_pow22523
template <bool WIDE>
BPG_DI fe16 fe16_pow22523(grp16& g, fe16 z) {
  fe16 t0 = fe16_sq<WIDE>(g, z);
  fe16 t1 = fe16_sqn<WIDE>(g, t0, 2);
  t1 = fe16_mul<WIDE>(g, z, t1);
  t0 = fe16_mul<WIDE>(g, t0, t1);
  t0 = fe16_sq<WIDE>(g, t0);
  t0 = fe16_mul<WIDE>(g, t1, t0);
  t1 = fe16_sqn<WIDE>(g, t0, 5);
  t0 = fe16_mul<WIDE>(g, t1, t0);
  t1 = fe16_sqn<WIDE>(g, t0, 10);
  t1 = fe16_mul<WIDE>(g, t1, t0);
  fe16 t2 = fe16_sqn<WIDE>(g, t1, 20);
  t1 = fe16_mul<WIDE>(g, t2, t1);
  t1 = fe16_sqn<WIDE>(g, t1, 10);
  t0 = fe16_mul<WIDE>(g, t1, t0);
  t1 = fe16_sqn<WIDE>(g, t0, 50);
  t1 = fe16_mul<WIDE>(g, t1, t0);
  t2 = fe16_sqn<WIDE>(g, t1, 100);
  t1 = fe16_mul<WIDE>(g, t2, t1);
  t1 = fe16_sqn<WIDE>(g, t1, 50);
  t0 = fe16_mul<WIDE>(g, t1, t0);
  t0 = fe16_sqn<WIDE>(g, t0, 2);
  return fe16_mul<WIDE>(g, t0, z);
}

// RFC 9496 §4.3.2 on a half-warp: the same steps as ge_encode (ge.cuh), hence the same bytes.
// `p` is replicated in the group's lanes; the canonical s comes back replicated.
template <bool WIDE>
BPG_DI fe ge_encode16(grp16& g, const ge_ext& p) {
  fe16 X = fe16_from_fe(g, p.X), Y = fe16_from_fe(g, p.Y), Z = fe16_from_fe(g, p.Z), T = fe16_from_fe(g, p.T);
  fe16 I = fe16_from_fe(g, fe_const(BPG_K(K_SQRT_M1)));
  fe16 zero;
  zero.l = 0;
  fe16 u1 = fe16_mul<WIDE>(g, fe16_add(g, Z, Y), fe16_sub(g, Z, Y));
  fe16 u2 = fe16_mul<WIDE>(g, X, Y);
  // invsqrt = sqrt_ratio_m1(1, v), v = u1 u2^2
  fe16 v = fe16_mul<WIDE>(g, u1, fe16_sq<WIDE>(g, u2));
  fe16 v3 = fe16_mul<WIDE>(g, fe16_sq<WIDE>(g, v), v);
  fe16 v7 = fe16_mul<WIDE>(g, fe16_sq<WIDE>(g, v3), v);
  fe16 r = fe16_mul<WIDE>(g, v3, fe16_pow22523<WIDE>(g, v7));
  fe check = fe16_to_fe(g, fe16_mul<WIDE>(g, v, fe16_sq<WIDE>(g, r)));
  // check against u = 1: -u = p - 1, -u i = -sqrt(-1)
  fe m1 = fe_canon(fe_neg(fe_one()));
  fe mi = fe_canon(fe_neg(fe_const(BPG_K(K_SQRT_M1))));
  uint32_t d1 = 0, di = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    d1 |= check.v[i] ^ m1.v[i];
    di |= check.v[i] ^ mi.v[i];
  }
  bool flip = (d1 == 0) | (di == 0);
  r = fe16_sel(flip, fe16_mul<WIDE>(g, r, I), r);
  bool rneg = fe16_to_fe(g, r).v[0] & 1u;
  fe16 invsqrt = fe16_sel(rneg, fe16_sub(g, zero, r), r);
  fe16 den1 = fe16_mul<WIDE>(g, invsqrt, u1);
  fe16 den2 = fe16_mul<WIDE>(g, invsqrt, u2);
  fe16 z_inv = fe16_mul<WIDE>(g, fe16_mul<WIDE>(g, den1, den2), T);
  fe16 ix0 = fe16_mul<WIDE>(g, X, I);
  fe16 iy0 = fe16_mul<WIDE>(g, Y, I);
  fe16 enchanted = fe16_mul<WIDE>(g, den1, fe16_from_fe(g, fe_const(BPG_K(K_INVSQRT_A_MINUS_D))));
  bool rotate = fe16_to_fe(g, fe16_mul<WIDE>(g, T, z_inv)).v[0] & 1u;
  fe16 x = fe16_sel(rotate, iy0, X);
  fe16 y = fe16_sel(rotate, ix0, Y);
  fe16 den_inv = fe16_sel(rotate, enchanted, den2);
  bool yneg = fe16_to_fe(g, fe16_mul<WIDE>(g, x, z_inv)).v[0] & 1u;
  y = fe16_sel(yneg, fe16_sub(g, zero, y), y);
  fe16 s = fe16_mul<WIDE>(g, den_inv, fe16_sub(g, Z, y));
  fe sc = fe16_to_fe(g, s);
  return fe_canon(fe_cneg(sc, sc.v[0] & 1u));
}

// RFC 9496 §4.3.1 on a half-warp (or a warp, WIDE): the same steps as ge_decode (ge.cuh), hence the same point and
// the same verdict.  `in` is replicated in the group's lanes; the extended point (Z = 1) comes back replicated.
template <bool WIDE>
BPG_DI bool ge_decode16(grp16& g, const uint8_t in[32], ge_ext& out) {
  fe s;
#pragma unroll
  for (int i = 0; i < 8; i++)
    s.v[i] = (uint32_t)in[4 * i] | ((uint32_t)in[4 * i + 1] << 8) | ((uint32_t)in[4 * i + 2] << 16) |
             ((uint32_t)in[4 * i + 3] << 24);
  // canonical (s < p) and non-negative
  fe c = fe_canon(s);
  uint32_t diff = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) diff |= c.v[i] ^ s.v[i];
  bool ok = (diff == 0) & ((s.v[0] & 1u) == 0);
  s = c;  // the arithmetic below wants a reduced operand; a non-canonical input is rejected anyway
  fe16 zero;
  zero.l = 0;
  const fe16 one = fe16_from_fe(g, fe_one());
  const fe16 I = fe16_from_fe(g, fe_const(BPG_K(K_SQRT_M1)));
  const fe16 s16 = fe16_from_fe(g, s);
  fe16 ss = fe16_sq<WIDE>(g, s16);
  fe16 u1 = fe16_sub(g, one, ss);
  fe16 u2 = fe16_add(g, one, ss);
  fe16 u2_sqr = fe16_sq<WIDE>(g, u2);
  fe16 du1 = fe16_mul<WIDE>(g, fe16_from_fe(g, fe_const(BPG_K(K_D))), fe16_sq<WIDE>(g, u1));
  fe16 v = fe16_sub(g, fe16_sub(g, zero, du1), u2_sqr);
  // (was_square, invsqrt) = sqrt_ratio_m1(1, w), w = v u2^2
  fe16 w = fe16_mul<WIDE>(g, v, u2_sqr);
  fe16 w3 = fe16_mul<WIDE>(g, fe16_sq<WIDE>(g, w), w);
  fe16 w7 = fe16_mul<WIDE>(g, fe16_sq<WIDE>(g, w3), w);
  fe16 r = fe16_mul<WIDE>(g, w3, fe16_pow22523<WIDE>(g, w7));
  fe check = fe16_to_fe(g, fe16_mul<WIDE>(g, w, fe16_sq<WIDE>(g, r)));
  const fe p1 = fe_one(), m1 = fe_canon(fe_neg(fe_one())), mi = fe_canon(fe_neg(fe_const(BPG_K(K_SQRT_M1))));
  uint32_t d0 = 0, d1 = 0, di = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    d0 |= check.v[i] ^ p1.v[i];
    d1 |= check.v[i] ^ m1.v[i];
    di |= check.v[i] ^ mi.v[i];
  }
  const bool correct = d0 == 0, flipped = d1 == 0, flipped_i = di == 0;
  r = fe16_sel(flipped | flipped_i, fe16_mul<WIDE>(g, r, I), r);
  const bool rneg = fe16_to_fe(g, r).v[0] & 1u;
  fe16 invsqrt = fe16_sel(rneg, fe16_sub(g, zero, r), r);
  const bool was_square = correct | flipped;
  fe16 den_x = fe16_mul<WIDE>(g, invsqrt, u2);
  fe16 den_y = fe16_mul<WIDE>(g, fe16_mul<WIDE>(g, invsqrt, den_x), v);
  fe16 x = fe16_mul<WIDE>(g, fe16_add(g, s16, s16), den_x);
  const bool xneg = fe16_to_fe(g, x).v[0] & 1u;
  x = fe16_sel(xneg, fe16_sub(g, zero, x), x);  // abs
  fe16 y = fe16_mul<WIDE>(g, u1, den_y);
  fe16 t = fe16_mul<WIDE>(g, x, y);
  out.X = fe16_to_fe(g, x);
  out.Y = fe16_to_fe(g, y);
  out.Z = fe_one();
  out.T = fe16_to_fe(g, t);
  uint32_t ynz = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) ynz |= out.Y.v[i];
  return ok & was_square & ((out.T.v[0] & 1u) == 0) & (ynz != 0);
}

}  // namespace bpg
