// ristretto255 group elements on edwards25519 (a = -1), extended coordinates,
// with affine-Niels (y+x, y-x, 2dxy) table entries for the 7-multiplication
// mixed addition the bucket kernels run.
//
// This is the group layer behind every `StarkPoint::msm*` call site of the
// reference; encode/decode follow RFC 9496 §4.3 so that
// results compare byte-for-byte with oracle/group.py.
#pragma once
#include "fe.cuh"

namespace bpg {

// Constants live in constant memory on the device; the host copies exist only for
// tests/hostsim (see fe.cuh).
#define BPG_DEF_CONST(name, ...)                                   \
  static __device__ __constant__ uint32_t name[8] = {__VA_ARGS__};        \
  static const uint32_t name##_h[8] = {__VA_ARGS__};
#if defined(__CUDA_ARCH__)
#define BPG_K(name) name
#else
#define BPG_K(name) name##_h
#endif

BPG_DEF_CONST(K_D, 0x135978a3u, 0x75eb4dcau, 0x4141d8abu, 0x00700a4du, 0x7779e898u, 0x8cc74079u, 0x2b6ffe73u,
              0x52036ceeu)
BPG_DEF_CONST(K_D2, 0x26b2f159u, 0xebd69b94u, 0x8283b156u, 0x00e0149au, 0xeef3d130u, 0x198e80f2u, 0x56dffce7u,
              0x2406d9dcu)
BPG_DEF_CONST(K_SQRT_M1, 0x4a0ea0b0u, 0xc4ee1b27u, 0xad2fe478u, 0x2f431806u, 0x3dfbd7a7u, 0x2b4d0099u,
              0x4fc1df0bu, 0x2b832480u)
BPG_DEF_CONST(K_INVSQRT_A_MINUS_D, 0x805d40eau, 0x99c8fdaau, 0x5a4172beu, 0x9d2f1617u, 0xfe01d840u,
              0x16c27b91u, 0xcfaffca2u, 0x786c8905u)

// RFC 9496 §4.1 constants of the Elligator map (element derivation, §4.3.4)
BPG_DEF_CONST(K_ONE_MINUS_D_SQ, 0x945fc176u, 0xe27c09c1u, 0xcd5e350fu, 0x2c81a138u, 0xbe70dfe4u, 0x9994abddu,
              0xb2b3e0d7u, 0x029072a8u)
BPG_DEF_CONST(K_D_MINUS_ONE_SQ, 0x44ed4d20u, 0x31ad5aaau, 0xb01e1999u, 0xd29e4a2cu, 0x529b4eebu, 0x4cdcd32fu,
              0xf66c2241u, 0x5968b37au)
BPG_DEF_CONST(K_SQRT_AD_MINUS_ONE, 0x497b2e1bu, 0x7e97f6a0u, 0x1b7854bdu, 0xaf9d8e0cu, 0x31f5d1fdu, 0x0f3cfcc9u,
              0x2b8348acu, 0x376931bfu)

BPG_DI fe fe_const(const uint32_t* k) {
  fe o;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = k[i];
  return o;
}

struct ge_ext {  // (X:Y:Z:T), T = XY/Z; all coordinates tight (< 2^255)
  fe X, Y, Z, T;
};
// Resident tables keep one affine-Niels entry (96 bytes of payload) per 128-byte line: an entry never straddles
// two lines, so a gathered entry costs one line of HBM traffic (unpadded entries cost 196 bytes on average,
// profiles/r1r_top_ncu_summary.json).  Comb tables (point_kernels.cuh, comb_kernels.cuh) stay packed.
constexpr int NIELS_WORDS = 32;
constexpr size_t NIELS_BYTES = 128;
struct ge_niels {  // affine: (y+x, y-x, 2dxy)
  fe ypx, ymx, t2d;
};

BPG_DI ge_ext ge_identity() {
  ge_ext r;
  r.X = fe_zero();
  r.Y = fe_one();
  r.Z = fe_one();
  r.T = fe_zero();
  return r;
}
BPG_DI ge_niels ge_niels_identity() {
  ge_niels r;
  r.ypx = fe_one();
  r.ymx = fe_one();
  r.t2d = fe_zero();
  return r;
}

// r = p + (neg ? -q : q), q affine-Niels.  7 fe_mul.  Unified/complete.
BPG_DI ge_ext ge_madd(const ge_ext& p, const ge_niels& q, bool neg) {
  fe ypx, ymx;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    ypx.v[i] = neg ? q.ymx.v[i] : q.ypx.v[i];
    ymx.v[i] = neg ? q.ypx.v[i] : q.ymx.v[i];
  }
  fe A = fe_mul(fe_sub(p.Y, p.X), ymx);
  fe B = fe_mul(fe_add_nc(p.Y, p.X), ypx);
  fe C = fe_mul(p.T, q.t2d);
  fe D = fe_add_nc(p.Z, p.Z);
  fe E = fe_sub(B, A);
  fe H = fe_add_nc(B, A);
  fe Fp = fe_sub(D, C);
  fe Gp = fe_add(D, C);
  fe F, G;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    F.v[i] = neg ? Gp.v[i] : Fp.v[i];
    G.v[i] = neg ? Fp.v[i] : Gp.v[i];
  }
  ge_ext r;
  r.X = fe_mul(E, F);
  r.Y = fe_mul(G, H);
  r.Z = fe_mul(F, G);
  r.T = fe_mul(E, H);
  return r;
}

// r = p + q, both extended.  9 fe_mul.
BPG_DI ge_ext ge_add(const ge_ext& p, const ge_ext& q) {
  fe A = fe_mul(fe_sub(p.Y, p.X), fe_sub(q.Y, q.X));
  fe B = fe_mul(fe_add_nc(p.Y, p.X), fe_add_nc(q.Y, q.X));
  fe C = fe_mul(fe_mul(p.T, q.T), fe_const(BPG_K(K_D2)));
  fe ZZ = fe_mul(p.Z, q.Z);
  fe D = fe_add_nc(ZZ, ZZ);
  fe E = fe_sub(B, A);
  fe H = fe_add_nc(B, A);
  fe F = fe_sub(D, C);
  fe G = fe_add(D, C);
  ge_ext r;
  r.X = fe_mul(E, F);
  r.Y = fe_mul(G, H);
  r.Z = fe_mul(F, G);
  r.T = fe_mul(E, H);
  return r;
}

BPG_DI ge_ext ge_neg(const ge_ext& p) {
  ge_ext r;
  r.X = fe_neg(p.X);
  r.Y = p.Y;
  r.Z = p.Z;
  r.T = fe_neg(p.T);
  // fe_neg returns loose; retighten is not needed for the formulas (they accept loose
  // operands in sub/mul) but ge_ext promises tight, so multiply-free fix: canon.
  r.X = fe_canon(r.X);
  r.T = fe_canon(r.T);
  return r;
}

// r = 2p.  4 squarings + 4 multiplications.
BPG_DI ge_ext ge_dbl(const ge_ext& p) {
  fe A = fe_sq(p.X);
  fe B = fe_sq(p.Y);
  fe ZZ = fe_sq(p.Z);
  fe C = fe_add_nc(ZZ, ZZ);
  fe AB = fe_add_nc(A, B);
  fe XY = fe_add_nc(p.X, p.Y);
  fe E = fe_sub(fe_sq(XY), AB);
  fe G = fe_sub(B, A);
  fe F = fe_sub(G, C);
  fe H = fe_neg(AB);
  ge_ext r;
  r.X = fe_mul(E, F);
  r.Y = fe_mul(G, H);
  r.Z = fe_mul(F, G);
  r.T = fe_mul(E, H);
  return r;
}

// extended (any Z) -> affine Niels.  One inversion; callers on bulk paths batch it.
BPG_DI ge_niels ge_to_niels(const ge_ext& p) {
  fe zi = fe_invert(p.Z);
  fe x = fe_mul(p.X, zi);
  fe y = fe_mul(p.Y, zi);
  ge_niels r;
  r.ypx = fe_mul(fe_add_nc(y, x), fe_one());  // tight
  r.ymx = fe_mul(fe_sub(y, x), fe_one());
  r.t2d = fe_mul(fe_mul(x, y), fe_const(BPG_K(K_D2)));
  return r;
}

// affine (x, y) -> Niels
BPG_DI ge_niels ge_affine_to_niels(const fe& x, const fe& y) {
  ge_niels r;
  r.ypx = fe_canon(fe_add(y, x));
  r.ymx = fe_canon(fe_sub(y, x));
  r.t2d = fe_mul(fe_mul(x, y), fe_const(BPG_K(K_D2)));
  return r;
}

// RFC 9496 §4.2: (was_square, r) with r = sqrt(u/v) or sqrt(i u/v), r non-negative.
BPG_DI bool fe_sqrt_ratio_m1(fe& r_out, const fe& u, const fe& v) {
  fe v3 = fe_mul(fe_sq(v), v);
  fe v7 = fe_mul(fe_sq(v3), v);
  fe r = fe_mul(fe_mul(u, v3), fe_pow22523(fe_mul(u, v7)));
  fe check = fe_mul(v, fe_sq(r));
  fe i = fe_const(BPG_K(K_SQRT_M1));
  fe neg_u = fe_neg(u);
  bool correct = fe_eq(check, u);
  bool flipped = fe_eq(check, neg_u);
  bool flipped_i = fe_eq(check, fe_mul(neg_u, i));
  fe r_prime = fe_mul(r, i);
  bool flip = flipped | flipped_i;
#pragma unroll
  for (int k = 0; k < 8; k++) r.v[k] = flip ? r_prime.v[k] : r.v[k];
  r_out = fe_abs(r);
  return correct | flipped;
}

BPG_DI void fe_to_bytes(uint8_t* out, const fe& a) {
  fe c = fe_canon(a);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    out[4 * i + 0] = (uint8_t)(c.v[i]);
    out[4 * i + 1] = (uint8_t)(c.v[i] >> 8);
    out[4 * i + 2] = (uint8_t)(c.v[i] >> 16);
    out[4 * i + 3] = (uint8_t)(c.v[i] >> 24);
  }
}

// RFC 9496 §4.3.2
BPG_DI void ge_encode(uint8_t out[32], const ge_ext& p) {
  fe u1 = fe_mul(fe_add_nc(p.Z, p.Y), fe_sub(p.Z, p.Y));
  fe u2 = fe_mul(p.X, p.Y);
  fe invsqrt;
  fe_sqrt_ratio_m1(invsqrt, fe_one(), fe_mul(u1, fe_sq(u2)));
  fe den1 = fe_mul(invsqrt, u1);
  fe den2 = fe_mul(invsqrt, u2);
  fe z_inv = fe_mul(fe_mul(den1, den2), p.T);
  fe i = fe_const(BPG_K(K_SQRT_M1));
  fe ix0 = fe_mul(p.X, i);
  fe iy0 = fe_mul(p.Y, i);
  fe enchanted = fe_mul(den1, fe_const(BPG_K(K_INVSQRT_A_MINUS_D)));
  bool rotate = fe_is_neg(fe_mul(p.T, z_inv));
  fe x, y, den_inv;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    x.v[k] = rotate ? iy0.v[k] : p.X.v[k];
    y.v[k] = rotate ? ix0.v[k] : p.Y.v[k];
    den_inv.v[k] = rotate ? enchanted.v[k] : den2.v[k];
  }
  y = fe_cneg(y, fe_is_neg(fe_mul(x, z_inv)));
  fe s = fe_abs(fe_mul(den_inv, fe_sub(p.Z, y)));
  fe_to_bytes(out, s);
}

// RFC 9496 §4.3.1.  Returns false on a non-canonical / invalid encoding.
BPG_DI bool ge_decode(ge_ext& r, const uint8_t in[32]) {
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
  fe one = fe_one();
  fe ss = fe_sq(s);
  fe u1 = fe_sub(one, ss);
  fe u2 = fe_add(one, ss);
  fe u2_sqr = fe_sq(u2);
  fe v = fe_sub(fe_neg(fe_mul(fe_const(BPG_K(K_D)), fe_sq(u1))), u2_sqr);
  fe invsqrt;
  bool was_square = fe_sqrt_ratio_m1(invsqrt, one, fe_mul(v, u2_sqr));
  fe den_x = fe_mul(invsqrt, u2);
  fe den_y = fe_mul(fe_mul(invsqrt, den_x), v);
  fe x = fe_abs(fe_mul(fe_add(s, s), den_x));
  fe y = fe_mul(u1, den_y);
  fe t = fe_mul(x, y);
  ok = ok & was_square & !fe_is_neg(t) & !fe_is_zero(y);
  r.X = x;
  r.Y = y;
  r.Z = one;
  r.T = t;
  return ok;
}

// RFC 9496 §4.3.4 MAP (Elligator 2): field element -> point of the prime-order group.
BPG_DI ge_ext ge_elligator_map(const fe& t) {
  fe one = fe_one();
  fe d = fe_const(BPG_K(K_D));
  fe r = fe_mul(fe_const(BPG_K(K_SQRT_M1)), fe_sq(t));
  fe u = fe_mul(fe_add(r, one), fe_const(BPG_K(K_ONE_MINUS_D_SQ)));
  fe v = fe_mul(fe_sub(fe_neg(one), fe_mul(r, d)), fe_add(r, d));
  fe s;
  bool was_square = fe_sqrt_ratio_m1(s, u, v);
  fe s_prime = fe_neg(fe_abs(fe_mul(s, t)));
  fe c = fe_neg(one);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    s.v[k] = was_square ? s.v[k] : s_prime.v[k];
    c.v[k] = was_square ? c.v[k] : r.v[k];
  }
  fe N = fe_sub(fe_mul(fe_mul(c, fe_sub(r, one)), fe_const(BPG_K(K_D_MINUS_ONE_SQ))), v);
  fe ss = fe_sq(s);
  fe w0 = fe_mul(fe_add(s, s), v);
  fe w1 = fe_mul(N, fe_const(BPG_K(K_SQRT_AD_MINUS_ONE)));
  fe w2 = fe_sub(one, ss);
  fe w3 = fe_add(one, ss);
  ge_ext q;
  q.X = fe_mul(w0, w3);
  q.Y = fe_mul(w2, w1);
  q.Z = fe_mul(w1, w3);
  q.T = fe_mul(w0, w2);
  return q;
}

// RFC 9496 §4.3.4 element derivation: 64 uniform bytes -> MAP(lo mod 2^255) + MAP(hi mod 2^255)
// (what dalek's RistrettoPoint::from_uniform_bytes computes for the generator chains,
// reference src/generators.rs:107-125 in its ristretto255 form).
BPG_DI fe fe_from_bytes_255(const uint8_t in[32]) {
  fe t;
#pragma unroll
  for (int i = 0; i < 8; i++)
    t.v[i] = (uint32_t)in[4 * i] | ((uint32_t)in[4 * i + 1] << 8) | ((uint32_t)in[4 * i + 2] << 16) |
             ((uint32_t)in[4 * i + 3] << 24);
  t.v[7] &= 0x7fffffffu;
  return t;
}

BPG_DI void ge_store_niels(uint32_t* p, const ge_niels& q) {
  fe_store(p, q.ypx);
  fe_store(p + 8, q.ymx);
  fe_store(p + 16, q.t2d);
}
BPG_DI void ge_load_niels(ge_niels& q, const uint32_t* p) {
  fe_load(q.ypx, p);
  fe_load(q.ymx, p + 8);
  fe_load(q.t2d, p + 16);
}
BPG_DI void ge_store_ext(uint32_t* p, const ge_ext& q) {
  fe_store(p, q.X);
  fe_store(p + 8, q.Y);
  fe_store(p + 16, q.Z);
  fe_store(p + 24, q.T);
}
BPG_DI void ge_load_ext(ge_ext& q, const uint32_t* p) {
  fe_load(q.X, p);
  fe_load(q.Y, p + 8);
  fe_load(q.Z, p + 16);
  fe_load(q.T, p + 24);
}

BPG_DEF_CONST(K_DINV, 0xcdc9f843u, 0x25e0f276u, 0x4279542eu, 0x0b5dd698u, 0xcdb9cf66u, 0x2b162114u, 0x14d5ce43u,
              0x40907ed2u)  // 1/d

// the point (+-) of a Niels entry as an extended point with Z = 2: one multiplication
BPG_DI ge_ext ge_from_niels(const ge_niels& q, bool neg) {
  ge_ext r;
  fe x2 = fe_sub(q.ypx, q.ymx);   // 2x
  fe t = fe_mul(q.t2d, fe_const(BPG_K(K_DINV)));  // 2xy
  r.X = fe_canon(fe_cneg(x2, neg));
  r.Y = fe_canon(fe_add(q.ypx, q.ymx));  // 2y
  r.Z = fe_zero();
  r.Z.v[0] = 2;
  r.T = fe_canon(fe_cneg(t, neg));
  return r;
}


// p + q with q in the cached layout (Y-X, Y+X, 2Z, 2dT): 8 multiplications
BPG_DI ge_ext ge_add_cached(const ge_ext& p, const fe& ymx, const fe& ypx, const fe& z2,
                                                const fe& t2d) {
  fe A = fe_mul(fe_sub(p.Y, p.X), ymx);
  fe B = fe_mul(fe_add_nc(p.Y, p.X), ypx);
  fe C = fe_mul(p.T, t2d);
  fe D = fe_mul(p.Z, z2);
  fe E = fe_sub(B, A), H = fe_add_nc(B, A), F = fe_sub(D, C), G = fe_add(D, C);
  ge_ext r;
  r.X = fe_mul(E, F);
  r.Y = fe_mul(G, H);
  r.Z = fe_mul(F, G);
  r.T = fe_mul(E, H);
  return r;
}

}  // namespace bpg
