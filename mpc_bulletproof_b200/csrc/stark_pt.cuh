// The Stark-curve policy, group layer:  y^2 = x^3 + x + beta  over F_p, prime order.
//
// Table entries are affine (x, y) in Montgomery form, 64 bytes, the identity flagged as (0, 0)
// (not on the curve: beta != 0).  Accumulators are extended Jacobian "XYZZ" points
// (X, Y, ZZ, ZZZ) with x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; the identity has ZZ = 0.  Mixed
// addition is 8M + 2S, full addition 12M + 2S, doubling 6M + 4S (Explicit-Formulas Database,
// "madd-2008-s", "add-2008-s", "dbl-2008-s-1", "mdbl-2008-s-1").  Short-Weierstrass formulas are
// not unified: P = Q falls through to a doubling and P = -Q to the identity, both handled here
// (a bucket can meet the same point twice when the input repeats a point).
#pragma once
#include "stark_fp.cuh"

namespace bpg {

struct sp_aff {  // affine, Montgomery; (0, 0) = identity
  fp x, y;
};
struct sp_xyzz {
  fp X, Y, ZZ, ZZZ;
};

BPG_DI sp_xyzz sp_identity() {
  sp_xyzz r;
  r.X = fp_zero();
  r.Y = fp_zero();
  r.ZZ = fp_zero();
  r.ZZZ = fp_zero();
  return r;
}
BPG_DI bool sp_is_identity(const sp_xyzz& p) { return fp_is_zero(p.ZZ); }
BPG_DI bool sp_aff_is_identity(const sp_aff& q) { return fp_is_zero(q.x) & fp_is_zero(q.y); }

BPG_DI sp_xyzz sp_from_aff(const sp_aff& q) {
  sp_xyzz r;
  bool id = sp_aff_is_identity(q);
  r.X = q.x;
  r.Y = q.y;
  r.ZZ = id ? fp_zero() : fp_one();
  r.ZZZ = r.ZZ;
  return r;
}

// 2 * q for an affine q (mdbl-2008-s-1; a = 1)
BPG_DI sp_xyzz sp_mdbl(const sp_aff& q) {
  if (sp_aff_is_identity(q) || fp_is_zero(q.y)) return sp_identity();
  fp U = fp_dbl(q.y);
  fp V = fp_sq(U);
  fp W = fp_mul(U, V);
  fp S = fp_mul(q.x, V);
  fp xx = fp_sq(q.x);
  fp M = fp_add(fp_add(fp_dbl(xx), xx), fp_one());
  sp_xyzz r;
  r.X = fp_sub(fp_sq(M), fp_dbl(S));
  r.Y = fp_sub(fp_mul(M, fp_sub(S, r.X)), fp_mul(W, q.y));
  r.ZZ = V;
  r.ZZZ = W;
  return r;
}

// 2 * p (dbl-2008-s-1; a = 1)
BPG_DI sp_xyzz sp_dbl(const sp_xyzz& p) {
  if (sp_is_identity(p) || fp_is_zero(p.Y)) return sp_identity();
  fp U = fp_dbl(p.Y);
  fp V = fp_sq(U);
  fp W = fp_mul(U, V);
  fp S = fp_mul(p.X, V);
  fp xx = fp_sq(p.X);
  fp M = fp_add(fp_add(fp_dbl(xx), xx), fp_sq(p.ZZ));
  sp_xyzz r;
  r.X = fp_sub(fp_sq(M), fp_dbl(S));
  r.Y = fp_sub(fp_mul(M, fp_sub(S, r.X)), fp_mul(W, p.Y));
  r.ZZ = fp_mul(V, p.ZZ);
  r.ZZZ = fp_mul(W, p.ZZZ);
  return r;
}

// p + (neg ? -q : q), q affine (madd-2008-s)
BPG_DI sp_xyzz sp_madd(const sp_xyzz& p, const sp_aff& q0, bool neg) {
  sp_aff q = q0;
  q.y = fp_sel(neg, fp_neg(q0.y), q0.y);
  if (sp_aff_is_identity(q0)) return p;
  if (sp_is_identity(p)) return sp_from_aff(q);
  fp U2 = fp_mul(q.x, p.ZZ);
  fp S2 = fp_mul(q.y, p.ZZZ);
  fp Pd = fp_sub(U2, p.X);
  fp Rd = fp_sub(S2, p.Y);
  if (fp_is_zero(Pd)) return fp_is_zero(Rd) ? sp_mdbl(q) : sp_identity();
  fp PP = fp_sq(Pd);
  fp PPP = fp_mul(Pd, PP);
  fp Q = fp_mul(p.X, PP);
  sp_xyzz r;
  r.X = fp_sub(fp_sub(fp_sq(Rd), PPP), fp_dbl(Q));
  r.Y = fp_sub(fp_mul(Rd, fp_sub(Q, r.X)), fp_mul(p.Y, PPP));
  r.ZZ = fp_mul(p.ZZ, PP);
  r.ZZZ = fp_mul(p.ZZZ, PPP);
  return r;
}

// p + q (add-2008-s)
BPG_DI sp_xyzz sp_add(const sp_xyzz& p, const sp_xyzz& q) {
  if (sp_is_identity(q)) return p;
  if (sp_is_identity(p)) return q;
  fp U1 = fp_mul(p.X, q.ZZ);
  fp U2 = fp_mul(q.X, p.ZZ);
  fp S1 = fp_mul(p.Y, q.ZZZ);
  fp S2 = fp_mul(q.Y, p.ZZZ);
  fp Pd = fp_sub(U2, U1);
  fp Rd = fp_sub(S2, S1);
  if (fp_is_zero(Pd)) return fp_is_zero(Rd) ? sp_dbl(p) : sp_identity();
  fp PP = fp_sq(Pd);
  fp PPP = fp_mul(Pd, PP);
  fp Q = fp_mul(U1, PP);
  sp_xyzz r;
  r.X = fp_sub(fp_sub(fp_sq(Rd), PPP), fp_dbl(Q));
  r.Y = fp_sub(fp_mul(Rd, fp_sub(Q, r.X)), fp_mul(S1, PPP));
  r.ZZ = fp_mul(fp_mul(p.ZZ, q.ZZ), PP);
  r.ZZZ = fp_mul(fp_mul(p.ZZZ, q.ZZZ), PPP);
  return r;
}

// XYZZ -> canonical affine words (x | y, 8 + 8 little-endian u32); the identity is all zero
BPG_DI void sp_to_affine_words(uint32_t out[16], const sp_xyzz& p) {
  if (sp_is_identity(p)) {
#pragma unroll
    for (int i = 0; i < 16; i++) out[i] = 0;
    return;
  }
  fp t = fp_invert(fp_mul(p.ZZ, p.ZZZ));
  fp x = fp_from_mont(fp_mul(p.X, fp_mul(t, p.ZZZ)));  // X / ZZ
  fp y = fp_from_mont(fp_mul(p.Y, fp_mul(t, p.ZZ)));   // Y / ZZZ
#pragma unroll
  for (int i = 0; i < 8; i++) {
    out[i] = x.v[i];
    out[8 + i] = y.v[i];
  }
}

// canonical affine words -> table entry; false if a coordinate is >= p or the point is off the curve
BPG_DI bool sp_from_affine_words(sp_aff& q, const uint32_t in[16]) {
  uint32_t any = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) any |= in[i];
  if (any == 0) {
    q.x = fp_zero();
    q.y = fp_zero();
    return true;
  }
  bool ok = fp_from_canonical(q.x, in) & fp_from_canonical(q.y, in + 8);
  // y^2 = x^3 + x + beta
  fp lhs = fp_sq(q.y);
  fp rhs = fp_add(fp_mul(fp_add(fp_sq(q.x), fp_one()), q.x), fp_const(BPG_K(KS_BETA)));
  return ok & fp_eq(lhs, rhs);
}

BPG_DI void sp_store(uint32_t* p, const sp_xyzz& a) {
  fp_store(p, a.X);
  fp_store(p + 8, a.Y);
  fp_store(p + 16, a.ZZ);
  fp_store(p + 24, a.ZZZ);
}
BPG_DI void sp_load(sp_xyzz& a, const uint32_t* p) {
  fp_load(a.X, p);
  fp_load(a.Y, p + 8);
  fp_load(a.ZZ, p + 16);
  fp_load(a.ZZZ, p + 24);
}
BPG_DI void sp_aff_store(uint32_t* p, const sp_aff& a) {
  fp_store(p, a.x);
  fp_store(p + 8, a.y);
}
BPG_DI void sp_aff_load(sp_aff& a, const uint32_t* p) {
  fp_load(a.x, p);
  fp_load(a.y, p + 8);
}

}  // namespace bpg
