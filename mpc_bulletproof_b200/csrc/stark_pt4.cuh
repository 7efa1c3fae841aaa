// Stark-curve policy: quad-cooperative XYZZ arithmetic for the latency-bound tails of the MSM
// (the counterpart of ge4.cuh).  Four adjacent lanes own one point, one coordinate each
// (lane&3 = 0:X 1:Y 2:ZZ 3:ZZZ); the independent multiplications of each formula level are one
// warp instruction stream: an addition is 4 multiplication levels instead of 14 dependent
// multiplications, a doubling 3 instead of 10.
//
// The short-Weierstrass formulas are not unified.  Identities are selected by flag; P = Q (same x,
// same y) needs the doubling, which is computed only when some quad of the warp needs it (a
// warp-uniform branch: the quad arithmetic shuffles warp-wide, so every lane of a warp must run
// the same instruction stream).
#pragma once
#include "stark_pt.cuh"

namespace bpg {

struct sp4 {
  fp c;  // this lane's coordinate of its quad's point
};

#ifndef BPG_FULL_MASK
#define BPG_FULL_MASK 0xffffffffu
#endif

// value of `x` held by lane (quad_base + src) — src may differ per lane
__device__ __forceinline__ fp fp_quad_get(const fp& x, int src) {
  fp o;
  int from = ((threadIdx.x & 31) & ~3) | src;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = __shfl_sync(BPG_FULL_MASK, x.v[i], from);
  return o;
}

__device__ __forceinline__ sp4 sp4_identity() {
  sp4 r;
  r.c = fp_zero();
  return r;
}
__device__ __forceinline__ sp4 sp4_load(const uint32_t* p) {
  sp4 r;
  fp_load(r.c, p + 8 * (threadIdx.x & 3));
  return r;
}
__device__ __forceinline__ void sp4_store(uint32_t* p, const sp4& a) { fp_store(p + 8 * (threadIdx.x & 3), a.c); }

// 2p: three levels (identity in, identity out: every product is zero)
__device__ __forceinline__ sp4 sp4_dbl(const sp4& p) {
  int q = threadIdx.x & 3;
  fp X = fp_quad_get(p.c, 0), Y = fp_quad_get(p.c, 1), ZZ = fp_quad_get(p.c, 2);
  fp U = fp_dbl(Y);
  // level 1: X^2 | U^2 = V | ZZ^2 | (idle)
  fp op1 = q == 0 ? X : (q == 1 ? U : ZZ);
  fp s1 = fp_sq(op1);
  fp XX = fp_quad_get(s1, 0), V = fp_quad_get(s1, 1), ZZsq = fp_quad_get(s1, 2);
  fp M = fp_add(fp_add(fp_dbl(XX), XX), ZZsq);  // 3 X^2 + a ZZ^2, a = 1
  // level 2: X V = S | U V = W | V ZZ = ZZ3 | M^2
  fp l2 = q == 0 ? X : (q == 1 ? U : (q == 2 ? V : M));
  fp r2 = q == 3 ? M : (q == 2 ? ZZ : V);
  fp m2 = fp_mul(l2, r2);
  fp S = fp_quad_get(m2, 0), W = fp_quad_get(m2, 1), MM = fp_quad_get(m2, 3);
  fp X3 = fp_sub(MM, fp_dbl(S));
  // level 3: (idle) | M (S - X3) | W Y | W ZZZ = ZZZ3
  fp l3 = q == 1 ? M : W;
  fp r3 = q == 1 ? fp_sub(S, X3) : (q == 2 ? Y : p.c /* lane 3: ZZZ */);
  fp m3 = fp_mul(l3, r3);
  fp Y3 = fp_sub(fp_quad_get(m3, 1), fp_quad_get(m3, 2));
  sp4 o;
  o.c = q == 0 ? X3 : (q == 1 ? Y3 : (q == 2 ? m2 /* V ZZ */ : m3 /* W ZZZ */));
  return o;
}

// p + q: four levels, exceptional cases by selection
__device__ __forceinline__ sp4 sp4_add(const sp4& p, const sp4& qq) {
  int q = threadIdx.x & 3;
  fp ZZ1 = fp_quad_get(p.c, 2), ZZ2 = fp_quad_get(qq.c, 2);
  bool idP = fp_is_zero(ZZ1), idQ = fp_is_zero(ZZ2);
  // level 1: X1 ZZ2 = U1 | X2 ZZ1 = U2 | Y1 ZZZ2 = S1 | Y2 ZZZ1 = S2
  int srcP = q == 1 ? 2 : (q == 2 ? 1 : q);             // X1, ZZ1, Y1, ZZZ1
  int srcQ = q == 0 ? 2 : (q == 1 ? 0 : (q == 2 ? 3 : 1));  // ZZ2, X2, ZZZ2, Y2
  fp a1 = fp_quad_get(p.c, srcP);
  fp b1 = fp_quad_get(qq.c, srcQ);
  fp m1 = fp_mul(a1, b1);
  fp U1 = fp_quad_get(m1, 0), U2 = fp_quad_get(m1, 1), S1 = fp_quad_get(m1, 2), S2 = fp_quad_get(m1, 3);
  fp Pd = fp_sub(U2, U1), Rd = fp_sub(S2, S1);
  bool same_x = fp_is_zero(Pd) && !idP && !idQ;
  bool need_dbl = same_x && fp_is_zero(Rd);
  // level 2: P^2 = PP | R^2 | ZZ1 ZZ2 | ZZZ1 ZZZ2
  fp l2 = q == 0 ? Pd : (q == 1 ? Rd : (q == 2 ? ZZ1 : p.c));
  fp r2 = q == 0 ? Pd : (q == 1 ? Rd : (q == 2 ? ZZ2 : qq.c));
  fp m2 = fp_mul(l2, r2);
  fp PP = fp_quad_get(m2, 0), RR = fp_quad_get(m2, 1);
  // level 3: U1 PP = Q | P PP = PPP | (ZZ1 ZZ2) PP = ZZ3 | (idle: repeats lane 1)
  fp l3 = q == 0 ? U1 : (q == 2 ? m2 : Pd);
  fp m3 = fp_mul(l3, PP);
  fp Qv = fp_quad_get(m3, 0), PPP = fp_quad_get(m3, 1);
  fp X3 = fp_sub(fp_sub(RR, PPP), fp_dbl(Qv));
  // level 4: (idle) | R (Q - X3) | S1 PPP | (ZZZ1 ZZZ2) PPP = ZZZ3
  fp l4 = q == 1 ? Rd : (q == 3 ? m2 : S1);
  fp r4 = q == 1 ? fp_sub(Qv, X3) : PPP;
  fp m4 = fp_mul(l4, r4);
  fp Y3 = fp_sub(fp_quad_get(m4, 1), fp_quad_get(m4, 2));
  sp4 o;
  o.c = q == 0 ? X3 : (q == 1 ? Y3 : (q == 2 ? m3 : m4));
  // exceptional cases
  if (__any_sync(BPG_FULL_MASK, need_dbl)) {
    sp4 d = sp4_dbl(p);
    o.c = fp_sel(need_dbl, d.c, o.c);
  }
  bool to_identity = same_x && !need_dbl;  // P = -Q
  o.c = fp_sel(to_identity, fp_zero(), o.c);
  o.c = fp_sel(idQ, p.c, o.c);
  o.c = fp_sel(idP && !idQ, qq.c, o.c);
  return o;
}

}  // namespace bpg
