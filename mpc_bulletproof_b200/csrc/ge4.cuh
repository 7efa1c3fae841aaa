// Quad-cooperative point arithmetic for the latency-bound tails of the MSM.
//
// A single warp keeps the INT32 multiply pipe of its SM sub-partition busy for
// ~260 ns per field multiplication no matter how many of its lanes are active
// (profiles/int32_peak.json, tools/fe_lat.cu), so the serial phases
// of Pippenger — running sums over buckets, the window combine, the doubling chain —
// cost (#dependent multiplications) x 260 ns when one thread owns a point.  Here FOUR
// adjacent lanes own one extended point, one coordinate each (lane&3 = 0:X 1:Y 2:Z
// 3:T), and the four independent multiplications of every formula level are one warp
// instruction stream: an addition is 2-3 multiplication levels instead of 8-9.
#pragma once
#include "ge.cuh"

namespace bpg {

struct ge4 {
  fe c;  // this lane's coordinate of its quad's point
};

#define BPG_FULL_MASK 0xffffffffu

// value of `x` held by lane (quad_base + src) for every lane of the quad
__device__ __forceinline__ fe fe_quad_get(const fe& x, int src) {
  fe o;
  int lane = (threadIdx.x & 31);
  int from = (lane & ~3) | src;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = __shfl_sync(BPG_FULL_MASK, x.v[i], from);
  return o;
}
__device__ __forceinline__ fe fe_sel(bool p, const fe& a, const fe& b) {
  fe o;
#pragma unroll
  for (int i = 0; i < 8; i++) o.v[i] = p ? a.v[i] : b.v[i];
  return o;
}

__device__ __forceinline__ ge4 ge4_identity() {
  int q = threadIdx.x & 3;
  ge4 r;
  r.c = (q == 1 || q == 2) ? fe_one() : fe_zero();
  return r;
}

// global layout of an extended point: 32 words X|Y|Z|T; lane q moves words [8q, 8q+8)
__device__ __forceinline__ ge4 ge4_load(const uint32_t* p) {
  ge4 r;
  fe_load(r.c, p + 8 * (threadIdx.x & 3));
  return r;
}
__device__ __forceinline__ void ge4_store(uint32_t* p, const ge4& a) { fe_store(p + 8 * (threadIdx.x & 3), a.c); }

// "cached" operand of an addition: lane0 Y-X, lane1 Y+X, lane2 2Z, lane3 2d*T.  One multiplication level.
__device__ __forceinline__ ge4 ge4_to_cached(const ge4& p) {
  int q = threadIdx.x & 3;
  fe other = fe_quad_get(p.c, q ^ 1);  // lanes 0,1 swap X/Y; lanes 2,3 swap Z/T (unused)
  fe ymx = fe_sub(other, p.c);         // lane0: Y - X
  fe ypx = fe_add_nc(other, p.c);      // lane1: X + Y; lane2: T + Z... only lanes 0,1 use these
  fe z2 = fe_add_nc(p.c, p.c);         // lane2: 2Z
  fe pre = q == 0 ? ymx : (q == 1 ? ypx : (q == 2 ? z2 : p.c));
  fe k = q == 3 ? fe_const(BPG_K(K_D2)) : fe_one();
  ge4 r;
  r.c = fe_mul(pre, k);  // tightens lanes 0..2, scales T on lane 3
  return r;
}

// p + q with q cached.  Two multiplication levels.  Note the cached layout: lane2 2Z, lane3 2dT.
__device__ __forceinline__ ge4 ge4_add_cached(const ge4& p, const ge4& qc) {
  int q = threadIdx.x & 3;
  fe other = fe_quad_get(p.c, q ^ 1);
  fe ymx = fe_sub(other, p.c);     // lane0
  fe ypx = fe_add_nc(other, p.c);  // lane1
  // first-level operand of P: lane0 Y1-X1, lane1 Y1+X1, lane2 Z1 (pairs with 2Z2), lane3 T1 (pairs with 2dT2)
  fe op = q == 0 ? ymx : (q == 1 ? ypx : p.c);
  fe m = fe_mul(op, qc.c);  // lane0 A, lane1 B, lane2 D = 2 Z1 Z2, lane3 C = 2d T1 T2
  fe A = fe_quad_get(m, 0), B = fe_quad_get(m, 1), D = fe_quad_get(m, 2), C = fe_quad_get(m, 3);
  fe E = fe_sub(B, A), H = fe_add_nc(B, A), F = fe_sub(D, C), Gg = fe_add_nc(D, C);
  // lane0 X3 = E F, lane1 Y3 = G H, lane2 Z3 = F G, lane3 T3 = E H
  fe l = (q == 0 || q == 3) ? E : (q == 1 ? Gg : F);
  fe r = (q == 0) ? F : ((q == 2) ? Gg : H);
  ge4 o;
  o.c = fe_mul(l, r);
  return o;
}

// p + q, both extended: three multiplication levels
__device__ __forceinline__ ge4 ge4_add(const ge4& p, const ge4& q) { return ge4_add_cached(p, ge4_to_cached(q)); }

// 2p: two levels (4 squarings, 4 multiplications)
__device__ __forceinline__ ge4 ge4_dbl(const ge4& p) {
  int q = threadIdx.x & 3;
  fe X = fe_quad_get(p.c, 0), Y = fe_quad_get(p.c, 1);
  fe xy = fe_add_nc(X, Y);
  fe op = q == 3 ? xy : p.c;  // lane0 X, lane1 Y, lane2 Z, lane3 X+Y
  fe s = fe_sq(op);
  fe A = fe_quad_get(s, 0), B = fe_quad_get(s, 1), ZZ = fe_quad_get(s, 2), S = fe_quad_get(s, 3);
  fe Cc = fe_add_nc(ZZ, ZZ);
  fe AB = fe_add_nc(A, B);
  fe E = fe_sub(S, AB);
  fe Gg = fe_sub(B, A);
  fe F = fe_sub(Gg, Cc);
  fe H = fe_neg(AB);
  fe l = (q == 0 || q == 3) ? E : (q == 1 ? Gg : F);
  fe r = (q == 0) ? F : ((q == 2) ? Gg : H);
  ge4 o;
  o.c = fe_mul(l, r);
  return o;
}

// gather the quad's point into every lane's registers (for single-lane epilogues such as encode)
__device__ __forceinline__ ge_ext ge4_gather(const ge4& p) {
  ge_ext r;
  r.X = fe_quad_get(p.c, 0);
  r.Y = fe_quad_get(p.c, 1);
  r.Z = fe_quad_get(p.c, 2);
  r.T = fe_quad_get(p.c, 3);
  return r;
}

// sum of one point per quad over the whole block -> quad 0 of warp 0.
// sm: [warps][32] words.  Every thread of the block must call it.
__device__ __forceinline__ ge4 block_sum_quads(ge4 p, uint32_t (*sm)[32]) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int off = 16; off >= 4; off >>= 1) {
    ge4 o;
#pragma unroll
    for (int i = 0; i < 8; i++) o.c.v[i] = __shfl_down_sync(BPG_FULL_MASK, p.c.v[i], off);
    p = ge4_add(p, o);
  }
  if (nw == 1) return p;
  if (lane < 4) ge4_store(sm[wid], p);
  __syncthreads();
  if (wid == 0) {
    int quad = lane >> 2;
    ge4 t = ge4_identity();
    // up to 32 warps: each quad folds warps quad, quad+8, ...
    for (int k = 0; k < (nw + 7) / 8; k++) {
      int w = quad + 8 * k;
      ge4 o = w < nw ? ge4_load(sm[w]) : ge4_identity();
      t = ge4_add(t, o);
    }
#pragma unroll
    for (int off = 16; off >= 4; off >>= 1) {
      ge4 o;
#pragma unroll
      for (int i = 0; i < 8; i++) o.c.v[i] = __shfl_down_sync(BPG_FULL_MASK, t.c.v[i], off);
      t = ge4_add(t, o);
    }
    p = t;
  }
  __syncthreads();
  return p;
}

}  // namespace bpg
