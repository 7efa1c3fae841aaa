// Candidate field multiplications for tools/fe_mul_bench.cu (measurement only; the shipped one is fe.cuh).
#pragma once
#include "../mpc_bulletproof_b200/csrc/fe.cuh"
namespace bpg {

// acc pair (lo, hi) += a*b, carry out of the pair counted in cnt: ONE wide multiply-add with carry OUT only
// (never the half-rate carry-IN form) + one add-with-carry on the other pipe.
BPG_DI void mac1(uint32_t& lo, uint32_t& hi, uint32_t& cnt, uint32_t a, uint32_t b) {
  mad_wide_cc(lo, hi, a, b);
  cnt = addc(cnt, 0u);
}

// V1: every product is a first link.  e pairs (e[2k], e[2k+1]) = words (2k, 2k+1), o pairs (o[2k], o[2k+1]) =
// words (2k+1, 2k+2); ce[k] counts carries into word 2k+2, co[k] into word 2k+3.
BPG_DI fe fe_mul_v1(const fe& A, const fe& B) {
  const uint32_t* a = A.v;
  const uint32_t* b = B.v;
  uint32_t e[16], o[16], ce[8], co[8];
#pragma unroll
  for (int k = 0; k < 8; k++) ce[k] = co[k] = 0;
  // first product of every pair: plain multiply.  pair e_k <- column 2k, pair o_k <- column 2k+1
#pragma unroll
  for (int k = 0; k < 8; k++) {
    // column 2k: (i, j) = (min(2k,7), 2k - min(2k,7))
    int i0 = 2 * k < 8 ? 2 * k : 7;
    if (2 * k <= 14) mul_wide(e[2 * k], e[2 * k + 1], a[i0], b[2 * k - i0]);
    int c1 = 2 * k + 1;
    int i1 = c1 < 8 ? c1 : 7;
    if (c1 <= 13) mul_wide(o[2 * k], o[2 * k + 1], a[i1], b[c1 - i1]);
  }
  o[14] = 0; o[15] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int c = i + j;
      int ifirst = c < 8 ? c : 7;  // the product used to initialise this column's pair
      if (i == ifirst) continue;
      if ((c & 1) == 0) mac1(e[c], e[c + 1], ce[c >> 1], a[i], b[j]);
      else mac1(o[c - 1], o[c], co[c >> 1], a[i], b[j]);
    }
  // words: w = e[w] + o[w-1] + counter(w); ce[k] -> word 2k+2, co[k] -> word 2k+3
  uint32_t r[16];
  r[0] = e[0];
  r[1] = add_cc(e[1], o[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) r[k] = addc_cc(e[k], o[k - 1]);
  r[15] = addc(e[15], o[14]);
  r[2] = add_cc(r[2], ce[0]);
#pragma unroll
  for (int w = 3; w < 15; w++) r[w] = addc_cc(r[w], (w & 1) ? co[(w - 3) >> 1] : ce[(w - 2) >> 1]);
  r[15] = addc(r[15], co[6]);
  return fe_reduce512(r);
}

// V2: as V1 but the reduction's eight folds are first links as well
BPG_DI fe fe_reduce512_v2(uint32_t r[16]) {
  uint32_t c[5] = {0, 0, 0, 0, 0};
  // (r0,r1)+=38 r8 ; (r2,r3)+=38 r10 ; (r4,r5)+=38 r12 ; (r6,r7)+=38 r14 : independent pairs, carries counted
  uint32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;
  mac1(r[0], r[1], x0, r[8], 38u);   // carry -> word 2
  mac1(r[2], r[3], x1, r[10], 38u);  // -> word 4
  mac1(r[4], r[5], x2, r[12], 38u);  // -> word 6
  mac1(r[6], r[7], x3, r[14], 38u);  // -> word 8 (2^256)
  // odd: (r1,r2)+=38 r9 ; (r3,r4)+=38 r11 ; (r5,r6)+=38 r13 ; (r7,t8)+=38 r15
  uint32_t y0 = 0, y1 = 0, y2 = 0, t8 = 0, y3 = 0;
  mac1(r[1], r[2], y0, r[9], 38u);   // -> word 3
  mac1(r[3], r[4], y1, r[11], 38u);  // -> word 5
  mac1(r[5], r[6], y2, r[13], 38u);  // -> word 7
  mac1(r[7], t8, y3, r[15], 38u);    // t8 = word 8; y3 always 0 (t8 starts at 0, product hi <= 37)
  (void)c;
  (void)y3;
  // add the counters: words 2..7, word 8 -> t8
  r[2] = add_cc(r[2], x0);
  r[3] = addc_cc(r[3], y0);
  r[4] = addc_cc(r[4], x1);
  r[5] = addc_cc(r[5], y1);
  r[6] = addc_cc(r[6], x2);
  r[7] = addc_cc(r[7], y2);
  t8 = addc(t8, x3);
  uint32_t top = (t8 << 1) | (r[7] >> 31);
  r[7] &= 0x7fffffffu;
  uint32_t f = top * 19u;
  fe out;
  out.v[0] = add_cc(r[0], f);
#pragma unroll
  for (int i = 1; i < 7; i++) out.v[i] = addc_cc(r[i], 0u);
  out.v[7] = addc(r[7], 0u);
  uint32_t bb = out.v[7] >> 31;
  out.v[7] &= 0x7fffffffu;
  out.v[0] += 19u * bb;
  return out;
}
BPG_DI fe fe_mul_v2(const fe& A, const fe& B) {
  const uint32_t* a = A.v;
  const uint32_t* b = B.v;
  uint32_t e[16], o[16], ce[8], co[8];
#pragma unroll
  for (int k = 0; k < 8; k++) ce[k] = co[k] = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    int i0 = 2 * k < 8 ? 2 * k : 7;
    if (2 * k <= 14) mul_wide(e[2 * k], e[2 * k + 1], a[i0], b[2 * k - i0]);
    int c1 = 2 * k + 1;
    int i1 = c1 < 8 ? c1 : 7;
    if (c1 <= 13) mul_wide(o[2 * k], o[2 * k + 1], a[i1], b[c1 - i1]);
  }
  o[14] = 0; o[15] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int c = i + j;
      int ifirst = c < 8 ? c : 7;
      if (i == ifirst) continue;
      if ((c & 1) == 0) mac1(e[c], e[c + 1], ce[c >> 1], a[i], b[j]);
      else mac1(o[c - 1], o[c], co[c >> 1], a[i], b[j]);
    }
  uint32_t r[16];
  r[0] = e[0];
  r[1] = add_cc(e[1], o[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) r[k] = addc_cc(e[k], o[k - 1]);
  r[15] = addc(e[15], o[14]);
  r[2] = add_cc(r[2], ce[0]);
#pragma unroll
  for (int w = 3; w < 15; w++) r[w] = addc_cc(r[w], (w & 1) ? co[(w - 3) >> 1] : ce[(w - 2) >> 1]);
  r[15] = addc(r[15], co[6]);
  return fe_reduce512_v2(r);
}
}  // namespace bpg
