// Stark-curve policy, point-level kernels: table decode, window multiples, final sum to affine bytes
// (the pipeline kernels are stark_msm.cuh; translation units include only what they launch).
#pragma once
#include "stark_pt.cuh"
#include "stark_pt4.cuh"
#include "stark_sc.cuh"

namespace bpg {

static __global__ void __launch_bounds__(128) k_stark_decode(const uint8_t* __restrict__ xy, uint32_t n,
                                                       uint32_t* __restrict__ table /*[n][16]*/,
                                                       uint32_t* __restrict__ bad_count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* src = reinterpret_cast<const uint4*>(xy + (size_t)i * 64);
  uint32_t w[16];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint4 a = src[k];
    w[4 * k] = a.x; w[4 * k + 1] = a.y; w[4 * k + 2] = a.z; w[4 * k + 3] = a.w;
  }
  sp_aff q;
  if (!sp_from_affine_words(q, w)) {
    q.x = fp_zero();
    q.y = fp_zero();
    atomicAdd(bad_count, 1u);
  }
  sp_aff_store(table + (size_t)i * 16, q);
}

// parts laid out [part][set][32 words]: sum over parts, then affine bytes (x || y, 32 LE each; identity = 0)
static __global__ void k_stark_finish(const uint32_t* __restrict__ parts, int nparts, int nsets, uint8_t* __restrict__ out_xy) {
  int set = blockIdx.x * blockDim.x + threadIdx.x;
  if (set >= nsets) return;
  sp_xyzz acc;
  sp_load(acc, parts + (size_t)set * 32);
  for (int p = 1; p < nparts; p++) {
    sp_xyzz o;
    sp_load(o, parts + ((size_t)p * nsets + set) * 32);
    acc = sp_add(acc, o);
  }
  uint32_t w[16];
  sp_to_affine_words(w, acc);
  uint32_t* dst = reinterpret_cast<uint32_t*>(out_xy + (size_t)set * 64);
#pragma unroll
  for (int i = 0; i < 16; i++) dst[i] = w[i];
}

// windowed tables: out[w][i] = 2^(c w) * P_i, affine, w < W.  One thread per point walks the doubling
// chain in XYZZ, parks the multiples and the running product of their ZZ*ZZZ in scratch, inverts
// once (Montgomery's trick) and converts every multiple back to affine (1/ZZ = t ZZZ, 1/ZZZ = t ZZ
// with t = 1/(ZZ ZZZ)).  One-time cost at table upload; it removes every doubling from later MSMs.
static __global__ void __launch_bounds__(128) k_stark_window_chain(const uint32_t* __restrict__ aff_in, uint32_t n_total,
                                                             uint32_t first, uint32_t count, int c, int W,
                                                             uint32_t* __restrict__ pt_scratch /*[W-1][count][32]*/,
                                                             uint32_t* __restrict__ dp_scratch /*[W-1][count][8]*/,
                                                             uint32_t* __restrict__ out /*[W][n_total][16]*/) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  uint32_t i = first + t;
  sp_aff q;
  sp_aff_load(q, aff_in + (size_t)i * 16);
  sp_aff_store(out + (size_t)i * 16, q);  // window 0
  if (sp_aff_is_identity(q)) {
    for (int w = 1; w < W; w++) sp_aff_store(out + ((size_t)w * n_total + i) * 16, q);
    return;
  }
  sp_xyzz p = sp_from_aff(q);
  fp dp = fp_one();
  for (int w = 1; w < W; w++) {
    for (int k = 0; k < c; k++) p = sp_dbl(p);  // prime order: never the identity
    dp = fp_mul(dp, fp_mul(p.ZZ, p.ZZZ));
    sp_store(pt_scratch + ((size_t)(w - 1) * count + t) * 32, p);
    fp_store(dp_scratch + ((size_t)(w - 1) * count + t) * 8, dp);
  }
  fp inv = fp_invert(dp);
  for (int w = W - 1; w >= 1; w--) {
    sp_xyzz e;
    sp_load(e, pt_scratch + ((size_t)(w - 1) * count + t) * 32);
    fp ti;
    if (w >= 2) {
      fp prev;
      fp_load(prev, dp_scratch + ((size_t)(w - 2) * count + t) * 8);
      ti = fp_mul(inv, prev);
    } else {
      ti = inv;
    }
    inv = fp_mul(inv, fp_mul(e.ZZ, e.ZZZ));
    sp_aff a;
    a.x = fp_mul(e.X, fp_mul(ti, e.ZZZ));
    a.y = fp_mul(e.Y, fp_mul(ti, e.ZZ));
    sp_aff_store(out + ((size_t)w * n_total + i) * 16, a);
  }
}


// ---------------------------------------------------------------------------
// The fork's generator chain (reference src/generators.rs:80-125): point = hash_to_scalar(state) * G with
// hash_to_scalar(low) = (low || keccak256(low)) read as a 512-bit little-endian integer mod the group order
// (src/util.rs:252-267).  The hash chain is sequential and runs on the host; the wide reduction and the
// fixed-base multiplication run here.  The comb of G holds (d+1) 16^j G, j < 64, d < 8, affine.
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(COMB_WINDOWS) k_stark_comb_build(const uint32_t* __restrict__ g_xy /*16 canonical words*/,
                                                                          uint32_t* __restrict__ comb /*[64][8][16]*/,
                                                                          uint32_t* __restrict__ bad) {
  const int j = threadIdx.x;
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = g_xy[i];
  sp_aff g;
  if (!sp_from_affine_words(g, w)) {
    if (j == 0) atomicAdd(bad, 1u);
    return;
  }
  sp_xyzz p = sp_from_aff(g);
  for (int k = 0; k < 4 * j; k++) p = sp_dbl(p);
  sp_xyzz m = p;
  for (int d = 0; d < 8; d++) {
    if (d) m = sp_add(m, p);
    sp_to_affine_words(w, m);
    sp_aff q;
    sp_from_affine_words(q, w);
    sp_aff_store(comb + (size_t)(j * 8 + d) * 16, q);
  }
}
// in: n x 64 bytes (low || high, little-endian); out: n x 64 bytes affine x || y of ((low + 2^256 high) mod order) * G
static __global__ void __launch_bounds__(128) k_stark_chain_points(const uint32_t* __restrict__ wide, uint32_t n,
                                                                   const uint32_t* __restrict__ comb, sc_bias bias4,
                                                                   uint8_t* __restrict__ out_xy) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sc lo, hi;
  sc_load(lo, wide + (size_t)i * 16);
  sc_load(hi, wide + (size_t)i * 16 + 8);
  const sc rr = sc_const(BPG_K(KSS_RR));
  sc one = sc_zero();
  one.v[0] = 1;
  // Montgomery form of lo + hi 2^256, then back: the canonical scalar
  sc km = scs_add(scs_montmul(lo, rr), scs_montmul(scs_montmul(hi, rr), rr));
  sc k = scs_montmul(km, one);
  const sc_recoded r = sc_recode(k.v, bias4);
  sp_xyzz acc = sp_identity();
  for (int j = 0; j < COMB_WINDOWS; j++) {
    int d = sc_digit(r, j, 4);
    if (d != 0) {
      int mag = d < 0 ? -d : d;
      sp_aff q;
      sp_aff_load(q, comb + (size_t)(j * 8 + mag - 1) * 16);
      acc = sp_madd(acc, q, d < 0);
    }
  }
  uint32_t w[16];
  sp_to_affine_words(w, acc);
  uint32_t* dst = reinterpret_cast<uint32_t*>(out_xy + (size_t)i * 64);
#pragma unroll
  for (int t = 0; t < 16; t++) dst[t] = w[t];
}

}  // namespace bpg
