// Comb construction kernels (throughput-bound: products stay inline; see comb_kernels.cuh for the scheme):
// the one-time comb of a resident table, the folded generators of a long inner-product argument
// (k_comb_materialize) and their per-proof combs (k_comb_chain, k_comb_multiples).
#pragma once
#include "comb_kernels.cuh"

namespace bpg {

// ---------------------------------------------------------------------------
// one-time: the comb of every point of a resident table.  Block = one point, thread j = window j:
// 4 j doublings, the eight multiples, ONE inversion for the eight (Montgomery's trick), affine Niels out.
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(COMB_WINDOWS) k_table_comb_build(const uint32_t* __restrict__ niels /*window 0 of the table*/,
                                                                   uint32_t first, uint32_t* __restrict__ comb) {
  const uint32_t i = first + blockIdx.x;
  const int j = threadIdx.x;
  ge_niels q;
  ge_load_niels(q, niels + (size_t)i * NIELS_WORDS);
  ge_ext p = ge_from_niels(q, false);
  for (int k = 0; k < 4 * j; k++) p = ge_dbl(p);
  // forward: the eight multiples, parked projectively (X | Y | Z = 24 words) in the entries they will become,
  // with the running product of their Z; backward: one inversion serves all eight (Montgomery's trick)
  uint32_t* out = comb + ((size_t)i * COMB_WINDOWS + j) * 8 * COMB_AFFINE_WORDS;
  ge_ext m = p;
  fe zp[8];
#pragma unroll
  for (int d = 0; d < 8; d++) {
    if (d) m = ge_add(m, p);
    zp[d] = d ? fe_mul(zp[d - 1], m.Z) : m.Z;
    fe_store(out + (size_t)d * COMB_AFFINE_WORDS, m.X);
    fe_store(out + (size_t)d * COMB_AFFINE_WORDS + 8, m.Y);
    fe_store(out + (size_t)d * COMB_AFFINE_WORDS + 16, m.Z);
  }
  fe inv = fe_invert(zp[7]);
#pragma unroll
  for (int d = 7; d >= 0; d--) {
    fe X, Y, Z;
    fe_load(X, out + (size_t)d * COMB_AFFINE_WORDS);
    fe_load(Y, out + (size_t)d * COMB_AFFINE_WORDS + 8);
    fe_load(Z, out + (size_t)d * COMB_AFFINE_WORDS + 16);
    fe zi = d ? fe_mul(inv, zp[d - 1]) : inv;
    inv = fe_mul(inv, Z);
    ge_store_niels(out + (size_t)d * COMB_AFFINE_WORDS, ge_affine_to_niels(fe_mul(X, zi), fe_mul(Y, zi)));
  }
}

// ---------------------------------------------------------------------------
// folded generators, once: out[g] for g < 2 m0 is G'_p (g = p) or H'_p (g = m0 + p),
//   G'_p = sum_{t < n/m0} wG(t m0 + p) * G_{t m0 + p}            (weights in Montgomery form)
// and out[2 m0] = q_mul * (point q_id) when q_mul is given.  One WARP per output; its lanes share the
// (term, window-slice) units, then a tree over the warp.
// ---------------------------------------------------------------------------
struct CombMat {
  const uint32_t* comb;  // generator combs (affine)
  uint32_t g_id, h_id, q_id;
  const uint32_t *wG, *wH;
  const uint32_t* q_mul;  // canonical scalar or null
  uint32_t n, m0;
  uint32_t msplit;  // warps per output: each takes 64 / msplit windows of every term and leaves a partial sum
  sc_bias bias4;
};
static __global__ void __launch_bounds__(CB_THREADS, 4) k_comb_materialize(CombMat M, uint32_t* __restrict__ out /*[msplit][2 m0][32] ext*/) {
  __shared__ __align__(16) uint32_t pts[CB_THREADS][32];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ngroups = 2 * M.m0 + (M.q_mul ? 1u : 0u);
  uint32_t gw = blockIdx.x * (CB_THREADS / 32) + warp;
  const bool live = gw < ngroups * M.msplit;
  if (!live) gw = 0;  // idle warps shadow group 0 (the quad arithmetic shuffles warp-wide) and do not store
  uint32_t g = gw / M.msplit;
  const uint32_t part = gw % M.msplit;
  const bool is_q = g == 2 * M.m0;
  const bool is_h = !is_q && g >= M.m0;
  const uint32_t p = is_q ? 0 : (is_h ? g - M.m0 : g);
  const uint32_t gterms = is_q ? 1u : M.n / M.m0;
  const uint32_t ws = gterms >= 32 ? 1u : 32u / gterms;  // window slices per term: gterms * ws >= 32 units
  const int per = COMB_WINDOWS / (int)(ws * M.msplit);
  ge_ext acc = ge_identity();
  for (uint32_t unit = lane; unit < gterms * ws; unit += 32) {
    const uint32_t t = unit / ws, slice = part * ws + unit % ws;
    sc v;
    uint32_t id;
    if (is_q) {
      sc_load(v, M.q_mul);
      id = M.q_id;
    } else {
      const uint32_t i = t * M.m0 + p;
      sc_load(v, (is_h ? M.wH : M.wG) + (size_t)i * 8);
      v = sc_from_mont(v);
      id = (is_h ? M.h_id : M.g_id) + i;
    }
    const sc_recoded r = sc_recode(v.v, M.bias4);
    acc = comb_windows<true>(M.comb + (size_t)id * COMB_ENTRIES * COMB_AFFINE_WORDS, r, (int)slice * per, (int)(slice + 1) * per, acc);
  }
  ge_store_ext(pts[threadIdx.x], acc);
  __syncwarp();
  // quad q of the warp sums lanes 4q..4q+3, then a shuffle tree over the eight quads
  const uint32_t quad = lane >> 2;
  ge4 tq = ge4_load(pts[warp * 32 + 4 * quad]);
#pragma unroll
  for (int k = 1; k < 4; k++) tq = ge4_add(tq, ge4_load(pts[warp * 32 + 4 * quad + k]));
#pragma unroll
  for (int off = 16; off >= 4; off >>= 1) {
    ge4 o;
#pragma unroll
    for (int w = 0; w < 8; w++) o.c.v[w] = __shfl_down_sync(BPG_FULL_MASK, tq.c.v[w], off);
    tq = ge4_add(tq, o);
  }
  if (live && lane < 4) ge4_store(out + ((size_t)part * ngroups + g) * 32, tq);
}

// ---------------------------------------------------------------------------
// combs of the folded generators (per proof): the doubling chain 16^j P on a quad per point, then the eight
// multiples of every (point, window) in the projective "cached" layout -- no inversion anywhere.
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(CB_THREADS) k_comb_chain(const uint32_t* __restrict__ pts_in /*[nparts][npts][32] ext*/,
                                                           uint32_t npts, uint32_t nparts,
                                                           uint32_t* __restrict__ chain /*[npts][64][32] ext*/) {
  uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const bool live = q < npts;
  if (!live) q = npts - 1;
  ge4 cur = ge4_load(pts_in + (size_t)q * 32);
  for (uint32_t p = 1; p < nparts; p++) cur = ge4_add(cur, ge4_load(pts_in + ((size_t)p * npts + q) * 32));
  uint32_t* dst = chain + (size_t)q * COMB_WINDOWS * 32;
  if (live) ge4_store(dst, cur);
#pragma unroll 1
  for (int j = 1; j < COMB_WINDOWS; j++) {
    cur = ge4_dbl(cur);
    cur = ge4_dbl(cur);
    cur = ge4_dbl(cur);
    cur = ge4_dbl(cur);
    if (live) ge4_store(dst + (size_t)j * 32, cur);
  }
}
static __global__ void __launch_bounds__(CB_THREADS) k_comb_multiples(const uint32_t* __restrict__ chain, uint32_t nwin /*npts * 64*/,
                                                               uint32_t* __restrict__ comb /*[npts][64][8][32] cached*/) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nwin) return;
  ge_ext base;
  ge_load_ext(base, chain + (size_t)t * 32);
  const fe ymx = fe_sub(base.Y, base.X), ypx = fe_add_nc(base.Y, base.X), z2 = fe_add_nc(base.Z, base.Z);
  const fe t2d = fe_mul(base.T, fe_const(BPG_K(K_D2)));
  uint32_t* out = comb + (size_t)t * 8 * COMB_CACHED_WORDS;
  fe_store(out, ymx);
  fe_store(out + 8, ypx);
  fe_store(out + 16, z2);
  fe_store(out + 24, t2d);
  ge_ext m = base;
#pragma unroll 1
  for (int d = 1; d < 8; d++) {
    m = ge_add_cached(m, ymx, ypx, z2, t2d);
    uint32_t* o = out + (size_t)d * COMB_CACHED_WORDS;
    fe_store(o, fe_sub(m.Y, m.X));
    fe_store(o + 8, fe_add_nc(m.Y, m.X));
    fe_store(o + 16, fe_add_nc(m.Z, m.Z));
    fe_store(o + 24, fe_mul(m.T, fe_const(BPG_K(K_D2))));
  }
}

// compressed -> extended (Z = 1), one WARP per point on the sixteen-lane field layer (fe16.cuh, whole-warp form): the
// decoding is an inverse square root, 254 dependent squarings, and it heads the chain a small verification waits for
// (75 us with a thread per point).  Invalid encodings count in *bad and become the identity.
static __global__ void __launch_bounds__(32) k_decode_ext(const uint8_t* __restrict__ comp, uint32_t n, uint32_t* __restrict__ ext,
                                                   uint32_t* __restrict__ bad) {
  __shared__ __align__(16) uint32_t sm[G16_WORDS];
  const uint32_t i = blockIdx.x;  // grid = n
  if (i >= n) return;
  uint8_t buf[32];
  for (int k = 0; k < 32; k++) buf[k] = comp[(size_t)i * 32 + k];
  grp16 g;
  g.sm = sm;
  g.k = threadIdx.x & 15u;
  g.half = (threadIdx.x >> 4) & 1u;
  g.par = 0;
  ge_ext p;
  const bool ok = ge_decode16<true>(g, buf, p);
  if (threadIdx.x == 0) {
    if (!ok) {
      p = ge_identity();
      atomicAdd(bad, 1u);
    }
    ge_store_ext(ext + (size_t)i * 32, p);
  }
}

}  // namespace bpg
