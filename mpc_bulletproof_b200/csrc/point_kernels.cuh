// Point-level kernels around the Pippenger pipeline, sm_100a: table construction (decode, window
// multiples), the fixed-base comb, generator derivation, the final sum + encoding, the fused peer
// exchange.  (The pipeline itself is msm_kernels.cuh; translation units include only what they launch.)
#pragma once
#include "ge.cuh"
#include "ge4.cuh"
#include "fe16.cuh"
#include "sc.cuh"

namespace bpg {

// point ids of up to four consecutive ranges [off_i, off_i + len_i) of one table
struct SegIds {
  uint32_t off[4], len[4];
  int n;
};
static __global__ void __launch_bounds__(256) k_seg_point_ids(SegIds sg, uint32_t total, uint32_t* __restrict__ ids) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  uint32_t r = t, id = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (i < sg.n) {
      if (r < sg.len[i]) {
        id = sg.off[i] + r;
        r = 0xffffffffu;
      } else if (r != 0xffffffffu) {
        r -= sg.len[i];
      }
    }
  }
  ids[t] = id;
}

// ---------------------------------------------------------------------------
// finishing: sum `nparts` partial sums per set (one per rank), encode
// ---------------------------------------------------------------------------
// parts layout: [part][set][32 words]
// One WARP per set: the sum of the parts is computed redundantly by its lanes, the encoding (one
// inverse square root, 252 dependent squarings) runs on the sixteen-lane field layer of fe16.cuh
// in its whole-warp form (the half-warps split every product).
constexpr int ENC_THREADS = 32;
__device__ __forceinline__ void store_s_bytes(uint8_t* out, const fe& s, uint32_t k) {
  uint32_t w = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) w = (k >> 1) == (uint32_t)i ? s.v[i] : w;
  w = (k & 1u) ? (w >> 16) : w;
  out[2 * k] = (uint8_t)w;
  out[2 * k + 1] = (uint8_t)(w >> 8);
}
__device__ __forceinline__ grp16 warp_group(uint32_t* sm_of_warp) {
  grp16 g;
  g.sm = sm_of_warp;
  g.k = threadIdx.x & 15u;
  g.half = (threadIdx.x >> 4) & 1u;
  g.par = 0;
  return g;
}
static __global__ void __launch_bounds__(ENC_THREADS) k_sum_encode(const uint32_t* __restrict__ parts, int nparts, int nsets,
                                                             uint8_t* __restrict__ out_bytes /*nsets*32*/,
                                                             uint32_t* __restrict__ out_ext /*nsets*32 words, may be null*/) {
  __shared__ __align__(16) uint32_t sm[G16_WORDS];
  const uint32_t set = blockIdx.x;  // grid = nsets
  grp16 g = warp_group(sm);
  ge_ext acc;
  ge_load_ext(acc, parts + (size_t)set * 32);
  for (int p = 1; p < nparts; p++) {
    ge_ext o;
    ge_load_ext(o, parts + ((size_t)p * nsets + set) * 32);
    acc = ge_add(acc, o);
  }
  if (out_ext && threadIdx.x == 0) ge_store_ext(out_ext + (size_t)set * 32, acc);
  if (out_bytes) {
    fe s = ge_encode16<true>(g, acc);
    if (threadIdx.x < 16) store_s_bytes(out_bytes + (size_t)set * 32, s, g.k);
  }
}

// out[set] = encode(sum_p decode(points[p][set])): the "open" of additively shared points (r1cs_mpc: the
// parties' shares of a commitment, exchanged as compressed encodings, reference
// src/r1cs_mpc/mpc_prover.rs:630-657, mpc_inner_product.rs:131,191).  One warp per set: lane p decodes
// part p (an inverse square root each), lane 0 adds, the warp encodes.  bad_count: invalid encodings.
static __global__ void __launch_bounds__(ENC_THREADS) k_points_sum(const uint8_t* __restrict__ points, int nparts, int nsets,
                                                             uint8_t* __restrict__ out_bytes,
                                                             uint32_t* __restrict__ bad_count) {
  __shared__ __align__(16) uint32_t sm[G16_WORDS];
  __shared__ __align__(16) uint32_t dec[32][32];
  const uint32_t set = blockIdx.x;
  grp16 g = warp_group(sm);
  if ((int)threadIdx.x < nparts) {
    uint8_t buf[32];
    const uint8_t* src = points + ((size_t)threadIdx.x * nsets + set) * 32;
    for (int i = 0; i < 32; i++) buf[i] = src[i];
    ge_ext p;
    if (!ge_decode(p, buf)) {
      atomicAdd(bad_count, 1u);
      p = ge_identity();
    }
    ge_store_ext(dec[threadIdx.x], p);
  }
  __syncwarp();
  ge_ext acc;
  ge_load_ext(acc, dec[0]);
  for (int p = 1; p < nparts; p++) {
    ge_ext o;
    ge_load_ext(o, dec[p]);
    acc = ge_add(acc, o);
  }
  fe s = ge_encode16<true>(g, acc);
  if (threadIdx.x < 16) store_s_bytes(out_bytes + (size_t)set * 32, s, g.k);
}

// Accept-iff-identity (Verifier::verify, reference src/r1cs/verifier.rs:549): no encoding, hence no
// inverse square root.  A ristretto255 element equals the identity iff X = 0 or Y = 0 (RFC 9496
// §4.3.3: X1 Y2 == Y1 X2 or Y1 Y2 == X1 X2 against (0 : 1 : 1 : 0)).  out: 32 zero bytes (the
// identity's encoding) or 0x01 0x00.. (odd, so not a canonical encoding of anything).
static __global__ void __launch_bounds__(32) k_sum_is_identity(const uint32_t* __restrict__ parts, int nparts,
                                                        uint8_t* __restrict__ out_bytes) {
  if (threadIdx.x != 0) return;
  ge_ext acc;
  ge_load_ext(acc, parts);
  for (int p = 1; p < nparts; p++) {
    ge_ext o;
    ge_load_ext(o, parts + (size_t)p * 32);
    acc = ge_add(acc, o);
  }
  bool id = fe_is_zero(acc.X) | fe_is_zero(acc.Y);
  uint32_t* w = reinterpret_cast<uint32_t*>(out_bytes);
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = 0;
  if (!id) w[0] = 1;
}

// ---------------------------------------------------------------------------
// Sharded MSM, the exchange step fused with the combine: ONE kernel per rank
//   1. stores this rank's partial sums (n_sets x 128 B) into slot [rank] of EVERY rank's exchange
//      buffer over NVLink (peer-mapped pointers, plain stores), fences system-wide and raises its
//      flag in every rank's flag array;
//   2. waits until all ranks' flags show this step's sequence number;
//   3. adds the `world` partials per set and encodes.
// The payload is 128 B per rank and set, so the cost is latency: this replaces an NCCL all-gather
// plus a separate combine launch.  Buffers are double-buffered by step parity: a rank can be at
// most one step ahead of the slowest one (it needs that rank's flag to finish a step).
// ---------------------------------------------------------------------------
// extended point from words written by a peer: volatile loads (never served from a stale L1 line)
__device__ __forceinline__ ge_ext ge_load_ext_volatile(const uint32_t* p) {
  const volatile uint32_t* src = p;
  ge_ext q;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    q.X.v[i] = src[i];
    q.Y.v[i] = src[8 + i];
    q.Z.v[i] = src[16 + i];
    q.T.v[i] = src[24 + i];
  }
  return q;
}
struct PeerPtrs {
  uint32_t* parts[8];  // rank p's parts buffer:  [2][world][max_sets][32] words
  uint32_t* flags[8];  // rank p's flags:          [2][world]
};
constexpr int XCH_THREADS = 256;
static __global__ void __launch_bounds__(XCH_THREADS) k_exchange_sum_encode(const uint32_t* __restrict__ local_part, PeerPtrs peers,
                                                                      int world, int rank, int nsets, int max_sets,
                                                                      uint32_t seq, uint8_t* __restrict__ out_bytes,
                                                                      uint32_t* __restrict__ out_ext,
                                                                      uint32_t* __restrict__ status /*0 ok, 1 timeout*/) {
  const uint32_t slot = seq & 1u;
  const size_t slot_words = (size_t)world * max_sets * 32;
  // 1. push
  for (int p = 0; p < world; p++) {
    uint32_t* dst = peers.parts[p] + slot * slot_words + (size_t)rank * max_sets * 32;
    for (int w = threadIdx.x; w < nsets * 32; w += blockDim.x) dst[w] = local_part[w];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < (uint32_t)world) {
    volatile uint32_t* f = peers.flags[threadIdx.x] + slot * world + rank;
    *f = seq;
  }
  // 2. wait for every rank's flag (bounded: a dead peer must not hang the GPU)
  __shared__ uint32_t timed_out;
  if (threadIdx.x == 0) timed_out = 0;
  __syncthreads();
  if (threadIdx.x < (uint32_t)world) {
    volatile uint32_t* f = peers.flags[rank] + slot * world + threadIdx.x;
    uint32_t spins = 0;
    while (*f != seq) {
      __nanosleep(64);
      if (++spins > (1u << 24)) {  // > 1 s
        timed_out = 1;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  if (timed_out) {
    if (threadIdx.x == 0) *status = 1;
    return;
  }
  // 3. combine: one warp per set (the sum computed by each of its lanes, the encoding on the
  // whole-warp form of fe16.cuh); partials read past the L1 (they were written by peers)
  __shared__ __align__(16) uint32_t sm16[(XCH_THREADS / 32) * G16_WORDS];
  grp16 g = warp_group(sm16 + (threadIdx.x >> 5) * G16_WORDS);
  const uint32_t* base = peers.parts[rank] + slot * slot_words;
  for (int first = 0; first < nsets; first += XCH_THREADS / 32) {  // block-uniform trip count
    int set = first + (int)(threadIdx.x >> 5);
    if (set >= nsets) continue;  // whole warps drop out: the exchanges are warp-wide
    ge_ext acc = ge_load_ext_volatile(base + (size_t)set * 32);
    for (int p = 1; p < world; p++) acc = ge_add(acc, ge_load_ext_volatile(base + ((size_t)p * max_sets + set) * 32));
    if (out_ext && (threadIdx.x & 31) == 0) ge_store_ext(out_ext + (size_t)set * 32, acc);
    if (out_bytes) {
      fe s = ge_encode16<true>(g, acc);
      if ((threadIdx.x & 31) < 16) store_s_bytes(out_bytes + (size_t)set * 32, s, g.k);
    }
  }
}

// ---------------------------------------------------------------------------
// table construction: compressed ristretto -> affine Niels
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(128) k_decode_to_niels(const uint8_t* __restrict__ comp, uint32_t n,
                                                          uint32_t* __restrict__ table,
                                                          uint32_t* __restrict__ bad_count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t buf[32];
  const uint4* src = reinterpret_cast<const uint4*>(comp + (size_t)i * 32);
  uint4 a = src[0], b = src[1];
  uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int k = 0; k < 8; k++) {
    buf[4 * k] = (uint8_t)w[k];
    buf[4 * k + 1] = (uint8_t)(w[k] >> 8);
    buf[4 * k + 2] = (uint8_t)(w[k] >> 16);
    buf[4 * k + 3] = (uint8_t)(w[k] >> 24);
  }
  ge_ext p;
  bool ok = ge_decode(p, buf);
  ge_niels q;
  if (ok) {
    q = ge_affine_to_niels(p.X, p.Y);
  } else {
    q = ge_niels_identity();
    atomicAdd(bad_count, 1u);
  }
  ge_store_niels(table + (size_t)i * NIELS_WORDS, q);
}

// extended -> compressed, one thread per point
static __global__ void __launch_bounds__(128) k_encode(const uint32_t* __restrict__ ext, uint32_t n,
                                                 uint8_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge_ext p;
  ge_load_ext(p, ext + (size_t)i * 32);
  ge_encode(out + (size_t)i * 32, p);
}

// ---------------------------------------------------------------------------
// K-FIXED: fixed-base comb.  tab[j][d] = (d+1) * 16^j * P  (j < 64, d < 8), affine
// Niels, so k*P is 64 mixed additions and no doublings.  Serves the two-term
// Pedersen commitments `v*B + v_blinding*B_blinding` (reference
// src/generators.rs:41-43; prover.rs:325,627-631,687) and synthetic point sets.
// ---------------------------------------------------------------------------

static __global__ void __launch_bounds__(COMB_WINDOWS) k_comb_build(const uint8_t* __restrict__ base32,
                                                              uint32_t* __restrict__ table,
                                                              uint32_t* __restrict__ bad_count) {
  int j = threadIdx.x;
  uint8_t buf[32];
  for (int i = 0; i < 32; i++) buf[i] = base32[i];
  ge_ext p;
  if (!ge_decode(p, buf)) {
    if (j == 0) atomicAdd(bad_count, 1u);
    p = ge_identity();
  }
  for (int i = 0; i < 4 * j; i++) p = ge_dbl(p);
  ge_ext m = p;
  for (int d = 0; d < 8; d++) {
    ge_store_niels(table + (size_t)(j * 8 + d) * 24, ge_to_niels(m));
    m = ge_add(m, p);
  }
}

// out[i] = sum_t scalars[t*n + i] * base_t  for `nbases` comb tables laid out back to back
static __global__ void __launch_bounds__(128) k_comb_mul(const uint32_t* __restrict__ tables, int nbases,
                                                   const uint32_t* __restrict__ scalars, uint32_t n, sc_bias bias4,
                                                   uint8_t* __restrict__ out_bytes,
                                                   uint32_t* __restrict__ out_ext) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge_ext acc = ge_identity();
  for (int t = 0; t < nbases; t++) {
    sc k;
    sc_load(k, scalars + ((size_t)t * n + i) * 8);
    sc_recoded r = sc_recode(k.v, bias4);
    const uint32_t* tab = tables + (size_t)t * COMB_ENTRIES * 24;
    for (int j = 0; j < COMB_WINDOWS; j++) {
      int d = sc_digit(r, j, 4);
      if (d != 0) {
        int mag = d < 0 ? -d : d;
        ge_niels q;
        ge_load_niels(q, tab + (size_t)(j * 8 + mag - 1) * 24);
        acc = ge_madd(acc, q, d < 0);
      }
    }
  }
  if (out_ext) ge_store_ext(out_ext + (size_t)i * 32, acc);
  if (out_bytes) ge_encode(out_bytes + (size_t)i * 32, acc);
}

// Few outputs (the five T_i of a proof, a V_j): one WARP per output.  The nbases*64 table lookups
// are spread over the lanes (a handful of mixed additions each), then a shuffle tree of five full
// additions; 128 dependent additions become ~4 + 5.
static __global__ void __launch_bounds__(128) k_comb_mul_warp(const uint32_t* __restrict__ tables, int nbases,
                                                        const uint32_t* __restrict__ scalars, uint32_t n, sc_bias bias4,
                                                        uint8_t* __restrict__ out_bytes, uint32_t* __restrict__ out_ext) {
  uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t lane = threadIdx.x & 31;
  bool live = i < n;
  uint32_t ii = live ? i : n - 1;  // idle warps shadow the last output (whole warps, shuffles stay uniform)
  ge_ext acc = ge_identity();
  for (int t = 0; t < nbases; t++) {
    sc k;
    sc_load(k, scalars + ((size_t)t * n + ii) * 8);
    sc_recoded r = sc_recode(k.v, bias4);
    const uint32_t* tab = tables + (size_t)t * COMB_ENTRIES * 24;
    for (int j = lane; j < COMB_WINDOWS; j += 32) {
      int d = sc_digit(r, j, 4);
      if (d != 0) {
        int mag = d < 0 ? -d : d;
        ge_niels q;
        ge_load_niels(q, tab + (size_t)(j * 8 + mag - 1) * 24);
        acc = ge_madd(acc, q, d < 0);
      }
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    ge_ext o;
#pragma unroll
    for (int w = 0; w < 8; w++) {
      o.X.v[w] = __shfl_down_sync(0xffffffffu, acc.X.v[w], off);
      o.Y.v[w] = __shfl_down_sync(0xffffffffu, acc.Y.v[w], off);
      o.Z.v[w] = __shfl_down_sync(0xffffffffu, acc.Z.v[w], off);
      o.T.v[w] = __shfl_down_sync(0xffffffffu, acc.T.v[w], off);
    }
    acc = ge_add(acc, o);
  }
  if (live && lane == 0 && out_ext) ge_store_ext(out_ext + (size_t)i * 32, acc);
  if (out_bytes) {
    // the total sits in lane 0: hand it to every lane, encode on the whole warp (fe16.cuh)
    __shared__ __align__(16) uint32_t sm[(128 / 32) * G16_WORDS];
#pragma unroll
    for (int w = 0; w < 8; w++) {
      acc.X.v[w] = __shfl_sync(0xffffffffu, acc.X.v[w], 0);
      acc.Y.v[w] = __shfl_sync(0xffffffffu, acc.Y.v[w], 0);
      acc.Z.v[w] = __shfl_sync(0xffffffffu, acc.Z.v[w], 0);
      acc.T.v[w] = __shfl_sync(0xffffffffu, acc.T.v[w], 0);
    }
    grp16 g = warp_group(sm + (threadIdx.x >> 5) * G16_WORDS);
    fe s = ge_encode16<true>(g, acc);
    if (live && lane < 16) store_s_bytes(out_bytes + (size_t)i * 32, s, g.k);
  }
}

// ---------------------------------------------------------------------------
// Generator chains (reference src/generators.rs:107-125, 210-235): point i of a
// chain = element derivation (RFC 9496 §4.3.4) of the i-th 64-byte block of the chain's XOF
// stream.  The stream is squeezed on the host (sequential, ~1 GB/s); the two Elligator maps, the
// addition and the encoding (three inverse-square-root chains, ~900 field products per point)
// run here, one thread per point.
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(128) k_from_uniform(const uint8_t* __restrict__ in /*n*64*/, uint32_t n,
                                                       uint8_t* __restrict__ out /*n*32*/) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ge_ext p = ge_add(ge_elligator_map(fe_from_bytes_255(in + (size_t)i * 64)),
                    ge_elligator_map(fe_from_bytes_255(in + (size_t)i * 64 + 32)));
  ge_encode(out + (size_t)i * 32, p);
}

// ---------------------------------------------------------------------------
// windowed tables: out[w][i] = 2^(c w) * P_i in affine Niels, w < W.
// One thread per point walks the doubling chain, parks the extended multiples and
// the running product of their Z in scratch, inverts once (Montgomery's trick) and
// converts every multiple back to affine.  One-time cost at table upload; it removes
// all doublings from every later MSM over the table.
// ---------------------------------------------------------------------------
BPG_DEF_CONST(K_INV2, 0xfffffff7u, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu,
              0x3fffffffu)  // (p+1)/2

static __global__ void __launch_bounds__(128) k_window_chain(const uint32_t* __restrict__ niels_in, uint32_t n_total,
                                                       uint32_t first, uint32_t count, int c, int W,
                                                       uint32_t* __restrict__ ext_scratch /*[W-1][count][32]*/,
                                                       uint32_t* __restrict__ zp_scratch /*[W-1][count][8]*/,
                                                       uint32_t* __restrict__ out /*[W][n_total][24]*/) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  uint32_t i = first + t;
  ge_niels q;
  ge_load_niels(q, niels_in + (size_t)i * NIELS_WORDS);
  ge_store_niels(out + (size_t)i * NIELS_WORDS, q);  // window 0
  ge_ext p;
  p.X = fe_sub(q.ypx, q.ymx);     // 2x
  p.Y = fe_add(q.ypx, q.ymx);     // 2y
  p.Z = fe_zero();
  p.Z.v[0] = 2;
  p.T = fe_mul(fe_mul(p.X, p.Y), fe_const(BPG_K(K_INV2)));  // XY/Z
  p.X = fe_mul(p.X, fe_one());    // tighten
  p.Y = fe_mul(p.Y, fe_one());
  fe zp = fe_one();
  for (int w = 1; w < W; w++) {
    for (int k = 0; k < c; k++) p = ge_dbl(p);
    zp = fe_mul(zp, p.Z);
    ge_store_ext(ext_scratch + ((size_t)(w - 1) * count + t) * 32, p);
    fe_store(zp_scratch + ((size_t)(w - 1) * count + t) * 8, zp);
  }
  fe inv = fe_invert(zp);
  for (int w = W - 1; w >= 1; w--) {
    ge_ext e;
    ge_load_ext(e, ext_scratch + ((size_t)(w - 1) * count + t) * 32);
    fe zi;
    if (w >= 2) {
      fe prev;
      fe_load(prev, zp_scratch + ((size_t)(w - 2) * count + t) * 8);
      zi = fe_mul(inv, prev);
    } else {
      zi = inv;
    }
    inv = fe_mul(inv, e.Z);
    fe x = fe_mul(e.X, zi), y = fe_mul(e.Y, zi);
    ge_store_niels(out + ((size_t)w * n_total + i) * NIELS_WORDS, ge_affine_to_niels(x, y));
  }
}

}  // namespace bpg
