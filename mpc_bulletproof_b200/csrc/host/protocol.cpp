// Host mirror of the reference's protocol layer on top of the CUDA engine:
// `InnerProductProof::{create, verification_scalars, verify, to_bytes, from_bytes}`
// (reference src/inner_product_proof.rs), `r1cs::Prover` / `r1cs::Verifier`
// (src/r1cs/prover.rs, src/r1cs/verifier.rs), `R1CSProof` bytes (src/r1cs/proof.rs),
// `PedersenGens` / `BulletproofGens` as resident tables (src/generators.rs).
//
// Everything group-valued is done by the device through include/bpgpu.h; what stays
// here is what the reference also does serially on the CPU: the constraint-system
// bookkeeping, the transcript, and O(n) scalar preparation.
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <new>
#include <vector>

#include <errno.h>
#include <sys/random.h>

#include "../../../include/bpgpu.h"
#include "merlin.hpp"
#include "sc_host.hpp"

using namespace bpg_host;

typedef std::array<uint8_t, 32> Bytes32;

// Growable array in page-locked host memory (trivially copyable T): the witness rows and the flat
// constraint terms are uploaded straight from it at PCIe rate.
template <typename T>
struct PinnedVec {
  T* p = nullptr;
  size_t n = 0, cap = 0;
  PinnedVec() = default;
  PinnedVec(const PinnedVec&) = delete;
  PinnedVec& operator=(const PinnedVec&) = delete;
  ~PinnedVec() { bpg_host_free(p); }
  void reserve(size_t want) {
    if (want <= cap) return;
    size_t nc = std::max<size_t>(want, std::max<size_t>(cap * 4, 4096));
    T* q = static_cast<T*>(bpg_host_alloc(nc * sizeof(T)));
    if (!q) throw std::bad_alloc();
    if (n) memcpy(q, p, n * sizeof(T));
    bpg_host_free(p);
    p = q;
    cap = nc;
  }
  void push_back(const T& x) {
    if (n == cap) reserve(n + 1);
    p[n++] = x;
  }
  size_t size() const { return n; }
  T* data() { return p; }
  const T* data() const { return p; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
};

// BPG_TRACE=1: per-stage wall-clock of prove/verify on stderr (development aid)
struct StageTimer {
  bool on;
  std::chrono::steady_clock::time_point t0;
  const char* what;
  explicit StageTimer(const char* w) : on(getenv("BPG_TRACE") != nullptr), t0(std::chrono::steady_clock::now()), what(w) {}
  void lap(const char* stage) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[bpg] %s %-22s %8.3f ms\n", what, stage, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

static inline void sc_bytes(const Scalar& s, uint8_t* out) { s.to_bytes(out); }
static std::vector<uint8_t> sc_vec_bytes(const std::vector<Scalar>& v) {
  std::vector<uint8_t> out(v.size() * 32);
  for (size_t i = 0; i < v.size(); i++) v[i].to_bytes(out.data() + 32 * i);
  return out;
}
static bool is_identity_enc(const uint8_t p[32]) {
  uint8_t acc = 0;
  for (int i = 0; i < 32; i++) acc |= p[i];
  return acc == 0;
}

// ---------------------------------------------------------------- transcript C ABI
struct bpg_transcript {
  Transcript t;
  explicit bpg_transcript(const uint8_t* l, size_t n) : t(l, n) {}
};
extern "C" bpg_transcript* bpg_transcript_new(const uint8_t* label, size_t len) {
  return new (std::nothrow) bpg_transcript(label, len);
}
extern "C" bpg_transcript* bpg_transcript_clone(const bpg_transcript* t) {
  return t ? new (std::nothrow) bpg_transcript(*t) : nullptr;
}
extern "C" void bpg_transcript_free(bpg_transcript* t) { delete t; }
extern "C" void bpg_transcript_append_message(bpg_transcript* t, const char* label, const uint8_t* msg, size_t len) {
  t->t.append_message(label, msg, len);
}
extern "C" void bpg_transcript_append_u64(bpg_transcript* t, const char* label, uint64_t v) { t->t.append_u64(label, v); }
extern "C" void bpg_transcript_challenge_bytes(bpg_transcript* t, const char* label, uint8_t* out, size_t len) {
  t->t.challenge_bytes(label, out, len);
}
extern "C" void bpg_transcript_challenge_scalar(bpg_transcript* t, const char* label, uint8_t out[32]) {
  t->t.challenge_scalar(label).to_bytes(out);
}

struct bpg_transcript_rng {
  TranscriptRng r;
  explicit bpg_transcript_rng(const TranscriptRng& x) : r(x) {}
};
extern "C" bpg_transcript_rng* bpg_transcript_build_rng(const bpg_transcript* t) {
  return t ? new (std::nothrow) bpg_transcript_rng(t->t.build_rng()) : nullptr;
}
extern "C" void bpg_transcript_rng_rekey_with_witness_bytes(bpg_transcript_rng* r, const char* label, const uint8_t* w,
                                                            size_t len) {
  r->r.rekey_with_witness_bytes(label, w, len);
}
extern "C" void bpg_transcript_rng_finalize(bpg_transcript_rng* r, const uint8_t random_bytes[32]) { r->r.finalize(random_bytes); }
extern "C" void bpg_transcript_rng_fill_bytes(bpg_transcript_rng* r, uint8_t* out, size_t len) { r->r.fill_bytes(out, len); }
extern "C" void bpg_transcript_rng_free(bpg_transcript_rng* r) { delete r; }

// ---------------------------------------------------------------- generators
// One windowed table [G (cap) | H (cap) | B | B_blinding] plus a comb for (B, B_blinding).
struct bpg_gens {
  bpg_ctx* ctx;
  size_t cap;
  bpg_table* table;
  bpg_comb* comb;
  Bytes32 B, Bb;
  size_t g_base() const { return 0; }
  size_t h_base() const { return cap; }
  size_t b_id() const { return 2 * cap; }
  size_t bb_id() const { return 2 * cap + 1; }
};

extern "C" int bpg_gens_new(bpg_ctx* ctx, const uint8_t* G, const uint8_t* H, size_t capacity, const uint8_t B[32],
                            const uint8_t B_blinding[32], bpg_gens** out) {
  if (!ctx || !B || !B_blinding || !out || (capacity && (!G || !H))) return BPG_ERR_ARG;
  std::vector<uint8_t> all((2 * capacity + 2) * 32);
  memcpy(all.data(), G, capacity * 32);
  memcpy(all.data() + capacity * 32, H, capacity * 32);
  memcpy(all.data() + 2 * capacity * 32, B, 32);
  memcpy(all.data() + (2 * capacity + 1) * 32, B_blinding, 32);
  bpg_gens* g = new (std::nothrow) bpg_gens();
  if (!g) return BPG_ERR_NOMEM;
  g->ctx = ctx;
  g->cap = capacity;
  g->table = nullptr;
  g->comb = nullptr;
  memcpy(g->B.data(), B, 32);
  memcpy(g->Bb.data(), B_blinding, 32);
  int rc = bpg_table_upload(ctx, all.data(), 2 * capacity + 2, &g->table);
  if (!rc) rc = bpg_table_set_windows(ctx, g->table, 0);
  if (!rc) {
    // combs for the inner-product rounds (48 KB per generator) when they fit the budget; without them the
    // rounds keep the bucket method throughout
    const char* e = getenv("BPG_COMB_MAX_GB");
    double max_gb = e ? atof(e) : 32.0;
    if ((double)(2 * capacity + 2) * 49152.0 <= max_gb * 1e9) {
      int rc2 = bpg_table_build_comb(ctx, g->table);
      if (rc2 && rc2 != BPG_ERR_NOMEM) rc = rc2;
    }
  }
  uint8_t bases[64];
  memcpy(bases, B, 32);
  memcpy(bases + 32, B_blinding, 32);
  if (!rc) rc = bpg_comb_create(ctx, bases, 2, &g->comb);
  if (rc) {
    if (g->table) bpg_table_free(g->table);
    if (g->comb) bpg_comb_free(g->comb);
    delete g;
    return rc;
  }
  *out = g;
  return BPG_OK;
}
// BulletproofGens::new(gens_capacity, ..) for one party (reference src/generators.rs:182-235) together
// with PedersenGens::default (src/generators.rs:61-71), ristretto255 instantiation: the chains
// "G" || u32le(party) and "H" || u32le(party) are squeezed here (a sequential XOF), the points are derived
// on the device (k_from_uniform), then the table is built as in bpg_gens_new.  G_out / H_out (capacity*32
// bytes each, may be NULL) and B_out / Bb_out (32 bytes, may be NULL) receive the compressed generators.
static const uint8_t RISTRETTO_BASEPOINT_COMPRESSED[32] = {
    0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
    0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};
extern "C" int bpg_gens_chain(bpg_ctx* ctx, const uint8_t* label, size_t label_len, size_t skip, size_t n, uint8_t* out) {
  if (!ctx || (n && !out) || (label_len && !label)) return BPG_ERR_ARG;
  bpg_host::KeccakSponge xof = bpg_host::shake256();
  xof.absorb((const uint8_t*)"GeneratorsChain", 15);
  xof.absorb(label, label_len);
  uint8_t discard[64];
  for (size_t i = 0; i < skip; i++) xof.squeeze(discard, 64);  // fast_forward (src/generators.rs:92-100)
  std::vector<uint8_t> stream(n * 64);
  xof.squeeze(stream.data(), n * 64);
  return bpg_points_from_uniform(ctx, stream.data(), n, out);
}
extern "C" int bpg_gens_derive(bpg_ctx* ctx, size_t gens_capacity, uint32_t party, uint8_t* G_out, uint8_t* H_out,
                               uint8_t* B_out, uint8_t* Bb_out, bpg_gens** out) {
  if (!ctx || !out) return BPG_ERR_ARG;
  std::vector<uint8_t> G(gens_capacity * 32), H(gens_capacity * 32);
  uint8_t label[5] = {'G', (uint8_t)party, (uint8_t)(party >> 8), (uint8_t)(party >> 16), (uint8_t)(party >> 24)};
  int rc = bpg_gens_chain(ctx, label, 5, 0, gens_capacity, G.data());
  if (rc) return rc;
  label[0] = 'H';
  rc = bpg_gens_chain(ctx, label, 5, 0, gens_capacity, H.data());
  if (rc) return rc;
  uint8_t digest[64], Bb[32];
  bpg_host::sha3_512(RISTRETTO_BASEPOINT_COMPRESSED, 32, digest);
  rc = bpg_points_from_uniform(ctx, digest, 1, Bb);
  if (rc) return rc;
  if (G_out) memcpy(G_out, G.data(), G.size());
  if (H_out) memcpy(H_out, H.data(), H.size());
  if (B_out) memcpy(B_out, RISTRETTO_BASEPOINT_COMPRESSED, 32);
  if (Bb_out) memcpy(Bb_out, Bb, 32);
  return bpg_gens_new(ctx, G.data(), H.data(), gens_capacity, RISTRETTO_BASEPOINT_COMPRESSED, Bb, out);
}
extern "C" void bpg_gens_free(bpg_gens* g) {
  if (!g) return;
  bpg_table_free(g->table);
  bpg_comb_free(g->comb);
  delete g;
}
extern "C" size_t bpg_gens_capacity(const bpg_gens* g) { return g ? g->cap : 0; }
extern "C" const bpg_table* bpg_gens_table(const bpg_gens* g) { return g ? g->table : nullptr; }

// PedersenGens::commit batched: out[i] = values[i]*B + blindings[i]*B_blinding (generators.rs:41-43)
extern "C" int bpg_pedersen_commit(bpg_ctx* ctx, const bpg_gens* g, const uint8_t* values, const uint8_t* blindings,
                                   size_t n, uint8_t* out) {
  if (!ctx || !g || !out || (n && (!values || !blindings))) return BPG_ERR_ARG;
  std::vector<uint8_t> sc(2 * n * 32);
  memcpy(sc.data(), values, n * 32);
  memcpy(sc.data() + n * 32, blindings, n * 32);
  return bpg_comb_mul(ctx, g->comb, sc.data(), n, out);
}

// ---------------------------------------------------------------- inner product proof
struct InnerProductProof {
  std::vector<Bytes32> L_vec, R_vec;
  Scalar a, b;

  size_t serialized_size() const { return (2 * L_vec.size() + 2) * 32; }
  void to_bytes(uint8_t* out) const {  // inner_product_proof.rs:388-397
    for (size_t i = 0; i < L_vec.size(); i++) {
      memcpy(out + 64 * i, L_vec[i].data(), 32);
      memcpy(out + 64 * i + 32, R_vec[i].data(), 32);
    }
    a.to_bytes(out + 64 * L_vec.size());
    b.to_bytes(out + 64 * L_vec.size() + 32);
  }
  // inner_product_proof.rs:418-455 (point validity is checked by the device when they are used)
  static int from_bytes(const uint8_t* s, size_t len, InnerProductProof* out) {
    if (len % 32 || len < 64) return BPG_ERR_DECODE;
    size_t num_elements = len / 32;
    if ((num_elements - 2) % 2) return BPG_ERR_DECODE;
    size_t lg_n = (num_elements - 2) / 2;
    if (lg_n >= 32) return BPG_ERR_DECODE;
    out->L_vec.resize(lg_n);
    out->R_vec.resize(lg_n);
    for (size_t i = 0; i < lg_n; i++) {
      memcpy(out->L_vec[i].data(), s + 64 * i, 32);
      memcpy(out->R_vec[i].data(), s + 64 * i + 32, 32);
    }
    if (!Scalar::from_bytes(s + 64 * lg_n, &out->a) || !Scalar::from_bytes(s + 64 * lg_n + 32, &out->b))
      return BPG_ERR_DECODE;
    return BPG_OK;
  }

  // the round loop of inner_product_proof.rs:49-193 around a device-resident state
  static int run_rounds(bpg_ipp* st, Transcript& tr, InnerProductProof* out) {
    int rc = BPG_OK;
    StageTimer tm("ipp");
    while (bpg_ipp_rounds_left(st)) {
      Bytes32 L, R;
      rc = bpg_ipp_round_LR(st, L.data(), R.data());
      if (rc) return rc;
      tm.lap("round L,R");
      out->L_vec.push_back(L);
      out->R_vec.push_back(R);
      tr.append_point("L", L.data());  // :119-120
      tr.append_point("R", R.data());
      Scalar u = tr.challenge_scalar("u");  // :122
      Scalar u_inv = u.invert();
      uint8_t ub[32], uib[32];
      u.to_bytes(ub);
      u_inv.to_bytes(uib);
      rc = bpg_ipp_round_fold(st, ub, uib);
      if (rc) return rc;
      tm.lap("challenge + fold");
    }
    uint8_t ab[32], bb[32];
    rc = bpg_ipp_finish(st, ab, bb);
    if (rc) return rc;
    tm.lap("finish");
    Scalar::from_bytes(ab, &out->a);
    Scalar::from_bytes(bb, &out->b);
    return BPG_OK;
  }

  // inner_product_proof.rs:254-298: transcript replay, challenges, batch inversion; the n-vector s
  // (:300-307) is produced on the device from (allinv, u_sq) by its closed form
  int verification_challenges(size_t n, Transcript& tr, std::vector<Scalar>& u_sq, std::vector<Scalar>& u_inv_sq,
                              Scalar& allinv) const {
    size_t lg_n = L_vec.size();
    if (lg_n >= 32) return BPG_ERR_VERIFY;
    if (n != ((size_t)1 << lg_n)) return BPG_ERR_VERIFY;
    tr.innerproduct_domain_sep(n);
    std::vector<Scalar> ch(lg_n), ch_inv(lg_n);
    for (size_t i = 0; i < lg_n; i++) {
      if (!tr.validate_and_append_point("L", L_vec[i].data())) return BPG_ERR_VERIFY;
      if (!tr.validate_and_append_point("R", R_vec[i].data())) return BPG_ERR_VERIFY;
      ch[i] = tr.challenge_scalar("u");
    }
    // batch inversion (Scalar::batch_inverse): one inversion + 3 lg n multiplications
    allinv = Scalar::one();
    if (lg_n) {
      std::vector<Scalar> pre(lg_n);
      Scalar acc = Scalar::one();
      for (size_t i = 0; i < lg_n; i++) {
        pre[i] = acc;
        acc = acc * ch[i];
      }
      Scalar inv = acc.invert();
      allinv = inv;
      for (size_t i = lg_n; i-- > 0;) {
        ch_inv[i] = inv * pre[i];
        inv = inv * ch[i];
      }
    }
    u_sq.resize(lg_n);
    u_inv_sq.resize(lg_n);
    for (size_t i = 0; i < lg_n; i++) {
      u_sq[i] = ch[i] * ch[i];
      u_inv_sq[i] = ch_inv[i] * ch_inv[i];
    }
    return BPG_OK;
  }
};

extern "C" int bpg_ipp_create(bpg_ctx* ctx, bpg_transcript* t, const uint8_t Q[32], const uint8_t* G_factors,
                              const uint8_t* H_factors, const bpg_table* G, size_t g_off, const bpg_table* H,
                              size_t h_off, size_t n, const uint8_t* a, const uint8_t* b, uint8_t* proof_out,
                              size_t proof_cap, size_t* proof_len) {
  if (!ctx || !t || !Q || !G || !H || !a || !b || !proof_out || !proof_len) return BPG_ERR_ARG;
  if (n == 0 || (n & (n - 1))) return BPG_ERR_POW2;  // assert!(n.is_power_of_two()) :69
  t->t.innerproduct_domain_sep(n);                   // :72
  bpg_ipp* st = nullptr;
  int rc = bpg_ipp_begin(ctx, G, g_off, H, h_off, n, Q, G_factors, H_factors, a, b, &st);
  if (rc) return rc;
  InnerProductProof proof;
  rc = InnerProductProof::run_rounds(st, t->t, &proof);
  bpg_ipp_free(st);
  if (rc) return rc;
  if (proof.serialized_size() > proof_cap) return BPG_ERR_ARG;
  proof.to_bytes(proof_out);
  *proof_len = proof.serialized_size();
  return BPG_OK;
}

// inner_product_proof.rs:317-372
extern "C" int bpg_ipp_verify(bpg_ctx* ctx, bpg_transcript* t, size_t n, const uint8_t* G_factors,
                              const uint8_t* H_factors, const uint8_t P[32], const uint8_t Q[32], const bpg_table* G,
                              size_t g_off, const bpg_table* H, size_t h_off, const uint8_t* proof, size_t proof_len) {
  if (!ctx || !t || !P || !Q || !G || !H || !proof) return BPG_ERR_ARG;
  InnerProductProof p;
  int rc = InnerProductProof::from_bytes(proof, proof_len, &p);
  if (rc) return rc;
  // the ad-hoc points [Q | L | R] of the check (:352-366) are known before their scalars: their doubling chains
  // start now, beside the transcript replay (bpg_adhoc_prefetch)
  size_t lg_n = p.L_vec.size();
  std::vector<uint8_t> pts((1 + 2 * lg_n) * 32);
  memcpy(pts.data(), Q, 32);
  for (size_t i = 0; i < lg_n; i++) {
    memcpy(pts.data() + 32 * (1 + i), p.L_vec[i].data(), 32);
    memcpy(pts.data() + 32 * (1 + lg_n + i), p.R_vec[i].data(), 32);
  }
  rc = bpg_adhoc_prefetch(ctx, pts.data(), 1 + 2 * lg_n);
  if (rc) return rc;
  std::vector<Scalar> u_sq, u_inv_sq;
  Scalar allinv;
  rc = p.verification_challenges(n, t->t, u_sq, u_inv_sq, allinv);
  if (rc) return rc;
  if (G_factors || H_factors)  // canonical scalars only (Scalar::from_canonical_bytes at the Rust boundary)
    for (size_t i = 0; i < n; i++) {
      if (G_factors && !Scalar::is_canonical(G_factors + 32 * i)) return BPG_ERR_DECODE;
      if (H_factors && !Scalar::is_canonical(H_factors + 32 * i)) return BPG_ERR_DECODE;
    }
  // ad-hoc terms [Q | L | R] with scalars [a*b | -u_sq | -u_inv_sq]; the 2n generator scalars
  // a s_i g_i, b s_i^-1 h_i (:335-351) are produced on the device
  std::vector<Scalar> sc;
  sc.reserve(1 + 2 * lg_n);
  sc.push_back(p.a * p.b);
  for (auto& x : u_sq) sc.push_back(-x);
  for (auto& x : u_inv_sq) sc.push_back(-x);
  bpg_ipp_verify_params vp;
  memset(&vp, 0, sizeof vp);
  for (size_t j = 0; j < lg_n; j++) memcpy(vp.u_sq[j], u_sq[j].v, 32);
  memcpy(vp.allinv, allinv.v, 32);
  memcpy(vp.a, p.a.v, 32);
  memcpy(vp.b, p.b.v, 32);
  vp.lg_n = (uint32_t)lg_n;
  vp.N = (uint32_t)n;
  uint8_t expect[32];
  std::vector<uint8_t> scb = sc_vec_bytes(sc);
  rc = bpg_ipp_verify_msm(ctx, G, g_off, H, h_off, pts.data(), scb.data(), 1 + 2 * lg_n, G_factors, H_factors, &vp,
                          expect);
  if (rc) return rc;
  return memcmp(expect, P, 32) == 0 ? BPG_OK : BPG_ERR_VERIFY;
}

// ---------------------------------------------------------------- constraint system
enum VarKind : uint8_t { V_LEFT = 1, V_RIGHT = 2, V_OUT = 3, V_COMMITTED = 4, V_ONE = 5, V_ZERO = 6 };
static inline bpg_var mkvar(VarKind k, uint64_t i) { return ((uint64_t)k << 56) | i; }
static inline VarKind var_kind(bpg_var v) { return (VarKind)(v >> 56); }
static inline uint64_t var_idx(bpg_var v) { return v & ((1ull << 56) - 1); }

struct Term {
  bpg_var var;
  Scalar coeff;
};
typedef std::vector<Term> LinComb;

// SplitMix64-seeded xoshiro256**, 64 bytes -> scalar mod l: the prover's blinding source
// (the reference draws from thread_rng, src/r1cs/prover.rs:435-445; here the seed is the input)
struct Xoshiro {
  uint64_t s[4];
  explicit Xoshiro(uint64_t seed) {
    for (int i = 0; i < 4; i++) {
      seed += 0x9E3779B97F4A7C15ULL;
      uint64_t z = seed;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
      s[i] = z ^ (z >> 31);
    }
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t result = rotl(s[1] * 5, 7) * 9;
    uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return result;
  }
  Scalar scalar() {
    uint8_t b[64];
    for (int i = 0; i < 8; i++) {
      uint64_t x = next();
      memcpy(b + 8 * i, &x, 8);
    }
    return Scalar::from_wide(b);
  }
};

// The prover's blinding source, consumed in the reference's draw order (prover.rs:457-462, 519-530,
// 621-625).  Production: merlin's TranscriptRng (transcript state + v_blinding witnesses + 32 bytes of
// external randomness, prover.rs:435-445); the two blinding vectors of a phase come from ONE 32-byte
// draw that keys a ChaCha20 stream expanded on the device (k_blind_vectors_chacha).  Tests and
// benches: xoshiro256** from a 64-bit seed (reproducible; not a secure source).
struct Blinder {
  bool keyed;
  Xoshiro xo;
  TranscriptRng rng;
  Blinder(uint64_t seed, const Transcript& tr) : keyed(false), xo(seed), rng(tr.build_rng()) {}
  Blinder(const Transcript& tr, const std::vector<Scalar>& v_blinding, const uint8_t random_bytes[32])
      : keyed(true), xo(0), rng(tr.build_rng()) {
    for (auto& vb : v_blinding) {
      uint8_t b[32];
      vb.to_bytes(b);
      rng.rekey_with_witness_bytes("v_blinding", b, 32);
    }
    rng.finalize(random_bytes);
  }
  Scalar scalar() { return keyed ? rng.scalar() : xo.scalar(); }
  // key of a phase's (s_L, s_R): 8 bytes of the seeded stream, or 32 bytes of the transcript RNG
  void vec_key(uint64_t* k64, uint8_t k32[32]) {
    if (keyed) rng.fill_bytes(k32, 32);
    else *k64 = xo.next();
  }
};

// 32 bytes from the operating system (what `thread_rng()` stands on): getrandom(2), /dev/urandom as a fallback
static bool os_random(uint8_t out[32]) {
  size_t got = 0;
  while (got < 32) {
    ssize_t r = getrandom(out + got, 32 - got, 0);
    if (r < 0) {
      if (errno == EINTR) continue;
      break;
    }
    got += (size_t)r;
  }
  if (got == 32) return true;
  FILE* f = fopen("/dev/urandom", "rb");
  if (!f) return false;
  bool ok = fread(out, 1, 32, f) == 32;
  fclose(f);
  return ok;
}

typedef int (*bpg_randomized_cb)(struct bpg_cs* cs, void* user);

struct bpg_cs {
  bool is_prover;
  bpg_ctx* ctx;
  const bpg_gens* gens;
  Transcript* tr;
  // constraints in flat (CSR-like) form, appended as the gadget code constrains: row q, variable
  // code = kind << 28 | index, Montgomery coefficient -- uploaded as is for the device flattening
  PinnedVec<uint32_t> t_code, t_row;
  PinnedVec<Scalar> t_coeff;
  PinnedVec<uint32_t> u_code, u_row;  // terms with coefficient +1 / -1 (bit 31 of the code): no coefficient stored
  size_t n_rows = 0;
  // prover
  PinnedVec<Scalar> a_L, a_R, a_O;
  std::vector<Scalar> v, v_blinding;
  // verifier
  size_t num_vars = 0;
  std::vector<Bytes32> V;
  std::vector<std::pair<bpg_randomized_cb, void*>> deferred;
  std::vector<std::shared_ptr<void>> owned;  // closures of native gadgets' deferred callbacks
  bool has_pending = false;
  size_t pending = 0;
  bool randomizing = false;

  size_t num_multipliers() const { return is_prover ? a_O.size() : num_vars; }

  Scalar eval(const LinComb& lc) const {  // prover.rs:178-194 (verifier: dummy zero, verifier.rs:168-174)
    Scalar tot = Scalar::zero();
    if (!is_prover) return tot;
    for (auto& t : lc) {
      uint64_t i = var_idx(t.var);
      switch (var_kind(t.var)) {
        case V_LEFT: tot += t.coeff * a_L[i]; break;
        case V_RIGHT: tot += t.coeff * a_R[i]; break;
        case V_OUT: tot += t.coeff * a_O[i]; break;
        case V_COMMITTED: tot += t.coeff * v[i]; break;
        case V_ONE: tot += t.coeff; break;
        default: break;
      }
    }
    return tot;
  }
  void add_constraint(const LinComb& lc) {
    static const Scalar one = Scalar::one(), minus_one = -Scalar::one();
    for (auto& t : lc) {
      const uint32_t code = ((uint32_t)var_kind(t.var) << 28) | (uint32_t)var_idx(t.var);
      if (t.coeff == one || t.coeff == minus_one) {
        u_code.push_back(t.coeff == one ? code : code | 0x80000000u);
        u_row.push_back((uint32_t)n_rows);
      } else {
        t_code.push_back(code);
        t_row.push_back((uint32_t)n_rows);
        t_coeff.push_back(t.coeff);
      }
    }
    n_rows++;
  }
  bool valid(const LinComb& lc) const {
    if (lc.size() >= (1u << 31)) return false;
    for (auto& t : lc) {
      uint64_t i = var_idx(t.var);
      switch (var_kind(t.var)) {
        case V_LEFT: case V_RIGHT: case V_OUT: if (i >= num_multipliers() || i >= (1u << 27)) return false; break;
        case V_COMMITTED: if (i >= (is_prover ? v.size() : V.size()) || i >= (1u << 27)) return false; break;
        case V_ONE: case V_ZERO: break;
        default: return false;
      }
    }
    return true;
  }
  // prover.rs:99-125 / verifier.rs:100-118
  void multiply(LinComb left, LinComb right, bpg_var out[3]) {
    size_t i;
    if (is_prover) {
      Scalar l = eval(left), r = eval(right);
      i = a_L.size();
      a_L.push_back(l);
      a_R.push_back(r);
      a_O.push_back(l * r);
    } else {
      i = num_vars++;
    }
    out[0] = mkvar(V_LEFT, i);
    out[1] = mkvar(V_RIGHT, i);
    out[2] = mkvar(V_OUT, i);
    left.push_back({out[0], -Scalar::one()});
    right.push_back({out[1], -Scalar::one()});
    add_constraint(left);
    add_constraint(right);
  }
  // prover.rs:127-146 / verifier.rs:120-134
  bpg_var allocate(const Scalar* assignment) {
    if (!has_pending) {
      size_t i;
      if (is_prover) {
        i = a_L.size();
        a_L.push_back(*assignment);
        a_R.push_back(Scalar::zero());
        a_O.push_back(Scalar::zero());
      } else {
        i = num_vars++;
      }
      has_pending = true;
      pending = i;
      return mkvar(V_LEFT, i);
    }
    size_t i = pending;
    has_pending = false;
    if (is_prover) {
      a_R[i] = *assignment;
      a_O[i] = a_L[i] * a_R[i];
    }
    return mkvar(V_RIGHT, i);
  }
  // prover.rs:148-164 / verifier.rs:136-150
  void allocate_multiplier(const Scalar* l, const Scalar* r, bpg_var out[3]) {
    size_t i;
    if (is_prover) {
      i = a_L.size();
      a_L.push_back(*l);
      a_R.push_back(*r);
      a_O.push_back(*l * *r);
    } else {
      i = num_vars++;
    }
    out[0] = mkvar(V_LEFT, i);
    out[1] = mkvar(V_RIGHT, i);
    out[2] = mkvar(V_OUT, i);
  }

  // prover.rs:342-379 / verifier.rs:323-362: sum_q z^(q+1) * row_q, evaluated on the device as a
  // sparse product over the flat terms; wL, wR, wO stay in HBM, wV and wc come back
  // the terms known so far start their way to the device (they depend on no challenge); constraints added
  // later (second phase) change the length and flattened_constraints uploads the whole list itself
  bpg_terms terms() const {
    bpg_terms T;
    T.n_terms = t_code.size();
    T.t_code = t_code.data();
    T.t_row = t_row.data();
    T.t_coeff = t_coeff.data();
    T.n_unit = u_code.size();
    T.u_code = u_code.data();
    T.u_row = u_row.data();
    return T;
  }
  int prefetch_terms(bool after_commit_uploads) const {
    bpg_terms T = terms();
    return bpg_r1cs_terms_prefetch(ctx, &T, after_commit_uploads);
  }
  int flattened_constraints(bpg_r1cs_dev* dv, const Scalar& z, std::vector<Scalar>& wV, Scalar& wc) const {
    size_t n = num_multipliers(), m = is_prover ? v.size() : V.size();
    uint32_t z_pow[32][8];
    Scalar b = z;
    for (int k = 0; k < 32; k++) {
      memcpy(z_pow[k], b.v, 32);
      b = b * b;
    }
    std::vector<Scalar> out(m + 1);
    bpg_terms T = terms();
    int rc = bpg_r1cs_dev_flatten_terms(dv, n, m, &T, z_pow, out.data());
    if (rc) return rc;
    wV.assign(out.begin(), out.begin() + m);
    wc = out[m];  // the prover ignores it (prover.rs:370-372)
    return BPG_OK;
  }

  // prover.rs:383-402 / verifier.rs:366-385
  int create_randomized_constraints() {
    has_pending = false;
    if (deferred.empty()) {
      tr->r1cs_1phase_domain_sep();
      return BPG_OK;
    }
    // the callbacks append constraints: a prefetch of the first-phase terms must have left the arrays (they may move)
    int wrc = bpg_r1cs_terms_wait(ctx);
    if (wrc) return wrc;
    tr->r1cs_2phase_domain_sep();
    randomizing = true;
    auto cbs = std::move(deferred);
    deferred.clear();
    for (auto& cb : cbs) {
      int rc = cb.first(this, cb.second);
      if (rc) return rc;
    }
    return BPG_OK;
  }
};

static size_t next_pow2(size_t n) {
  size_t p = 1;
  while (p < n) p <<= 1;
  return p;
}

static int parse_lc(const bpg_cs* cs, const bpg_term* t, size_t n, LinComb* out) {
  out->resize(n);
  for (size_t i = 0; i < n; i++) {
    (*out)[i].var = t[i].var;
    if (!Scalar::from_bytes(t[i].coeff, &(*out)[i].coeff)) return BPG_ERR_DECODE;
  }
  return cs->valid(*out) ? BPG_OK : BPG_ERR_ARG;
}

extern "C" int bpg_prover_new(bpg_ctx* ctx, const bpg_gens* gens, bpg_transcript* t, bpg_cs** out) {
  if (!ctx || !gens || !t || !out) return BPG_ERR_ARG;
  bpg_cs* cs = new (std::nothrow) bpg_cs();
  if (!cs) return BPG_ERR_NOMEM;
  cs->is_prover = true;
  cs->ctx = ctx;
  cs->gens = gens;
  cs->tr = &t->t;
  cs->tr->r1cs_domain_sep();  // prover.rs:286
  *out = cs;
  return BPG_OK;
}
extern "C" int bpg_verifier_new(bpg_ctx* ctx, const bpg_gens* gens, bpg_transcript* t, bpg_cs** out) {
  if (!ctx || !gens || !t || !out) return BPG_ERR_ARG;
  bpg_cs* cs = new (std::nothrow) bpg_cs();
  if (!cs) return BPG_ERR_NOMEM;
  cs->is_prover = false;
  cs->ctx = ctx;
  cs->gens = gens;
  cs->tr = &t->t;
  cs->tr->r1cs_domain_sep();  // verifier.rs:271
  *out = cs;
  return BPG_OK;
}
extern "C" void bpg_cs_free(bpg_cs* cs) { delete cs; }
extern "C" size_t bpg_cs_num_multipliers(const bpg_cs* cs) { return cs ? cs->num_multipliers() : 0; }
extern "C" size_t bpg_cs_num_constraints(const bpg_cs* cs) { return cs ? cs->n_rows : 0; }
extern "C" bpg_var bpg_var_one(void) { return mkvar(V_ONE, 0); }

// prover.rs:319-329
extern "C" int bpg_prover_commit(bpg_cs* cs, const uint8_t v[32], const uint8_t v_blinding[32], uint8_t V_out[32],
                                 bpg_var* var) {
  if (!cs || !cs->is_prover || !v || !v_blinding || !V_out || !var) return BPG_ERR_ARG;
  Scalar sv, sb;
  if (!Scalar::from_bytes(v, &sv) || !Scalar::from_bytes(v_blinding, &sb)) return BPG_ERR_DECODE;
  int rc = bpg_pedersen_commit(cs->ctx, cs->gens, v, v_blinding, 1, V_out);
  if (rc) return rc;
  size_t i = cs->v.size();
  cs->v.push_back(sv);
  cs->v_blinding.push_back(sb);
  cs->tr->append_point("V", V_out);
  *var = mkvar(V_COMMITTED, i);
  return BPG_OK;
}
// n calls of Prover::commit in one: ONE batched fixed-base launch for the n commitments, then the transcript
// appends in order -- the same transcript and variables as committing one by one (prover.rs:319-329), without a
// launch and a synchronisation per value (a shuffle of 64 values commits 128 of them).
extern "C" int bpg_prover_commit_batch(bpg_cs* cs, const uint8_t* v, const uint8_t* v_blinding, size_t n, uint8_t* V_out,
                                       bpg_var* vars) {
  if (!cs || !cs->is_prover || (n && (!v || !v_blinding || !V_out || !vars))) return BPG_ERR_ARG;
  std::vector<Scalar> sv(n), sb(n);
  for (size_t i = 0; i < n; i++)
    if (!Scalar::from_bytes(v + 32 * i, &sv[i]) || !Scalar::from_bytes(v_blinding + 32 * i, &sb[i])) return BPG_ERR_DECODE;
  int rc = bpg_pedersen_commit(cs->ctx, cs->gens, v, v_blinding, n, V_out);
  if (rc) return rc;
  for (size_t i = 0; i < n; i++) {
    size_t j = cs->v.size();
    cs->v.push_back(sv[i]);
    cs->v_blinding.push_back(sb[i]);
    cs->tr->append_point("V", V_out + 32 * i);
    vars[i] = mkvar(V_COMMITTED, j);
  }
  return BPG_OK;
}
// verifier.rs:298-308
extern "C" int bpg_verifier_commit(bpg_cs* cs, const uint8_t V[32], bpg_var* var) {
  if (!cs || cs->is_prover || !V || !var) return BPG_ERR_ARG;
  Bytes32 b;
  memcpy(b.data(), V, 32);
  size_t i = cs->V.size();
  cs->V.push_back(b);
  cs->tr->append_point("V", V);
  *var = mkvar(V_COMMITTED, i);
  return BPG_OK;
}
// prover.rs:169-171 / verifier.rs:153-160: blinding factor one
extern "C" int bpg_cs_commit_public(bpg_cs* cs, const uint8_t value[32], bpg_var* var) {
  if (!cs || !value || !var) return BPG_ERR_ARG;
  uint8_t one[32] = {1};
  uint8_t V[32];
  if (cs->is_prover) return bpg_prover_commit(cs, value, one, V, var);
  int rc = bpg_pedersen_commit(cs->ctx, cs->gens, value, one, 1, V);
  if (rc) return rc;
  return bpg_verifier_commit(cs, V, var);
}
extern "C" int bpg_cs_multiply(bpg_cs* cs, const bpg_term* left, size_t nl, const bpg_term* right, size_t nr,
                               bpg_var out[3]) {
  if (!cs || !out) return BPG_ERR_ARG;
  LinComb l, r;
  int rc = parse_lc(cs, left, nl, &l);
  if (!rc) rc = parse_lc(cs, right, nr, &r);
  if (rc) return rc;
  cs->multiply(std::move(l), std::move(r), out);
  return BPG_OK;
}
extern "C" int bpg_cs_allocate(bpg_cs* cs, const uint8_t* assignment, bpg_var* out) {
  if (!cs || !out) return BPG_ERR_ARG;
  Scalar a = Scalar::zero();
  if (cs->is_prover) {
    if (!assignment) return BPG_ERR_ARG;  // R1CSError::MissingAssignment
    if (!Scalar::from_bytes(assignment, &a)) return BPG_ERR_DECODE;
  }
  *out = cs->allocate(&a);
  return BPG_OK;
}
extern "C" int bpg_cs_allocate_multiplier(bpg_cs* cs, const uint8_t* l, const uint8_t* r, bpg_var out[3]) {
  if (!cs || !out) return BPG_ERR_ARG;
  Scalar sl = Scalar::zero(), sr = Scalar::zero();
  if (cs->is_prover) {
    if (!l || !r) return BPG_ERR_ARG;  // R1CSError::MissingAssignment
    if (!Scalar::from_bytes(l, &sl) || !Scalar::from_bytes(r, &sr)) return BPG_ERR_DECODE;
  }
  cs->allocate_multiplier(&sl, &sr, out);
  return BPG_OK;
}
extern "C" int bpg_cs_constrain(bpg_cs* cs, const bpg_term* lc, size_t n) {
  if (!cs) return BPG_ERR_ARG;
  LinComb l;
  int rc = parse_lc(cs, lc, n, &l);
  if (rc) return rc;
  cs->add_constraint(l);
  return BPG_OK;
}
extern "C" int bpg_cs_specify_randomized_constraints(bpg_cs* cs, bpg_randomized_cb cb, void* user) {
  if (!cs || !cb) return BPG_ERR_ARG;
  cs->deferred.push_back({cb, user});
  return BPG_OK;
}
extern "C" int bpg_cs_challenge_scalar(bpg_cs* cs, const char* label, uint8_t out[32]) {
  if (!cs || !label || !out || !cs->randomizing) return BPG_ERR_ARG;
  cs->tr->challenge_scalar(label).to_bytes(out);
  return BPG_OK;
}
// the prover's assignment of a linear combination (ConstraintSystem::eval)
extern "C" int bpg_cs_eval(bpg_cs* cs, const bpg_term* lc, size_t n, uint8_t out[32]) {
  if (!cs || !out) return BPG_ERR_ARG;
  LinComb l;
  int rc = parse_lc(cs, lc, n, &l);
  if (rc) return rc;
  cs->eval(l).to_bytes(out);
  return BPG_OK;
}

// The reference's benchmark circuit (benches/r1cs.rs:24-32, DummyCircuit): starting from `var`,
// n chained squarings  var <- var * var.  Native so that building 2^16 multipliers is not
// dominated by per-call binding overhead; semantically n calls of bpg_cs_multiply.
extern "C" int bpg_gadget_square_chain(bpg_cs* cs, bpg_var var, size_t n, bpg_var* out) {
  if (!cs) return BPG_ERR_ARG;
  LinComb l(1);
  l[0].var = var;
  l[0].coeff = Scalar::one();
  if (!cs->valid(l)) return BPG_ERR_ARG;
  bpg_var o[3] = {var, var, var};
  for (size_t i = 0; i < n; i++) {
    LinComb a(1), b(1);
    a[0].var = b[0].var = o[2];
    a[0].coeff = b[0].coeff = Scalar::one();
    cs->multiply(std::move(a), std::move(b), o);
  }
  if (out) *out = o[2];
  return BPG_OK;
}

// The shuffle gadget of the reference's benches and tests (benches/shuffle.rs:30-69, tests/r1cs.rs:22-63): y is a
// permutation of x through the randomized product check prod (x_i - z) = prod (y_i - z).  Native for the same
// reason as the square chain: 2 (k - 1) multipliers through a foreign-function callback cost more than proving
// them (config 1, k = 64: 1.2 of 2.5 ms were the Python callback).  Same constraint system as the callback form.
struct ShuffleClosure {
  std::vector<bpg_var> x, y;
};
static LinComb lc_var_minus(bpg_var v, const Scalar& z) {
  LinComb l(2);
  l[0].var = v;
  l[0].coeff = Scalar::one();
  l[1].var = mkvar(V_ONE, 0);
  l[1].coeff = -z;
  return l;
}
static int shuffle_randomized(bpg_cs* cs, void* user) {
  const ShuffleClosure& c = *static_cast<const ShuffleClosure*>(user);
  const size_t k = c.x.size();
  const Scalar z = cs->tr->challenge_scalar("shuffle challenge");
  auto product = [&](const std::vector<bpg_var>& v) {
    bpg_var o[3];
    cs->multiply(lc_var_minus(v[k - 1], z), lc_var_minus(v[k - 2], z), o);
    for (size_t i = k - 2; i-- > 0;) {
      LinComb prev(1);
      prev[0].var = o[2];
      prev[0].coeff = Scalar::one();
      cs->multiply(std::move(prev), lc_var_minus(v[i], z), o);
    }
    return o[2];
  };
  const bpg_var px = product(c.x), py = product(c.y);
  LinComb d(2);
  d[0].var = px;
  d[0].coeff = Scalar::one();
  d[1].var = py;
  d[1].coeff = -Scalar::one();
  cs->add_constraint(d);
  return BPG_OK;
}
extern "C" int bpg_gadget_shuffle(bpg_cs* cs, const bpg_var* x, const bpg_var* y, size_t k) {
  if (!cs || !x || !y || k == 0) return BPG_ERR_ARG;
  LinComb probe(2 * k);
  for (size_t i = 0; i < k; i++) {
    probe[i].var = x[i];
    probe[k + i].var = y[i];
    probe[i].coeff = probe[k + i].coeff = Scalar::one();
  }
  if (!cs->valid(probe)) return BPG_ERR_ARG;
  if (k == 1) {
    LinComb d(2);
    d[0].var = y[0];
    d[0].coeff = Scalar::one();
    d[1].var = x[0];
    d[1].coeff = -Scalar::one();
    cs->add_constraint(d);
    return BPG_OK;
  }
  auto c = std::make_shared<ShuffleClosure>();
  c->x.assign(x, x + k);
  c->y.assign(y, y + k);
  cs->owned.push_back(c);
  cs->deferred.push_back({shuffle_randomized, c.get()});
  return BPG_OK;
}

// BASELINE.json config 4 / SURVEY.md 8d: the synthetic random circuit.  n_mult multipliers with uniform
// a_L, a_R and n_cons constraints  c1 a_L[i1] + c2 a_R[i2] + c3 a_O[i3] + c4 v[j] - c0 = 0  over the
// variables committed so far; indices and coefficients from xoshiro256**(seed) in a fixed draw order
// (oracle/gadgets.py random_circuit is the same definition).  The prover writes the public constants
// c0 (n_cons x 32 bytes, fixed from its witness), the verifier reads them.
extern "C" int bpg_gadget_random_circuit(bpg_cs* cs, uint64_t seed, size_t n_mult, size_t n_cons, uint8_t* c0) {
  if (!cs || !c0 || n_mult == 0) return BPG_ERR_ARG;
  size_t m = cs->is_prover ? cs->v.size() : cs->V.size();
  if (m == 0) return BPG_ERR_ARG;
  Xoshiro rng(seed);
  size_t base = cs->num_multipliers();
  for (size_t i = 0; i < n_mult; i++) {
    Scalar l = rng.scalar(), r = rng.scalar();
    bpg_var o[3];
    cs->allocate_multiplier(&l, &r, o);
  }
  const Scalar minus_one = -Scalar::one();
  for (size_t q = 0; q < n_cons; q++) {
    uint64_t i1 = rng.next() % n_mult, i2 = rng.next() % n_mult, i3 = rng.next() % n_mult, j = rng.next() % m;
    Scalar c1 = rng.scalar(), c2 = rng.scalar(), c3 = rng.scalar(), c4 = rng.scalar(), k0;
    if (cs->is_prover) {
      k0 = c1 * cs->a_L[base + i1] + c2 * cs->a_R[base + i2] + c3 * cs->a_O[base + i3] + c4 * cs->v[j];
      k0.to_bytes(c0 + 32 * q);
    } else if (!Scalar::from_bytes(c0 + 32 * q, &k0)) {
      return BPG_ERR_DECODE;
    }
    LinComb lc(5);
    lc[0] = {mkvar(V_LEFT, base + i1), c1};
    lc[1] = {mkvar(V_RIGHT, base + i2), c2};
    lc[2] = {mkvar(V_OUT, base + i3), c3};
    lc[3] = {mkvar(V_COMMITTED, j), c4};
    lc[4] = {mkvar(V_ONE, 0), minus_one * k0};
    cs->add_constraint(lc);
  }
  return BPG_OK;
}

// ---------------------------------------------------------------- R1CS proof bytes
struct R1CSProof {
  Bytes32 A_I1, A_O1, S1, A_I2, A_O2, S2, T_1, T_3, T_4, T_5, T_6;
  Scalar t_x, t_x_blinding, e_blinding;
  InnerProductProof ipp;

  bool missing_phase2() const {  // proof.rs:121-123
    return is_identity_enc(A_I2.data()) && is_identity_enc(A_O2.data()) && is_identity_enc(S2.data());
  }
  size_t serialized_size() const { return 1 + (missing_phase2() ? 11 : 14) * 32 + ipp.serialized_size(); }
  void to_bytes(uint8_t* out) const {  // proof.rs:82-108
    uint8_t* p = out;
    auto put = [&](const Bytes32& b) { memcpy(p, b.data(), 32); p += 32; };
    if (missing_phase2()) {
      *p++ = 0;
      put(A_I1); put(A_O1); put(S1);
    } else {
      *p++ = 1;
      put(A_I1); put(A_O1); put(S1); put(A_I2); put(A_O2); put(S2);
    }
    put(T_1); put(T_3); put(T_4); put(T_5); put(T_6);
    t_x.to_bytes(p); p += 32;
    t_x_blinding.to_bytes(p); p += 32;
    e_blinding.to_bytes(p); p += 32;
    ipp.to_bytes(p);
  }
  static int from_bytes(const uint8_t* s, size_t len, R1CSProof* out) {  // proof.rs:128-207
    if (len == 0) return BPG_ERR_DECODE;
    uint8_t version = s[0];
    s++;
    len--;
    if (len % 32) return BPG_ERR_DECODE;
    size_t npts;
    if (version == 0) npts = 8;
    else if (version == 1) npts = 11;
    else return BPG_ERR_DECODE;
    if (len < (npts + 3) * 32) return BPG_ERR_DECODE;
    auto get = [&](Bytes32& b) { memcpy(b.data(), s, 32); s += 32; len -= 32; };
    get(out->A_I1); get(out->A_O1); get(out->S1);
    if (version == 0) {
      out->A_I2.fill(0); out->A_O2.fill(0); out->S2.fill(0);
    } else {
      get(out->A_I2); get(out->A_O2); get(out->S2);
    }
    get(out->T_1); get(out->T_3); get(out->T_4); get(out->T_5); get(out->T_6);
    if (!Scalar::from_bytes(s, &out->t_x) || !Scalar::from_bytes(s + 32, &out->t_x_blinding) ||
        !Scalar::from_bytes(s + 64, &out->e_blinding))
      return BPG_ERR_DECODE;
    s += 96;
    len -= 96;
    return InnerProductProof::from_bytes(s, len, &out->ipp);
  }
};

// ---------------------------------------------------------------- Prover::prove (prover.rs:412-727)
// base^(2^k), k < 32, Montgomery limbs: lets a device thread form base^i in <= lg i products
static void pow_table(const Scalar& base, uint32_t out[32][8]) {
  Scalar b = base;
  for (int k = 0; k < 32; k++) {
    memcpy(out[k], b.v, 32);
    b = b * b;
  }
}
struct DevGuard {  // frees the resident prover state on every exit path
  bpg_r1cs_dev* p = nullptr;
  ~DevGuard() { bpg_r1cs_dev_free(p); }
};

// a prefetch of the term arrays must not outlive the call that issued it (the arrays belong to the constraint
// system, which the caller may free right after an early return); after a flatten there is nothing to wait for
struct TermsGuard {
  bpg_ctx* ctx;
  ~TermsGuard() { bpg_r1cs_terms_wait(ctx); }
};

static int prover_prove(bpg_cs* cs, bool keyed, uint64_t rng_seed, const uint8_t* random_bytes, uint8_t* proof_out,
                        size_t proof_cap, size_t* proof_len) {
  if (!cs || !cs->is_prover || !proof_out || !proof_len) return BPG_ERR_ARG;
  Transcript& tr = *cs->tr;
  const bpg_gens* g = cs->gens;
  R1CSProof proof;
  StageTimer tm("prove");
  tr.append_u64("m", cs->v.size());  // :420
  // :435-445: the RNG is forked from the transcript here, keyed with the v_blindings and external randomness
  Blinder rng = keyed ? Blinder(tr, cs->v_blinding, random_bytes) : Blinder(rng_seed, tr);
  size_t n1 = cs->a_L.size();
  if (g->cap < n1) return BPG_ERR_CAPACITY;  // :450-452
  // Blinding draws in the reference's order (:457-462); the 2 n1 vector blindings s_L, s_R are
  // generated on the device from one drawn key (none drawn for an empty phase)
  Scalar i_b1 = rng.scalar(), o_b1 = rng.scalar(), s_b1 = rng.scalar();
  uint64_t vec_key = 0;
  uint8_t vec_key32[32] = {0};
  if (n1) rng.vec_key(&vec_key, vec_key32);
  auto commit = [&](bpg_r1cs_dev* dvp, size_t first, size_t cnt, const uint8_t* b3, uint8_t* out3) {
    return keyed ? bpg_r1cs_dev_commit_keyed(dvp, g->table, g->g_base(), g->h_base(), g->bb_id(), first, cnt,
                                             cs->a_L.data() + first, cs->a_R.data() + first, cs->a_O.data() + first,
                                             vec_key32, b3, out3)
                 : bpg_r1cs_dev_commit(dvp, g->table, g->g_base(), g->h_base(), g->bb_id(), first, cnt,
                                       cs->a_L.data() + first, cs->a_R.data() + first, cs->a_O.data() + first, vec_key, b3,
                                       out3);
  };
  DevGuard dv;
  int rc = bpg_r1cs_dev_new(cs->ctx, next_pow2(std::max<size_t>(n1, 1)), &dv.p);
  if (rc) return rc;
  uint8_t c3[96], blind3[96];
  TermsGuard terms_guard{cs->ctx};
  rc = cs->prefetch_terms(n1 != 0);  // behind the witness rows of the first commitment when there is one
  if (rc) return rc;
  tm.lap("blindings s_L s_R");
  i_b1.to_bytes(blind3);
  o_b1.to_bytes(blind3 + 32);
  s_b1.to_bytes(blind3 + 64);
  rc = commit(dv.p, 0, n1, blind3, c3);  // :465-494
  if (rc) return rc;
  tm.lap("A_I1 A_O1 S1 msm");
  memcpy(proof.A_I1.data(), c3, 32);
  memcpy(proof.A_O1.data(), c3 + 32, 32);
  memcpy(proof.S1.data(), c3 + 64, 32);
  tr.append_point("A_I1", proof.A_I1.data());
  tr.append_point("A_O1", proof.A_O1.data());
  tr.append_point("S1", proof.S1.data());
  rc = cs->create_randomized_constraints();  // :501
  if (rc) return rc;
  size_t n = cs->a_L.size(), n2 = n - n1, padded_n = next_pow2(n);
  if (g->cap < padded_n) return BPG_ERR_CAPACITY;  // :511-513
  rc = bpg_r1cs_dev_reserve(&dv.p, padded_n);
  if (rc) return rc;
  Scalar i_b2 = Scalar::zero(), o_b2 = Scalar::zero(), s_b2 = Scalar::zero();
  if (n2 > 0) {  // :519-530
    i_b2 = rng.scalar();
    o_b2 = rng.scalar();
    s_b2 = rng.scalar();
    rng.vec_key(&vec_key, vec_key32);
    i_b2.to_bytes(blind3);
    o_b2.to_bytes(blind3 + 32);
    s_b2.to_bytes(blind3 + 64);
    rc = commit(dv.p, n1, n2, blind3, c3);  // :532-565
    if (rc) return rc;
    memcpy(proof.A_I2.data(), c3, 32);
    memcpy(proof.A_O2.data(), c3 + 32, 32);
    memcpy(proof.S2.data(), c3 + 64, 32);
  } else {
    proof.A_I2.fill(0);  // identity (:566-576)
    proof.A_O2.fill(0);
    proof.S2.fill(0);
  }
  tr.append_point("A_I2", proof.A_I2.data());
  tr.append_point("A_O2", proof.A_O2.data());
  tr.append_point("S2", proof.S2.data());
  Scalar y = tr.challenge_scalar("y"), z = tr.challenge_scalar("z");  // :584-585
  std::vector<Scalar> wV;
  Scalar wc;
  tm.lap("phase 2");
  rc = cs->flattened_constraints(dv.p, z, wV, wc);
  if (rc) return rc;
  tm.lap("flattened_constraints");
  // l(X), r(X) coefficient vectors (:596-617) and t_1..t_6 (util.rs:152-170) in HBM
  Scalar y_inv = y.invert();
  uint32_t y_pow[32][8], y_inv_pow[32][8];
  pow_table(y, y_pow);
  pow_table(y_inv, y_inv_pow);
  uint8_t tbytes[192];
  rc = bpg_r1cs_dev_poly_t(dv.p, n, y_pow, y_inv_pow, tbytes);
  if (rc) return rc;
  Scalar t[6];
  for (int k = 0; k < 6; k++) Scalar::from_bytes(tbytes + 32 * k, &t[k]);
  tm.lap("l/r polys + t_i");
  Scalar tb1 = rng.scalar(), tb3 = rng.scalar(), tb4 = rng.scalar(), tb5 = rng.scalar(), tb6 = rng.scalar();  // :621-625
  {
    uint8_t vals[160], blinds[160], Ts[160];
    memcpy(vals, tbytes, 32);            // t_1
    memcpy(vals + 32, tbytes + 64, 128);  // t_3 .. t_6
    tb1.to_bytes(blinds); tb3.to_bytes(blinds + 32); tb4.to_bytes(blinds + 64); tb5.to_bytes(blinds + 96);
    tb6.to_bytes(blinds + 128);
    rc = bpg_pedersen_commit(cs->ctx, g, vals, blinds, 5, Ts);  // :627-631
    if (rc) return rc;
    memcpy(proof.T_1.data(), Ts, 32);
    memcpy(proof.T_3.data(), Ts + 32, 32);
    memcpy(proof.T_4.data(), Ts + 64, 32);
    memcpy(proof.T_5.data(), Ts + 96, 32);
    memcpy(proof.T_6.data(), Ts + 128, 32);
  }
  tr.append_point("T_1", proof.T_1.data());
  tr.append_point("T_3", proof.T_3.data());
  tr.append_point("T_4", proof.T_4.data());
  tr.append_point("T_5", proof.T_5.data());
  tr.append_point("T_6", proof.T_6.data());
  Scalar u = tr.challenge_scalar("u"), x = tr.challenge_scalar("x");  // :639-640
  Scalar tb2 = Scalar::zero();
  for (size_t i = 0; i < wV.size(); i++) tb2 += wV[i] * cs->v_blinding[i];  // :644-648
  auto poly6 = [&](const Scalar& c1, const Scalar& c2, const Scalar& c3_, const Scalar& c4, const Scalar& c5,
                   const Scalar& c6) { return x * (c1 + x * (c2 + x * (c3_ + x * (c4 + x * (c5 + x * c6))))); };
  proof.t_x = poly6(t[0], t[1], t[2], t[3], t[4], t[5]);
  proof.t_x_blinding = poly6(tb1, tb2, tb3, tb4, tb5, tb6);
  Scalar i_b = i_b1 + u * i_b2, o_b = o_b1 + u * o_b2, s_b = s_b1 + u * s_b2;
  proof.e_blinding = x * (i_b + x * (o_b + x * s_b));
  tr.append_scalar("t_x", proof.t_x);
  tr.append_scalar("t_x_blinding", proof.t_x_blinding);
  tr.append_scalar("e_blinding", proof.e_blinding);
  Scalar w = tr.challenge_scalar("w");  // :686; Q = w*B is never materialised: its scalar rides on B
  tm.lap("T commits");
  tr.innerproduct_domain_sep(padded_n);  // inner_product_proof.rs:72
  uint8_t wb[32];
  w.to_bytes(wb);
  bpg_ipp* st = nullptr;
  // l(x), r(x), the padding (:661-672) and the G/H factors (:689-697) are evaluated on the device
  rc = bpg_r1cs_dev_ipp_begin(dv.p, g->table, g->g_base(), g->h_base(), g->b_id(), wb, n, n1, padded_n, x.v, u.v, y_pow,
                              y_inv_pow, &st);
  if (rc) return rc;
  tm.lap("l/r eval + ipp begin");
  rc = InnerProductProof::run_rounds(st, tr, &proof.ipp);
  bpg_ipp_free(st);
  if (rc) return rc;
  tm.lap("ipp rounds");
  if (proof.serialized_size() > proof_cap) return BPG_ERR_ARG;
  proof.to_bytes(proof_out);
  *proof_len = proof.serialized_size();
  return BPG_OK;
}

// Prover::prove (prover.rs:412-727).  Blindings: merlin TranscriptRng bound to the transcript, the
// v_blinding witnesses and 32 bytes from the operating system (what `thread_rng()` supplies at :443-444).
extern "C" int bpg_prover_prove(bpg_cs* cs, uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
  uint8_t rb[32];
  if (!os_random(rb)) return BPG_ERR_ARG;
  int rc = prover_prove(cs, true, 0, rb, proof_out, proof_cap, proof_len);
  volatile uint8_t* z = rb;
  for (int i = 0; i < 32; i++) z[i] = 0;
  return rc;
}
// same with the 32 bytes `finalize(&mut rng)` draws supplied by the caller (its own CSPRNG; fixed bytes
// make the proof reproducible, which is how the parity tests compare proof bytes with the oracle)
extern "C" int bpg_prover_prove_with_rng_bytes(bpg_cs* cs, const uint8_t rng_bytes[32], uint8_t* proof_out,
                                               size_t proof_cap, size_t* proof_len) {
  if (!rng_bytes) return BPG_ERR_ARG;
  return prover_prove(cs, true, 0, rng_bytes, proof_out, proof_cap, proof_len);
}
// TEST / BENCH ONLY: every blinding from xoshiro256**(rng_seed) -- reproducible, NOT hiding
extern "C" int bpg_prover_prove_deterministic(bpg_cs* cs, uint64_t rng_seed, uint8_t* proof_out, size_t proof_cap,
                                              size_t* proof_len) {
  return prover_prove(cs, false, rng_seed, nullptr, proof_out, proof_cap, proof_len);
}

// ---------------------------------------------------------------- Verifier::verify (verifier.rs:393-554)
// r_bytes == nullptr: r = challenge_scalar("r") as the mounted fork does (verifier.rs:506).  Otherwise r is
// drawn from a TranscriptRng finalized with r_bytes (upstream's build_rng().finalize(&mut thread_rng())),
// so that a prover cannot predict the scalar that batches the two checks.
// Everything of Verifier::verify (verifier.rs:393-549) up to the final multiscalar multiplication: the ad-hoc points
// and their scalars, the parameters of the device scalar preparation, the scalar of B_blinding; the flattened
// weights are resident in `dv`.  Shared by the single verification and the batch.
struct VerifyPrep {
  std::vector<uint8_t> pts;  // n_adhoc x 32: [A_I1 A_O1 S1 A_I2 A_O2 S2 | V_* | T_* | L_* | R_*]
  std::vector<Scalar> sc;    // their scalars
  bpg_verify_params vp;
  Scalar bb;                 // scalar of B_blinding
  size_t n_adhoc = 0, padded_n = 0;
};
static int verifier_prepare(bpg_cs* cs, const uint8_t* proof_bytes, size_t proof_len, const uint8_t* r_bytes,
                            bool prefetch_points, DevGuard& dv, VerifyPrep& out) {
  if (!cs || cs->is_prover || !proof_bytes) return BPG_ERR_ARG;
  R1CSProof proof;
  int rc = R1CSProof::from_bytes(proof_bytes, proof_len, &proof);
  if (rc) return rc;
  Transcript& tr = *cs->tr;
  const bpg_gens* g = cs->gens;
  StageTimer tm("verify");
  TermsGuard terms_guard{cs->ctx};
  rc = cs->prefetch_terms(false);
  if (rc) return rc;
  // every point of the final check (:516-547) is known now: [A_I1 A_O1 S1 A_I2 A_O2 S2 | V_* | T_* | L_* | R_*];
  // their doubling chains start beside the transcript replay (bpg_adhoc_prefetch)
  const size_t lg_n = proof.ipp.L_vec.size(), m = cs->V.size();
  const size_t n_adhoc = 6 + m + 5 + 2 * lg_n;
  std::vector<uint8_t>& pts = out.pts;
  pts.assign(n_adhoc * 32, 0);
  {
    uint8_t* pp = pts.data();
    auto putp = [&](const uint8_t* p) {
      memcpy(pp, p, 32);
      pp += 32;
    };
    putp(proof.A_I1.data());
    putp(proof.A_O1.data());
    putp(proof.S1.data());
    putp(proof.A_I2.data());
    putp(proof.A_O2.data());
    putp(proof.S2.data());
    for (size_t j = 0; j < m; j++) putp(cs->V[j].data());
    putp(proof.T_1.data());
    putp(proof.T_3.data());
    putp(proof.T_4.data());
    putp(proof.T_5.data());
    putp(proof.T_6.data());
    for (size_t j = 0; j < lg_n; j++) putp(proof.ipp.L_vec[j].data());
    for (size_t j = 0; j < lg_n; j++) putp(proof.ipp.R_vec[j].data());
  }
  if (prefetch_points) {
    rc = bpg_adhoc_prefetch(cs->ctx, pts.data(), n_adhoc);
    if (rc) return rc;
  }
  tr.append_u64("m", cs->V.size());
  size_t n1 = cs->num_vars;
  if (!tr.validate_and_append_point("A_I1", proof.A_I1.data())) return BPG_ERR_VERIFY;
  if (!tr.validate_and_append_point("A_O1", proof.A_O1.data())) return BPG_ERR_VERIFY;
  if (!tr.validate_and_append_point("S1", proof.S1.data())) return BPG_ERR_VERIFY;
  rc = cs->create_randomized_constraints();
  if (rc) return rc;
  size_t n = cs->num_vars, padded_n = next_pow2(n);
  if (g->cap < padded_n) return BPG_ERR_CAPACITY;
  tr.append_point("A_I2", proof.A_I2.data());
  tr.append_point("A_O2", proof.A_O2.data());
  tr.append_point("S2", proof.S2.data());
  Scalar y = tr.challenge_scalar("y"), z = tr.challenge_scalar("z");
  if (!tr.validate_and_append_point("T_1", proof.T_1.data())) return BPG_ERR_VERIFY;
  if (!tr.validate_and_append_point("T_3", proof.T_3.data())) return BPG_ERR_VERIFY;
  if (!tr.validate_and_append_point("T_4", proof.T_4.data())) return BPG_ERR_VERIFY;
  if (!tr.validate_and_append_point("T_5", proof.T_5.data())) return BPG_ERR_VERIFY;
  if (!tr.validate_and_append_point("T_6", proof.T_6.data())) return BPG_ERR_VERIFY;
  Scalar u = tr.challenge_scalar("u"), x = tr.challenge_scalar("x");
  tr.append_scalar("t_x", proof.t_x);
  tr.append_scalar("t_x_blinding", proof.t_x_blinding);
  tr.append_scalar("e_blinding", proof.e_blinding);
  Scalar w = tr.challenge_scalar("w");
  std::vector<Scalar> wV;
  Scalar wc;
  tm.lap("replay");
  rc = bpg_r1cs_dev_new(cs->ctx, std::max<size_t>(n, 1), &dv.p);
  if (rc) return rc;
  tm.lap("state");
  rc = cs->flattened_constraints(dv.p, z, wV, wc);
  if (rc) return rc;
  tm.lap("flatten");
  std::vector<Scalar> u_sq, u_inv_sq;
  Scalar allinv;
  rc = proof.ipp.verification_challenges(padded_n, tr, u_sq, u_inv_sq, allinv);
  if (rc) return BPG_ERR_VERIFY;  // map_err(|_| VerificationError) :463
  tm.lap("verification challenges");
  const Scalar &a = proof.ipp.a, &b = proof.ipp.b;
  Scalar y_inv = y.invert();
  Scalar r;
  if (r_bytes) {
    TranscriptRng vr = tr.build_rng();
    vr.finalize(r_bytes);
    r = vr.scalar();
  } else {
    r = tr.challenge_scalar("r");  // :506
  }
  Scalar xx = x * x, rxx = r * xx, xxx = x * xx;
  // scalars of the ad-hoc points, in the order of `pts` above
  std::vector<Scalar>& sc = out.sc;
  sc.clear();
  sc.reserve(n_adhoc);
  sc.push_back(x);
  sc.push_back(xx);
  sc.push_back(xxx);
  sc.push_back(u * x);
  sc.push_back(u * xx);
  sc.push_back(u * xxx);
  for (size_t j = 0; j < m; j++) sc.push_back(wV[j] * rxx);
  sc.push_back(r * x);
  sc.push_back(rxx * x);
  sc.push_back(rxx * xx);
  sc.push_back(rxx * xxx);
  sc.push_back(rxx * xx * xx);
  for (size_t j = 0; j < lg_n; j++) sc.push_back(u_sq[j]);
  for (size_t j = 0; j < lg_n; j++) sc.push_back(u_inv_sq[j]);
  // y^-i, s, delta = <y^-n o w_R, w_L> (:468-479), g_scalars (:487-491), h_scalars (:493-501) and the
  // scalar of B = w (t_x - a b) + r (x^2 (w_c + delta) - t_x) = c0 + c1 delta (:527-529): device
  bpg_verify_params& vp = out.vp;
  memset(&vp, 0, sizeof vp);
  pow_table(y_inv, vp.y_inv_pow);
  for (size_t j = 0; j < lg_n; j++) memcpy(vp.u_sq[j], u_sq[j].v, 32);
  Scalar c0 = w * (proof.t_x - a * b) + r * (xx * wc - proof.t_x), c1 = rxx;
  memcpy(vp.allinv, allinv.v, 32);
  memcpy(vp.x, x.v, 32);
  memcpy(vp.a, a.v, 32);
  memcpy(vp.b, b.v, 32);
  memcpy(vp.u, u.v, 32);
  memcpy(vp.c0, c0.v, 32);
  memcpy(vp.c1, c1.v, 32);
  vp.lg_n = (uint32_t)lg_n;
  vp.n = (uint32_t)n;
  vp.n1 = (uint32_t)n1;
  vp.N = (uint32_t)padded_n;
  out.bb = -proof.e_blinding - r * proof.t_x_blinding;  // B_blinding
  out.n_adhoc = n_adhoc;
  out.padded_n = padded_n;
  tm.lap("adhoc scalars");
  return BPG_OK;
}
static int verifier_verify(bpg_cs* cs, const uint8_t* proof_bytes, size_t proof_len, const uint8_t* r_bytes) {
  DevGuard dv;
  VerifyPrep P;
  int rc = verifier_prepare(cs, proof_bytes, proof_len, r_bytes, true, dv, P);
  if (rc) return rc;
  const bpg_gens* g = cs->gens;
  StageTimer tm("verify");
  uint8_t bb_scalar[32], mega[32];
  P.bb.to_bytes(bb_scalar);
  std::vector<uint8_t> scb = sc_vec_bytes(P.sc);
  rc = bpg_r1cs_dev_verify_msm(dv.p, g->table, g->g_base(), g->h_base(), g->b_id(), P.pts.data(), scb.data(), P.n_adhoc,
                               bb_scalar, &P.vp, mega);
  tm.lap("g/h scalars + mega msm");
  if (rc == BPG_ERR_DECODE) return BPG_ERR_DECODE;  // a proof point that is not a valid encoding: FormatError
  if (rc) return rc;
  return is_identity_enc(mega) ? BPG_OK : BPG_ERR_VERIFY;  // :549
}
extern "C" int bpg_verifier_verify(bpg_cs* cs, const uint8_t* proof_bytes, size_t proof_len) {
  return verifier_verify(cs, proof_bytes, proof_len, nullptr);
}
// Hardened form: the batching scalar r is bound to the whole transcript AND to 32 bytes the prover cannot
// know (rng_bytes; NULL = read from the operating system).  Accepts exactly the proofs bpg_verifier_verify
// accepts (an honest proof satisfies both checks separately), rejects with overwhelming probability otherwise.
extern "C" int bpg_verifier_verify_with_rng_bytes(bpg_cs* cs, const uint8_t* proof_bytes, size_t proof_len,
                                                  const uint8_t* rng_bytes) {
  uint8_t rb[32];
  if (!rng_bytes) {
    if (!os_random(rb)) return BPG_ERR_ARG;
    rng_bytes = rb;
  }
  return verifier_verify(cs, proof_bytes, proof_len, rng_bytes);
}

// Batch verification (BASELINE.json config 4): the reference verifies proof by proof
// (`Verifier::verify`, verifier.rs:393), so per-proof accept/reject is the parity surface.  Every
// verifier is consumed.  ok[i] = 1 iff proof i verifies; a malformed proof is a reject, not an
// error.  Returns non-zero only for failures of the machinery (bad handle, CUDA, memory).
// One context, one generator set, equal padded sizes: the proofs' final checks are combined.  Every proof k leaves
// its 2 + 2N generator scalars in a slot on the device; sum_k rho_k * (check_k) with rho_0 = 1 and 128-bit rho_k from
// the operating system is ONE multiscalar multiplication over [all proof points | B | B_blinding | G | H] instead of
// one per proof, and it is the identity iff every check is (error 2^-128).  If it is not -- or a point does not
// decode -- each proof is checked from its own slot, so the per-proof answers are the reference's in every case.
static int batch_verify_chunk(bpg_cs* const* verifiers, const uint8_t* const* proofs, const size_t* proof_lens, size_t n,
                              uint8_t* ok) {
  bpg_ctx* ctx = verifiers[0]->ctx;
  const bpg_gens* g = verifiers[0]->gens;
  struct VBGuard {
    bpg_vbatch* p = nullptr;
    ~VBGuard() { bpg_vbatch_free(p); }
  } vb;
  std::vector<VerifyPrep> preps(n);
  std::vector<size_t> slot_of;  // proofs that made it into the batch, in slot order
  size_t N0 = 0;
  for (size_t i = 0; i < n; i++) {
    ok[i] = 0;
    DevGuard dv;
    int rc = verifier_prepare(verifiers[i], proofs[i], proof_lens[i], nullptr, false, dv, preps[i]);
    if (rc == BPG_ERR_VERIFY || rc == BPG_ERR_DECODE || rc == BPG_ERR_CAPACITY) continue;  // rejected before the check
    if (rc) return rc;
    if (!vb.p) {
      N0 = preps[i].padded_n;
      rc = bpg_vbatch_new(ctx, N0, n, &vb.p);
      if (rc) return rc;
    }
    uint8_t bb[32];
    preps[i].bb.to_bytes(bb);
    if (preps[i].padded_n != N0) {  // another size: on its own
      uint8_t mega[32];
      std::vector<uint8_t> scb = sc_vec_bytes(preps[i].sc);
      rc = bpg_r1cs_dev_verify_msm(dv.p, g->table, g->g_base(), g->h_base(), g->b_id(), preps[i].pts.data(), scb.data(),
                                   preps[i].n_adhoc, bb, &preps[i].vp, mega);
      if (rc == BPG_OK) ok[i] = is_identity_enc(mega) ? 1 : 0;
      else if (rc != BPG_ERR_DECODE) return rc;
      continue;
    }
    rc = bpg_vbatch_put(vb.p, slot_of.size(), dv.p, bb, &preps[i].vp);
    if (rc) return rc;
    slot_of.push_back(i);
  }
  if (slot_of.empty()) return BPG_OK;
  StageTimer tmb("batch");
  auto check = [&](const std::vector<uint32_t>& slots, const std::vector<Scalar>& rho, uint8_t mega[32]) {
    std::vector<uint8_t> pts, scb, rb(slots.size() * 32);
    for (size_t j = 0; j < slots.size(); j++) {
      const VerifyPrep& P = preps[slot_of[slots[j]]];
      pts.insert(pts.end(), P.pts.begin(), P.pts.end());
      for (const Scalar& x : P.sc) {
        uint8_t b32[32];
        (x * rho[j]).to_bytes(b32);
        scb.insert(scb.end(), b32, b32 + 32);
      }
      rho[j].to_bytes(rb.data() + 32 * j);
    }
    return bpg_vbatch_check(vb.p, g->table, g->g_base(), g->h_base(), g->b_id(), slots.data(), slots.size(), rb.data(),
                            pts.data(), scb.data(), pts.size() / 32, mega);
  };
  // does the random combination of these slots pass?  (a point that does not decode counts as a failure)
  int err = BPG_OK;
  auto passes = [&](const std::vector<uint32_t>& slots) {
    std::vector<Scalar> rho(slots.size());
    for (size_t j = 0; j < slots.size(); j++) {
      if (j == 0) {
        rho[j] = Scalar::one();
      } else {
        uint8_t rb[32], lo[32] = {0};
        if (!os_random(rb)) {
          err = BPG_ERR_ARG;
          return false;
        }
        memcpy(lo, rb, 16);  // 128 bits
        rho[j] = Scalar::from_bytes_mod_order(lo);
      }
    }
    uint8_t mega[32];
    int rc = check(slots, rho, mega);
    tmb.lap("combined check");
    if (rc != BPG_OK && rc != BPG_ERR_DECODE) err = rc;
    return rc == BPG_OK && is_identity_enc(mega);
  };
  // `failed`: the combination of exactly these slots is known not to pass.  Halves are tested and the failing ones
  // split again: one bad proof among n costs about 2 lg n checks instead of n.
  std::function<void(const std::vector<uint32_t>&, bool)> solve = [&](const std::vector<uint32_t>& slots, bool failed) {
    if (err) return;
    if (!failed && passes(slots)) {
      for (uint32_t j : slots) ok[slot_of[j]] = 1;
      return;
    }
    if (err || slots.size() == 1) return;  // a single slot that does not pass: rejected (ok stays 0)
    const size_t half = slots.size() / 2;
    std::vector<uint32_t> lo(slots.begin(), slots.begin() + half), hi(slots.begin() + half, slots.end());
    const bool lo_ok = passes(lo);
    if (err) return;
    if (lo_ok) {
      for (uint32_t j : lo) ok[slot_of[j]] = 1;
      solve(hi, true);  // the failure is in the other half
    } else {
      solve(lo, true);
      solve(hi, false);
    }
  };
  std::vector<uint32_t> all(slot_of.size());
  for (size_t j = 0; j < all.size(); j++) all[j] = (uint32_t)j;
  solve(all, false);
  return err;
}
extern "C" int bpg_batch_verify(bpg_cs* const* verifiers, const uint8_t* const* proofs, const size_t* proof_lens,
                                size_t n, uint8_t* ok) {
  if (n && (!verifiers || !proofs || !proof_lens || !ok)) return BPG_ERR_ARG;
  for (size_t i = 0; i < n; i++)
    if (!verifiers[i] || verifiers[i]->is_prover || !proofs[i]) return BPG_ERR_ARG;
  // chunks of proofs that share a context and a generator set; the slots of a chunk stay within ~512 MB
  size_t i = 0;
  while (i < n) {
    const size_t cap = verifiers[i]->gens ? verifiers[i]->gens->cap : 0;
    const size_t per = std::max<size_t>(1, std::min<size_t>(256, ((size_t)512 << 20) / ((2 * std::max<size_t>(cap, 1) + 2) * 32)));
    size_t j = i + 1;
    while (j < n && j - i < per && verifiers[j]->ctx == verifiers[i]->ctx && verifiers[j]->gens == verifiers[i]->gens) j++;
    int rc = batch_verify_chunk(verifiers + i, proofs + i, proof_lens + i, j - i, ok + i);
    if (rc) return rc;
    i = j;
  }
  return BPG_OK;
}

// ---------------------------------------------------------------- Stark-curve conventions of the mounted fork
// (the part of its transcript / generator layer that IS in /root/reference: src/util.rs:252-289,
// src/generators.rs:80-125; `pad_label` and the HashChainTranscript construction live in the un-vendored merlin fork)
extern "C" void bpg_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) { keccak256(data, len, out); }

// hash_to_scalar (src/util.rs:252-267): (low || keccak256(low)) as a 512-bit little-endian integer mod the Stark
// group order; out = the scalar, 32 bytes little-endian
extern "C" void bpg_stark_hash_to_scalar(const uint8_t low[32], uint8_t out[32]) {
  static const uint64_t N[4] = {0x1e66a241adc64d2fULL, 0xb781126dcae7b232ULL, 0xffffffffffffffffULL, 0x0800000000000010ULL};
  uint8_t wide[64];
  memcpy(wide, low, 32);
  keccak256(low, 32, wide + 32);
  uint64_t r[4] = {0, 0, 0, 0};
  for (int bit = 511; bit >= 0; bit--) {  // r = 2 r + bit, reduced: r < N < 2^252, so 2 r + 1 < 2^253 never overflows
    uint64_t in = (wide[bit >> 3] >> (bit & 7)) & 1;
    for (int i = 3; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 63);
    r[0] = (r[0] << 1) | in;
    bool ge = true;
    for (int i = 3; i >= 0; i--)
      if (r[i] != N[i]) {
        ge = r[i] > N[i];
        break;
      }
    if (ge) {
      unsigned __int128 borrow = 0;
      for (int i = 0; i < 4; i++) {
        unsigned __int128 t = (unsigned __int128)r[i] - N[i] - (uint64_t)borrow;
        r[i] = (uint64_t)t;
        borrow = (t >> 64) & 1;
      }
    }
  }
  for (int i = 0; i < 4; i++)
    for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(r[i] >> (8 * b));
}

extern "C" int bpg_stark_wide_mul_generator(bpg_ctx* ctx, const uint8_t* wide64, size_t n, uint8_t* out_xy);
// GeneratorsChain (src/generators.rs:80-125) from its initial state state0 = keccak256(pad_label("GeneratorsChain" ||
// label)) -- the caller forms it (pad_label is the merlin fork's).  Points [skip, skip + n): per point
// state <- keccak256(state), scalar = hash_to_scalar(state), point = scalar * G; out = n x 64 bytes affine x || y.
extern "C" int bpg_stark_gens_chain(bpg_ctx* ctx, const uint8_t state0[32], size_t skip, size_t n, uint8_t* out_xy) {
  if (!ctx || !state0 || (n && !out_xy)) return BPG_ERR_ARG;
  uint8_t st[32], nx[32];
  memcpy(st, state0, 32);
  for (size_t i = 0; i < skip; i++) {  // fast_forward (:92-100)
    keccak256(st, 32, nx);
    memcpy(st, nx, 32);
  }
  std::vector<uint8_t> wide(n * 64);
  for (size_t i = 0; i < n; i++) {
    keccak256(st, 32, nx);
    memcpy(st, nx, 32);
    memcpy(wide.data() + 64 * i, st, 32);
    keccak256(st, 32, wide.data() + 64 * i + 32);
  }
  return bpg_stark_wide_mul_generator(ctx, wide.data(), n, out_xy);
}
