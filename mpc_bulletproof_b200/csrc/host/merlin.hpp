// Merlin v1.0 transcript (STROBE-128 over Keccak-f[1600]) and the
// `TranscriptProtocol` of the reference (src/transcript.rs:25-121) with the
// ristretto255 payload conventions of SURVEY.md A.6.  Host-side, sequential,
// <= 64 bytes per step: it is the per-round synchronisation point of the
// proof, not part of the data-parallel path.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "sc_host.hpp"

namespace bpg_host {

inline void keccak_f1600(uint64_t st[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
      0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
      0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
      0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
      0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  auto rol = [](uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; };
  for (int r = 0; r < 24; r++) {
    uint64_t c[5], d[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) st[i] ^= d[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(st[x + 5 * y], ROT[x + 5 * y]);
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) st[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    st[0] ^= RC[r];
  }
}

// Keccak sponge for the generator chains (reference src/generators.rs:80-125 in its ristretto255
// form: SHAKE256("GeneratorsChain" || label) as an XOF, 64 bytes per point) and for
// PedersenGens::default's B_blinding = hash-to-group(SHA3-512(B)) (src/generators.rs:61-71).
class KeccakSponge {
 public:
  KeccakSponge(size_t rate, uint8_t suffix) : rate_(rate), suffix_(suffix), pos_(0), squeezing_(false) { memset(st_, 0, sizeof st_); }
  void absorb(const uint8_t* d, size_t n) {
    uint8_t* s = reinterpret_cast<uint8_t*>(st_);
    for (size_t i = 0; i < n; i++) {
      s[pos_++] ^= d[i];
      if (pos_ == rate_) { keccak_f1600(st_); pos_ = 0; }
    }
  }
  void squeeze(uint8_t* out, size_t n) {
    uint8_t* s = reinterpret_cast<uint8_t*>(st_);
    if (!squeezing_) {
      s[pos_] ^= suffix_;
      s[rate_ - 1] ^= 0x80;
      keccak_f1600(st_);
      pos_ = 0;
      squeezing_ = true;
    }
    for (size_t i = 0; i < n; i++) {
      if (pos_ == rate_) { keccak_f1600(st_); pos_ = 0; }
      out[i] = s[pos_++];
    }
  }

 private:
  uint64_t st_[25];
  size_t rate_;
  uint8_t suffix_;
  size_t pos_;
  bool squeezing_;
};
inline KeccakSponge shake256() { return KeccakSponge(136, 0x1f); }
// legacy Keccak-256 (pad 0x01): `merlin::keccak256` of the fork (reference src/generators.rs:84-123, src/util.rs:252-267)
inline void keccak256(const uint8_t* d, size_t n, uint8_t out[32]) {
  KeccakSponge k(136, 0x01);
  k.absorb(d, n);
  k.squeeze(out, 32);
}
inline void sha3_512(const uint8_t* d, size_t n, uint8_t out[64]) {
  KeccakSponge k(72, 0x06);
  k.absorb(d, n);
  k.squeeze(out, 64);
}

class Strobe128 {
 public:
  explicit Strobe128(const char* protocol_label) {
    memset(st_, 0, sizeof st_);
    uint8_t* s = bytes();
    const uint8_t init[6] = {1, R + 2, 1, 0, 1, 96};
    memcpy(s, init, 6);
    memcpy(s + 6, "STROBEv1.0.2", 12);
    keccak_f1600(st_);
    pos_ = pos_begin_ = cur_flags_ = 0;
    meta_ad((const uint8_t*)protocol_label, strlen(protocol_label), false);
  }
  void meta_ad(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_M | FLAG_A, more); absorb(d, n); }
  void ad(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_A, more); absorb(d, n); }
  void prf(uint8_t* out, size_t n, bool more) { begin_op(FLAG_I | FLAG_A | FLAG_C, more); squeeze(out, n); }
  void key(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_A | FLAG_C, more); overwrite(d, n); }

 private:
  static constexpr uint8_t R = 166;
  static constexpr uint8_t FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32;
  uint64_t st_[25];
  uint8_t pos_, pos_begin_, cur_flags_;
  uint8_t* bytes() { return reinterpret_cast<uint8_t*>(st_); }
  void run_f() {
    uint8_t* s = bytes();
    s[pos_] ^= pos_begin_;
    s[pos_ + 1] ^= 0x04;
    s[R + 1] ^= 0x80;
    keccak_f1600(st_);
    pos_ = 0;
    pos_begin_ = 0;
  }
  void absorb(const uint8_t* d, size_t n) {
    uint8_t* s = bytes();
    for (size_t i = 0; i < n; i++) {
      s[pos_] ^= d[i];
      if (++pos_ == R) run_f();
    }
  }
  void overwrite(const uint8_t* d, size_t n) {
    uint8_t* s = bytes();
    for (size_t i = 0; i < n; i++) {
      s[pos_] = d[i];
      if (++pos_ == R) run_f();
    }
  }
  void squeeze(uint8_t* out, size_t n) {
    uint8_t* s = bytes();
    for (size_t i = 0; i < n; i++) {
      out[i] = s[pos_];
      s[pos_] = 0;
      if (++pos_ == R) run_f();
    }
  }
  void begin_op(uint8_t flags, bool more) {
    if (more) return;
    uint8_t old_begin = pos_begin_;
    pos_begin_ = pos_ + 1;
    cur_flags_ = flags;
    uint8_t hdr[2] = {old_begin, flags};
    absorb(hdr, 2);
    if ((flags & (FLAG_C | FLAG_K)) && pos_ != 0) run_f();
  }
};

// merlin::TranscriptRng (merlin transcript.rs: TranscriptRngBuilder / TranscriptRng): a clone of the
// transcript's STROBE state rekeyed with witness bytes and finalized with 32 bytes of external
// randomness; every fill is a length-framed PRF output.  The prover's blinding source (reference
// src/r1cs/prover.rs:435-445) and, in the hardened form, the verifier's batching scalar.
class TranscriptRng {
 public:
  explicit TranscriptRng(const Strobe128& s) : strobe_(s) {}
  void rekey_with_witness_bytes(const char* label, const uint8_t* w, size_t n) {
    uint32_t len = (uint32_t)n;
    uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
    strobe_.meta_ad((const uint8_t*)label, strlen(label), false);
    strobe_.meta_ad(le, 4, true);
    strobe_.key(w, n, false);
  }
  void finalize(const uint8_t random_bytes[32]) {
    strobe_.meta_ad((const uint8_t*)"rng", 3, false);
    strobe_.key(random_bytes, 32, false);
  }
  void fill_bytes(uint8_t* out, size_t n) {
    uint32_t len = (uint32_t)n;
    uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
    strobe_.meta_ad(le, 4, false);
    strobe_.prf(out, n, false);
  }
  Scalar scalar() {  // Scalar::random(&mut rng): 64 bytes, wide reduction
    uint8_t b[64];
    fill_bytes(b, 64);
    return Scalar::from_wide(b);
  }

 private:
  Strobe128 strobe_;
};

// merlin::Transcript + the reference's TranscriptProtocol
class Transcript {
 public:
  explicit Transcript(const uint8_t* label, size_t n) : strobe_("Merlin v1.0") { append_message("dom-sep", label, n); }
  explicit Transcript(const char* label) : Transcript((const uint8_t*)label, strlen(label)) {}

  void append_message(const char* label, const uint8_t* msg, size_t n) {
    uint32_t len = (uint32_t)n;
    uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
    strobe_.meta_ad((const uint8_t*)label, strlen(label), false);
    strobe_.meta_ad(le, 4, true);
    strobe_.ad(msg, n, false);
  }
  void append_u64(const char* label, uint64_t x) {
    uint8_t le[8];
    for (int i = 0; i < 8; i++) le[i] = (uint8_t)(x >> (8 * i));
    append_message(label, le, 8);
  }
  void challenge_bytes(const char* label, uint8_t* out, size_t n) {
    uint32_t len = (uint32_t)n;
    uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
    strobe_.meta_ad((const uint8_t*)label, strlen(label), false);
    strobe_.meta_ad(le, 4, true);
    strobe_.prf(out, n, false);
  }
  // ---- TranscriptProtocol (reference src/transcript.rs:63-121) ----
  void innerproduct_domain_sep(uint64_t n) {
    append_message("dom-sep", (const uint8_t*)"ipp v1", 6);
    append_u64("n", n);
  }
  void r1cs_domain_sep() { append_message("dom-sep", (const uint8_t*)"r1cs v1", 7); }
  void r1cs_1phase_domain_sep() { append_message("dom-sep", (const uint8_t*)"r1cs-1phase", 11); }
  void r1cs_2phase_domain_sep() { append_message("dom-sep", (const uint8_t*)"r1cs-2phase", 11); }
  void append_scalar(const char* label, const Scalar& s) {
    uint8_t b[32];
    s.to_bytes(b);
    append_message(label, b, 32);
  }
  void append_point(const char* label, const uint8_t p[32]) { append_message(label, p, 32); }
  // false (and nothing appended) if p is the identity (all-zero encoding)
  bool validate_and_append_point(const char* label, const uint8_t p[32]) {
    uint8_t acc = 0;
    for (int i = 0; i < 32; i++) acc |= p[i];
    if (!acc) return false;
    append_message(label, p, 32);
    return true;
  }
  Scalar challenge_scalar(const char* label) {
    uint8_t b[64];
    challenge_bytes(label, b, 64);
    return Scalar::from_wide(b);
  }
  TranscriptRng build_rng() const { return TranscriptRng(strobe_); }

 private:
  Strobe128 strobe_;
};

}  // namespace bpg_host
