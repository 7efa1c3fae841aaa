// Host-side scalars mod l (the ristretto255 group order), 4x64 Montgomery limbs.
// Used by the host mirror of the reference's protocol layer for the O(lg n)
// bookkeeping that stays on the CPU exactly as in the reference: challenges,
// their inverses, polynomial evaluation at a challenge, proof (de)serialisation
// (reference src/inner_product_proof.rs:119-123, src/r1cs/prover.rs:584-697).
#pragma once
#include <cstdint>
#include <cstring>

namespace bpg_host {

typedef unsigned __int128 u128;

struct Scalar {
  uint64_t v[4];  // Montgomery form, R = 2^256, fully reduced

  static const uint64_t* L() {
    static const uint64_t l[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0x0ULL, 0x1000000000000000ULL};
    return l;
  }
  static constexpr uint64_t NINV = 0xd2b51da312547e1bULL;  // -l^{-1} mod 2^64
  static const uint64_t* RR() {
    static const uint64_t rr[4] = {0xa40611e3449c0f01ULL, 0xd00e1ba768859347ULL, 0xceec73d217f5be65ULL,
                                   0x0399411b7c309a3dULL};
    return rr;
  }

  static bool geq_l(const uint64_t x[4]) {
    const uint64_t* l = L();
    for (int i = 3; i >= 0; i--) {
      if (x[i] > l[i]) return true;
      if (x[i] < l[i]) return false;
    }
    return true;
  }
  static void sub_l(uint64_t x[4]) {
    const uint64_t* l = L();
    u128 bw = 0;
    for (int i = 0; i < 4; i++) {
      u128 t = (u128)x[i] - l[i] - bw;
      x[i] = (uint64_t)t;
      bw = (t >> 64) & 1;
    }
  }
  static Scalar montmul(const uint64_t a[4], const uint64_t b[4]) {
    const uint64_t* l = L();
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      u128 c = 0;
      for (int j = 0; j < 4; j++) {
        u128 s = (u128)a[j] * b[i] + t[j] + c;
        t[j] = (uint64_t)s;
        c = s >> 64;
      }
      u128 s = (u128)t[4] + c;
      t[4] = (uint64_t)s;
      t[5] = (uint64_t)(s >> 64);
      uint64_t m = t[0] * NINV;
      c = ((u128)m * l[0] + t[0]) >> 64;
      for (int j = 1; j < 4; j++) {
        u128 s2 = (u128)m * l[j] + t[j] + c;
        t[j - 1] = (uint64_t)s2;
        c = s2 >> 64;
      }
      s = (u128)t[4] + c;
      t[3] = (uint64_t)s;
      t[4] = t[5] + (uint64_t)(s >> 64);
      t[5] = 0;
    }
    Scalar r;
    memcpy(r.v, t, 32);
    if (t[4] || geq_l(r.v)) sub_l(r.v);
    return r;
  }

  static Scalar zero() { Scalar s; memset(s.v, 0, 32); return s; }
  static Scalar from_u64(uint64_t x) {
    uint64_t a[4] = {x, 0, 0, 0};
    return montmul(a, RR());
  }
  static Scalar one() { return from_u64(1); }
  // canonical little-endian bytes; returns false if >= l
  static bool from_bytes(const uint8_t b[32], Scalar* out) {
    uint64_t a[4];
    memcpy(a, b, 32);
    if (geq_l(a)) return false;
    *out = montmul(a, RR());
    return true;
  }
  // canonical encoding? (x < l), without converting
  static bool is_canonical(const uint8_t b[32]) {
    uint64_t a[4];
    memcpy(a, b, 32);
    return !geq_l(a);
  }
  // any 32 bytes, reduced mod l
  static Scalar from_bytes_mod_order(const uint8_t b[32]) {
    uint64_t a[4];
    memcpy(a, b, 32);
    // a < 2^256 < 16 l: montmul reduces (a * R^2 / R = aR mod l) for any a < 2^256
    return montmul(a, RR());
  }
  // 64 uniform bytes -> mod l (challenge scalars)
  static Scalar from_wide(const uint8_t b[64]) {
    uint64_t lo[4], hi[4];
    memcpy(lo, b, 32);
    memcpy(hi, b + 32, 32);
    Scalar l_ = montmul(lo, RR());          // lo * R
    Scalar h_ = montmul(hi, RR());          // hi * R
    Scalar r2 = montmul(RR(), RR());        // R^2 * R  (Montgomery form of R^2)... see below
    // hi * 2^256 = hi * R: in Montgomery form (hi R) * (R * R)/R = hi R R -> montmul(h_, RRm) where RRm = R^2 (= Mont form of R)
    (void)r2;
    Scalar rm;  // Montgomery form of R mod l is R^2 mod l
    memcpy(rm.v, RR(), 32);
    Scalar hR = montmul(h_.v, rm.v);
    return l_ + hR;
  }
  void to_bytes(uint8_t out[32]) const {
    uint64_t one_[4] = {1, 0, 0, 0};
    Scalar n = montmul(v, one_);
    memcpy(out, n.v, 32);
  }
  Scalar operator*(const Scalar& o) const { return montmul(v, o.v); }
  Scalar operator+(const Scalar& o) const {
    Scalar r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      u128 s = (u128)v[i] + o.v[i] + c;
      r.v[i] = (uint64_t)s;
      c = s >> 64;
    }
    if (geq_l(r.v)) sub_l(r.v);
    return r;
  }
  Scalar operator-(const Scalar& o) const {
    Scalar r;
    u128 bw = 0;
    for (int i = 0; i < 4; i++) {
      u128 t = (u128)v[i] - o.v[i] - bw;
      r.v[i] = (uint64_t)t;
      bw = (t >> 64) & 1;
    }
    if (bw) {
      const uint64_t* l = L();
      u128 c = 0;
      for (int i = 0; i < 4; i++) {
        u128 s = (u128)r.v[i] + l[i] + c;
        r.v[i] = (uint64_t)s;
        c = s >> 64;
      }
    }
    return r;
  }
  Scalar operator-() const { return zero() - *this; }
  Scalar& operator+=(const Scalar& o) { *this = *this + o; return *this; }
  Scalar& operator-=(const Scalar& o) { *this = *this - o; return *this; }
  Scalar& operator*=(const Scalar& o) { *this = *this * o; return *this; }
  bool operator==(const Scalar& o) const { return memcmp(v, o.v, 32) == 0; }
  bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
  // x^(l-2): the reference's Scalar::invert as an exponentiation (kept as the cross-check of invert())
  Scalar invert_fermat() const {
    // l - 2 = 2^252 + 27742317777372353535851937790883648493 - 2
    static const uint64_t e[4] = {0x5812631a5cf5d3ebULL, 0x14def9dea2f79cd6ULL, 0x0ULL, 0x1000000000000000ULL};
    Scalar r = one(), base = *this;
    for (int i = 0; i < 256; i++) {
      if ((e[i >> 6] >> (i & 63)) & 1) r = r * base;
      base = base * base;
    }
    return r;
  }
  // The same value by Kaliski's binary "almost Montgomery inverse": about 1.4 x 253 steps of shifts and
  // subtractions on four limbs (a^-1 2^k mod l), then one multiplication by 2^(768-k) to land in Montgomery form.
  // The prover inverts one challenge per inner-product round with the GPU waiting on it: the exponentiation was
  // 10-14 us of every round, this is about 3.  (Variable time in its input: challenges are public.)
  Scalar invert() const {
    if (is_zero()) return zero();
    const uint64_t* l = L();
    uint64_t u[4] = {l[0], l[1], l[2], l[3]}, w[4] = {v[0], v[1], v[2], v[3]}, r[4] = {0, 0, 0, 0}, q[4] = {1, 0, 0, 0};
    auto shr = [](uint64_t x[4], int n) {  // 1 <= n <= 63
      x[0] = (x[0] >> n) | (x[1] << (64 - n));
      x[1] = (x[1] >> n) | (x[2] << (64 - n));
      x[2] = (x[2] >> n) | (x[3] << (64 - n));
      x[3] >>= n;
    };
    auto shl = [](uint64_t x[4], int n) {
      x[3] = (x[3] << n) | (x[2] >> (64 - n));
      x[2] = (x[2] << n) | (x[1] >> (64 - n));
      x[1] = (x[1] << n) | (x[0] >> (64 - n));
      x[0] <<= n;
    };
    auto sub = [](uint64_t x[4], const uint64_t y[4]) {  // x -= y, x >= y
      u128 bw = 0;
      for (int i = 0; i < 4; i++) {
        u128 t = (u128)x[i] - y[i] - bw;
        x[i] = (uint64_t)t;
        bw = (t >> 64) & 1;
      }
    };
    auto add = [](uint64_t x[4], const uint64_t y[4]) {  // x += y (both < 2^255)
      u128 c = 0;
      for (int i = 0; i < 4; i++) {
        u128 t = (u128)x[i] + y[i] + c;
        x[i] = (uint64_t)t;
        c = t >> 64;
      }
    };
    auto gt = [](const uint64_t x[4], const uint64_t y[4]) {
      for (int i = 3; i >= 0; i--) {
        if (x[i] != y[i]) return x[i] > y[i];
      }
      return false;
    };
    // x is nonzero and even: strip its trailing zeros from x, apply the same shift (upwards) to y
    auto strip = [&](uint64_t x[4], uint64_t y[4], int& k) {
      while (!(x[0] & 1)) {
        int n = x[0] ? __builtin_ctzll(x[0]) : 63;
        shr(x, n);
        shl(y, n);
        k += n;
      }
    };
    int k = 0;
    // Kaliski's steps with every run of halvings taken at once: u, w odd at the top of the loop
    if (!(w[0] & 1)) strip(w, r, k);
    while (true) {
      if (gt(u, w)) {
        sub(u, w);
        add(r, q);
        strip(u, q, k);
      } else {
        sub(w, u);
        add(q, r);
        if (!(w[0] | w[1] | w[2] | w[3])) {  // w = u = gcd = 1 met: the last step of the unrolled loop is r <<= 1
          shl(r, 1);
          k++;
          break;
        }
        strip(w, r, k);
      }
    }
    if (geq_l(r)) sub_l(r);  // r < 2 l
    uint64_t res[4] = {l[0], l[1], l[2], l[3]};
    sub(res, r);             // l - r = x^-1 2^k mod l, 253 <= k <= 506
    return montmul(res, pow2_mod_l(768 - k));
  }
  // 2^e mod l as a plain integer, e <= 515
  static const uint64_t* pow2_mod_l(int e) {
    struct Table {
      uint64_t t[516][4];
      Table() {
        uint64_t x[4] = {1, 0, 0, 0};
        for (int i = 0; i < 516; i++) {
          memcpy(t[i], x, 32);
          // x = 2x mod l  (x < l < 2^253: no overflow)
          x[3] = (x[3] << 1) | (x[2] >> 63);
          x[2] = (x[2] << 1) | (x[1] >> 63);
          x[1] = (x[1] << 1) | (x[0] >> 63);
          x[0] <<= 1;
          if (geq_l(x)) sub_l(x);
        }
      }
    };
    static const Table T;
    return T.t[e];
  }
};

}  // namespace bpg_host
