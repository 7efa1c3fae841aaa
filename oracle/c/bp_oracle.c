/* ORACLE / CPU BASELINE — test and measurement infrastructure only.
 *
 * Plain-C restatement of the CPU algorithm behind the reference's group calls
 * for the ristretto255 instantiation: 64-bit-limb field
 * arithmetic (radix 2^51, the layout of curve25519-dalek's `u64` backend),
 * extended twisted-Edwards points, RFC 9496 encode/decode, and the vartime
 * multiscalar multiplication `StarkPoint::msm_iter` resolves to at the call
 * sites reference src/inner_product_proof.rs:90-114, src/r1cs/verifier.rs:516-547:
 * Straus with width-5 NAF below 190 terms, signed radix-2^w Pippenger above with
 * w = 6 (<500 terms), 7 (<800), 8 otherwise — the schedule of the dalek
 * `vartime_multiscalar_mul` the north-star names as the CPU path.  The group
 * dependency itself (mpc-stark ^0.2 / curve25519-dalek) is not vendored in
 * /root/reference; this file restates the published algorithms and is pinned
 * against oracle/group.py (RFC 9496 vectors) by tests/test_oracle_c.py.
 *
 * Threads: `threads <= 1` is the serial algorithm as the reference runs it;
 * `threads > 1` splits the Pippenger digit columns over OpenMP threads (the way
 * ark-ec's rayon feature parallelises windows) — used by bench.py's reference arm.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  Nothing in the product links it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* ------------------------------------------------------------------ field */
typedef struct { u64 v[5]; } fe;
#define MASK51 ((1ULL << 51) - 1)

static const fe FE_ZERO = {{0, 0, 0, 0, 0}};
static const fe FE_ONE = {{1, 0, 0, 0, 0}};

static void fe_frombytes(fe* h, const uint8_t s[32]) {
  u64 w[4];
  memcpy(w, s, 32);
  h->v[0] = w[0] & MASK51;
  h->v[1] = ((w[0] >> 51) | (w[1] << 13)) & MASK51;
  h->v[2] = ((w[1] >> 38) | (w[2] << 26)) & MASK51;
  h->v[3] = ((w[2] >> 25) | (w[3] << 39)) & MASK51;
  h->v[4] = (w[3] >> 12) & MASK51; /* drops bit 255 */
}
static void fe_carry(fe* h) {
  u64 c;
  c = h->v[0] >> 51; h->v[0] &= MASK51; h->v[1] += c;
  c = h->v[1] >> 51; h->v[1] &= MASK51; h->v[2] += c;
  c = h->v[2] >> 51; h->v[2] &= MASK51; h->v[3] += c;
  c = h->v[3] >> 51; h->v[3] &= MASK51; h->v[4] += c;
  c = h->v[4] >> 51; h->v[4] &= MASK51; h->v[0] += c * 19;
}
static void fe_tobytes(uint8_t s[32], const fe* f) {
  fe h = *f;
  fe_carry(&h);
  fe_carry(&h);
  /* h < 2^255 + small; compute h mod p: q = (h + 19) >> 255 */
  u64 q = (h.v[0] + 19) >> 51;
  q = (h.v[1] + q) >> 51;
  q = (h.v[2] + q) >> 51;
  q = (h.v[3] + q) >> 51;
  q = (h.v[4] + q) >> 51;
  h.v[0] += 19 * q;
  u64 c;
  c = h.v[0] >> 51; h.v[0] &= MASK51; h.v[1] += c;
  c = h.v[1] >> 51; h.v[1] &= MASK51; h.v[2] += c;
  c = h.v[2] >> 51; h.v[2] &= MASK51; h.v[3] += c;
  c = h.v[3] >> 51; h.v[3] &= MASK51; h.v[4] += c;
  h.v[4] &= MASK51;
  u64 w[4];
  w[0] = h.v[0] | (h.v[1] << 51);
  w[1] = (h.v[1] >> 13) | (h.v[2] << 38);
  w[2] = (h.v[2] >> 26) | (h.v[3] << 25);
  w[3] = (h.v[3] >> 39) | (h.v[4] << 12);
  memcpy(s, w, 32);
}
static inline void fe_add(fe* h, const fe* f, const fe* g) {
  for (int i = 0; i < 5; i++) h->v[i] = f->v[i] + g->v[i];
}
/* h = f - g; adds 4p so limbs stay non-negative for inputs with limbs < 2^53 */
static inline void fe_sub(fe* h, const fe* f, const fe* g) {
  h->v[0] = f->v[0] + 0x1FFFFFFFFFFFB4ULL - g->v[0];
  h->v[1] = f->v[1] + 0x1FFFFFFFFFFFFCULL - g->v[1];
  h->v[2] = f->v[2] + 0x1FFFFFFFFFFFFCULL - g->v[2];
  h->v[3] = f->v[3] + 0x1FFFFFFFFFFFFCULL - g->v[3];
  h->v[4] = f->v[4] + 0x1FFFFFFFFFFFFCULL - g->v[4];
  fe_carry(h);
}
static inline void fe_neg(fe* h, const fe* f) { fe_sub(h, &FE_ZERO, f); }
static inline void fe_mul(fe* h, const fe* f, const fe* g) {
  u64 f0 = f->v[0], f1 = f->v[1], f2 = f->v[2], f3 = f->v[3], f4 = f->v[4];
  u64 g0 = g->v[0], g1 = g->v[1], g2 = g->v[2], g3 = g->v[3], g4 = g->v[4];
  u64 g1_19 = g1 * 19, g2_19 = g2 * 19, g3_19 = g3 * 19, g4_19 = g4 * 19;
  u128 r0 = (u128)f0 * g0 + (u128)f1 * g4_19 + (u128)f2 * g3_19 + (u128)f3 * g2_19 + (u128)f4 * g1_19;
  u128 r1 = (u128)f0 * g1 + (u128)f1 * g0 + (u128)f2 * g4_19 + (u128)f3 * g3_19 + (u128)f4 * g2_19;
  u128 r2 = (u128)f0 * g2 + (u128)f1 * g1 + (u128)f2 * g0 + (u128)f3 * g4_19 + (u128)f4 * g3_19;
  u128 r3 = (u128)f0 * g3 + (u128)f1 * g2 + (u128)f2 * g1 + (u128)f3 * g0 + (u128)f4 * g4_19;
  u128 r4 = (u128)f0 * g4 + (u128)f1 * g3 + (u128)f2 * g2 + (u128)f3 * g1 + (u128)f4 * g0;
  u64 c;
  r1 += (u64)(r0 >> 51); u64 h0 = (u64)r0 & MASK51;
  r2 += (u64)(r1 >> 51); u64 h1 = (u64)r1 & MASK51;
  r3 += (u64)(r2 >> 51); u64 h2 = (u64)r2 & MASK51;
  r4 += (u64)(r3 >> 51); u64 h3 = (u64)r3 & MASK51;
  c = (u64)(r4 >> 51);   u64 h4 = (u64)r4 & MASK51;
  h0 += c * 19;
  c = h0 >> 51; h0 &= MASK51; h1 += c;
  h->v[0] = h0; h->v[1] = h1; h->v[2] = h2; h->v[3] = h3; h->v[4] = h4;
}
static inline void fe_sq(fe* h, const fe* f) { fe_mul(h, f, f); }
static void fe_sqn(fe* h, const fe* f, int n) {
  fe t = *f;
  for (int i = 0; i < n; i++) fe_sq(&t, &t);
  *h = t;
}
/* z^(2^252 - 3) */
static void fe_pow22523(fe* out, const fe* z) {
  fe t0, t1, t2;
  fe_sq(&t0, z);
  fe_sqn(&t1, &t0, 2);
  fe_mul(&t1, z, &t1);
  fe_mul(&t0, &t0, &t1);
  fe_sq(&t0, &t0);
  fe_mul(&t0, &t1, &t0);
  fe_sqn(&t1, &t0, 5);
  fe_mul(&t0, &t1, &t0);
  fe_sqn(&t1, &t0, 10);
  fe_mul(&t1, &t1, &t0);
  fe_sqn(&t2, &t1, 20);
  fe_mul(&t1, &t2, &t1);
  fe_sqn(&t1, &t1, 10);
  fe_mul(&t0, &t1, &t0);
  fe_sqn(&t1, &t0, 50);
  fe_mul(&t1, &t1, &t0);
  fe_sqn(&t2, &t1, 100);
  fe_mul(&t1, &t2, &t1);
  fe_sqn(&t1, &t1, 50);
  fe_mul(&t0, &t1, &t0);
  fe_sqn(&t0, &t0, 2);
  fe_mul(out, &t0, z);
}
static void fe_invert(fe* out, const fe* z) {
  fe t, z3;
  fe_pow22523(&t, z);
  fe_sqn(&t, &t, 3);
  fe_sq(&z3, z);
  fe_mul(&z3, &z3, z);
  fe_mul(out, &t, &z3);
}
static int fe_iszero(const fe* f) {
  uint8_t s[32];
  fe_tobytes(s, f);
  uint8_t r = 0;
  for (int i = 0; i < 32; i++) r |= s[i];
  return r == 0;
}
static int fe_isneg(const fe* f) {
  uint8_t s[32];
  fe_tobytes(s, f);
  return s[0] & 1;
}
static int fe_eq(const fe* a, const fe* b) {
  fe d;
  fe_sub(&d, a, b);
  return fe_iszero(&d);
}
static void fe_cneg(fe* h, const fe* f, int b) {
  if (b) fe_neg(h, f); else *h = *f;
}
static void fe_abs(fe* h, const fe* f) { fe_cneg(h, f, fe_isneg(f)); }

static void fe_fromhex_le_words(fe* h, const uint32_t w[8]) {
  uint8_t s[32];
  memcpy(s, w, 32);
  fe_frombytes(h, s);
}
static const uint32_t W_D[8] = {0x135978a3u, 0x75eb4dcau, 0x4141d8abu, 0x00700a4du, 0x7779e898u, 0x8cc74079u, 0x2b6ffe73u, 0x52036ceeu};
static const uint32_t W_D2[8] = {0x26b2f159u, 0xebd69b94u, 0x8283b156u, 0x00e0149au, 0xeef3d130u, 0x198e80f2u, 0x56dffce7u, 0x2406d9dcu};
static const uint32_t W_SQRT_M1[8] = {0x4a0ea0b0u, 0xc4ee1b27u, 0xad2fe478u, 0x2f431806u, 0x3dfbd7a7u, 0x2b4d0099u, 0x4fc1df0bu, 0x2b832480u};
static const uint32_t W_INVSQRT_A_MINUS_D[8] = {0x805d40eau, 0x99c8fdaau, 0x5a4172beu, 0x9d2f1617u, 0xfe01d840u, 0x16c27b91u, 0xcfaffca2u, 0x786c8905u};
static fe K_D, K_D2, K_SQRT_M1, K_INVSQRT_A_MINUS_D;
static int k_init_done = 0;
static void k_init(void) {
  if (k_init_done) return;
  fe_fromhex_le_words(&K_D, W_D);
  fe_fromhex_le_words(&K_D2, W_D2);
  fe_fromhex_le_words(&K_SQRT_M1, W_SQRT_M1);
  fe_fromhex_le_words(&K_INVSQRT_A_MINUS_D, W_INVSQRT_A_MINUS_D);
  k_init_done = 1;
}

/* RFC 9496 §4.2 */
static int fe_sqrt_ratio_m1(fe* r_out, const fe* u, const fe* v) {
  fe v3, v7, r, check, t, neg_u, neg_u_i, r_prime;
  fe_sq(&v3, v); fe_mul(&v3, &v3, v);
  fe_sq(&v7, &v3); fe_mul(&v7, &v7, v);
  fe_mul(&t, u, &v7);
  fe_pow22523(&t, &t);
  fe_mul(&r, u, &v3); fe_mul(&r, &r, &t);
  fe_sq(&check, &r); fe_mul(&check, &check, v);
  fe_neg(&neg_u, u);
  fe_mul(&neg_u_i, &neg_u, &K_SQRT_M1);
  int correct = fe_eq(&check, u);
  int flipped = fe_eq(&check, &neg_u);
  int flipped_i = fe_eq(&check, &neg_u_i);
  fe_mul(&r_prime, &r, &K_SQRT_M1);
  if (flipped | flipped_i) r = r_prime;
  fe_abs(r_out, &r);
  return correct | flipped;
}

/* ------------------------------------------------------------------ group */
typedef struct { fe X, Y, Z, T; } ge;            /* extended */
typedef struct { fe YpX, YmX, Z, T2d; } ge_cached; /* projective Niels */

static void ge_identity(ge* r) { r->X = FE_ZERO; r->Y = FE_ONE; r->Z = FE_ONE; r->T = FE_ZERO; }
static void ge_to_cached(ge_cached* c, const ge* p) {
  fe_add(&c->YpX, &p->Y, &p->X);
  fe_sub(&c->YmX, &p->Y, &p->X);
  c->Z = p->Z;
  fe_mul(&c->T2d, &p->T, &K_D2);
}
/* r = p + q (sub: p - q) */
static void ge_add_cached(ge* r, const ge* p, const ge_cached* q, int sub) {
  fe a, b, c, d, e, f, g, h, t;
  fe_sub(&t, &p->Y, &p->X);
  fe_mul(&a, &t, sub ? &q->YpX : &q->YmX);
  fe_add(&t, &p->Y, &p->X);
  fe_mul(&b, &t, sub ? &q->YmX : &q->YpX);
  fe_mul(&c, &p->T, &q->T2d);
  fe_mul(&d, &p->Z, &q->Z);
  fe_add(&d, &d, &d);
  fe_sub(&e, &b, &a);
  fe_add(&h, &b, &a);
  if (sub) { fe_add(&f, &d, &c); fe_sub(&g, &d, &c); }
  else { fe_sub(&f, &d, &c); fe_add(&g, &d, &c); }
  fe_carry(&f); fe_carry(&g); fe_carry(&h);
  fe_mul(&r->X, &e, &f);
  fe_mul(&r->Y, &g, &h);
  fe_mul(&r->Z, &f, &g);
  fe_mul(&r->T, &e, &h);
}
static void ge_add(ge* r, const ge* p, const ge* q) {
  ge_cached c;
  ge_to_cached(&c, q);
  ge_add_cached(r, p, &c, 0);
}
static void ge_dbl(ge* r, const ge* p) {
  fe a, b, c, ab, xy, e, g, f, h;
  fe_sq(&a, &p->X);
  fe_sq(&b, &p->Y);
  fe_sq(&c, &p->Z);
  fe_add(&c, &c, &c);
  fe_add(&ab, &a, &b);
  fe_add(&xy, &p->X, &p->Y);
  fe_sq(&xy, &xy);
  fe_sub(&e, &xy, &ab);
  fe_sub(&g, &b, &a);
  fe_sub(&f, &g, &c);
  fe_neg(&h, &ab);
  fe_mul(&r->X, &e, &f);
  fe_mul(&r->Y, &g, &h);
  fe_mul(&r->Z, &f, &g);
  fe_mul(&r->T, &e, &h);
}

/* RFC 9496 §4.3.1 */
static int ge_decode(ge* r, const uint8_t in[32]) {
  k_init();
  fe s, ss, u1, u2, u2_sqr, v, t, invsqrt, den_x, den_y, x, y;
  uint8_t chk[32];
  fe_frombytes(&s, in);
  fe_tobytes(chk, &s);
  if (memcmp(chk, in, 32) != 0 || (in[0] & 1)) return 0;
  fe_sq(&ss, &s);
  fe_sub(&u1, &FE_ONE, &ss);
  fe_add(&u2, &FE_ONE, &ss);
  fe_sq(&u2_sqr, &u2);
  fe_sq(&t, &u1); fe_mul(&t, &t, &K_D); fe_neg(&t, &t);
  fe_sub(&v, &t, &u2_sqr);
  fe_mul(&t, &v, &u2_sqr);
  int was_square = fe_sqrt_ratio_m1(&invsqrt, &FE_ONE, &t);
  fe_mul(&den_x, &invsqrt, &u2);
  fe_mul(&den_y, &invsqrt, &den_x); fe_mul(&den_y, &den_y, &v);
  fe_add(&t, &s, &s); fe_mul(&t, &t, &den_x);
  fe_abs(&x, &t);
  fe_mul(&y, &u1, &den_y);
  fe_mul(&t, &x, &y);
  if (!was_square || fe_isneg(&t) || fe_iszero(&y)) return 0;
  r->X = x; r->Y = y; r->Z = FE_ONE; r->T = t;
  return 1;
}
/* RFC 9496 §4.3.2 */
static void ge_encode(uint8_t out[32], const ge* p) {
  k_init();
  fe u1, u2, t, t2, invsqrt, den1, den2, z_inv, ix0, iy0, ench, x, y, den_inv, s;
  fe_add(&t, &p->Z, &p->Y); fe_sub(&t2, &p->Z, &p->Y); fe_mul(&u1, &t, &t2);
  fe_mul(&u2, &p->X, &p->Y);
  fe_sq(&t, &u2); fe_mul(&t, &t, &u1);
  fe_sqrt_ratio_m1(&invsqrt, &FE_ONE, &t);
  fe_mul(&den1, &invsqrt, &u1);
  fe_mul(&den2, &invsqrt, &u2);
  fe_mul(&z_inv, &den1, &den2); fe_mul(&z_inv, &z_inv, &p->T);
  fe_mul(&ix0, &p->X, &K_SQRT_M1);
  fe_mul(&iy0, &p->Y, &K_SQRT_M1);
  fe_mul(&ench, &den1, &K_INVSQRT_A_MINUS_D);
  fe_mul(&t, &p->T, &z_inv);
  if (fe_isneg(&t)) { x = iy0; y = ix0; den_inv = ench; }
  else { x = p->X; y = p->Y; den_inv = den2; }
  fe_mul(&t, &x, &z_inv);
  if (fe_isneg(&t)) fe_neg(&y, &y);
  fe_sub(&t, &p->Z, &y); fe_mul(&s, &den_inv, &t);
  fe_abs(&s, &s);
  fe_tobytes(out, &s);
}

/* ----------------------------------------------------------------- scalars */
/* signed radix-2^w digits of a canonical 256-bit little-endian scalar
 * (dalek Scalar::as_radix_2w): digits in [-2^(w-1), 2^(w-1)), top digit absorbs carry */
static int radix_2w_count(int w) { return w == 8 ? (256 + w - 1) / w + 1 : (256 + w - 1) / w; }
static void to_radix_2w(int8_t* digits, const uint8_t s[32], int w) {
  u64 sc[4];
  memcpy(sc, s, 32);
  u64 radix = 1ULL << w, window_mask = radix - 1, carry = 0;
  int count = (256 + w - 1) / w;
  for (int i = 0; i < count; i++) {
    int bit_offset = i * w, u64_idx = bit_offset / 64, bit_idx = bit_offset % 64;
    u64 bit_buf;
    if (bit_idx < 64 - w || u64_idx == 3) bit_buf = sc[u64_idx] >> bit_idx;
    else bit_buf = (sc[u64_idx] >> bit_idx) | (sc[1 + u64_idx] << (64 - bit_idx));
    u64 coef = carry + (bit_buf & window_mask);
    carry = (coef + radix / 2) >> w;
    digits[i] = (int8_t)((int64_t)coef - (int64_t)(carry << w));
  }
  if (w == 8) digits[count] += (int8_t)carry;
  else digits[count - 1] += (int8_t)(carry << w);
}
/* width-5 non-adjacent form, 256 entries */
static void sc_naf5(int8_t naf[256], const uint8_t s[32]) {
  u64 x[5] = {0, 0, 0, 0, 0};
  memcpy(x, s, 32);
  memset(naf, 0, 256);
  const int w = 5;
  u64 width = 1ULL << w, window_mask = width - 1;
  int pos = 0, carry = 0;
  while (pos < 256) {
    int idx = pos / 64, bit = pos % 64;
    u64 bit_buf;
    if (bit < 64 - w) bit_buf = x[idx] >> bit;
    else bit_buf = (x[idx] >> bit) | (x[1 + idx] << (64 - bit));
    u64 window = carry + (bit_buf & window_mask);
    if ((window & 1) == 0) { pos += 1; continue; }
    if (window < width / 2) { carry = 0; naf[pos] = (int8_t)window; }
    else { carry = 1; naf[pos] = (int8_t)((int64_t)window - (int64_t)width); }
    pos += w;
  }
}

/* ---------------------------------------------------------------- MSM */
static void straus(ge* out, const uint8_t* scalars, const ge* pts, size_t n) {
  int8_t (*nafs)[256] = malloc(n * 256);
  ge_cached (*tables)[8] = malloc(n * sizeof(ge_cached[8]));
  for (size_t i = 0; i < n; i++) {
    sc_naf5(nafs[i], scalars + 32 * i);
    ge p2, cur = pts[i];
    ge_dbl(&p2, &pts[i]);
    ge_cached c2;
    ge_to_cached(&c2, &p2);
    ge_to_cached(&tables[i][0], &cur);
    for (int j = 1; j < 8; j++) {
      ge_add_cached(&cur, &cur, &c2, 0);
      ge_to_cached(&tables[i][j], &cur);
    }
  }
  ge r;
  ge_identity(&r);
  for (int i = 255; i >= 0; i--) {
    ge_dbl(&r, &r);
    for (size_t k = 0; k < n; k++) {
      int d = nafs[k][i];
      if (d > 0) ge_add_cached(&r, &r, &tables[k][d / 2], 0);
      else if (d < 0) ge_add_cached(&r, &r, &tables[k][(-d) / 2], 1);
    }
  }
  *out = r;
  free(nafs);
  free(tables);
}

static void pippenger_column(ge* colsum, const int8_t* digits, int ndig, int col, const ge_cached* cached,
                             size_t n, int w, ge* buckets) {
  size_t nb = (size_t)1 << (w - 1);
  for (size_t b = 0; b < nb; b++) ge_identity(&buckets[b]);
  for (size_t i = 0; i < n; i++) {
    int d = digits[i * ndig + col];
    if (d > 0) ge_add_cached(&buckets[d - 1], &buckets[d - 1], &cached[i], 0);
    else if (d < 0) ge_add_cached(&buckets[-d - 1], &buckets[-d - 1], &cached[i], 1);
  }
  ge inter = buckets[nb - 1], sum = buckets[nb - 1];
  for (size_t b = nb - 1; b-- > 0;) {
    ge_add(&inter, &inter, &buckets[b]);
    ge_add(&sum, &sum, &inter);
  }
  *colsum = sum;
}

static void pippenger(ge* out, const uint8_t* scalars, const ge* pts, size_t n, int threads) {
  int w = n < 500 ? 6 : (n < 800 ? 7 : 8);
  int ndig = radix_2w_count(w);
  size_t nb = (size_t)1 << (w - 1);
  int8_t* digits = calloc(n * ndig, 1);
  ge_cached* cached = malloc(n * sizeof(ge_cached));
  for (size_t i = 0; i < n; i++) {
    to_radix_2w(digits + i * ndig, scalars + 32 * i, w);
    ge_to_cached(&cached[i], &pts[i]);
  }
  ge* cols = malloc(ndig * sizeof(ge));
  if (threads <= 1) {
    ge* buckets = malloc(nb * sizeof(ge));
    for (int c = 0; c < ndig; c++) pippenger_column(&cols[c], digits, ndig, c, cached, n, w, buckets);
    free(buckets);
  } else {
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
    {
      ge* buckets = malloc(nb * sizeof(ge));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
      for (int c = 0; c < ndig; c++) pippenger_column(&cols[c], digits, ndig, c, cached, n, w, buckets);
      free(buckets);
    }
  }
  ge total = cols[ndig - 1];
  for (int c = ndig - 2; c >= 0; c--) {
    for (int k = 0; k < w; k++) ge_dbl(&total, &total);
    ge_add(&total, &total, &cols[c]);
  }
  *out = total;
  free(cols);
  free(cached);
  free(digits);
}

/* out = sum scalars[i] * points[i]; returns 0, or -5 on an invalid point encoding */
int oracle_msm(const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t out[32], int threads) {
  k_init();
  ge* pts = malloc((n ? n : 1) * sizeof(ge));
  int bad = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads > 1 ? threads : 1) reduction(| : bad)
#endif
  for (size_t i = 0; i < n; i++)
    if (!ge_decode(&pts[i], points + 32 * i)) bad |= 1;
  if (bad) { free(pts); return -5; }
  ge r;
  if (n == 0) ge_identity(&r);
  else if (n < 190) straus(&r, scalars, pts, n);
  else pippenger(&r, scalars, pts, n, threads);
  ge_encode(out, &r);
  free(pts);
  return 0;
}

/* Same, points already decoded by oracle_decode_points (so that the timed region
 * of a baseline run can exclude decompression, like the GPU's resident table). */
typedef struct { ge* p; size_t n; } oracle_points;
oracle_points* oracle_points_decode(const uint8_t* points, size_t n, int threads) {
  k_init();
  oracle_points* t = malloc(sizeof *t);
  t->p = malloc((n ? n : 1) * sizeof(ge));
  t->n = n;
  int bad = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads > 1 ? threads : 1) reduction(| : bad)
#endif
  for (size_t i = 0; i < n; i++)
    if (!ge_decode(&t->p[i], points + 32 * i)) bad |= 1;
  if (bad) { free(t->p); free(t); return NULL; }
  return t;
}
void oracle_points_free(oracle_points* t) { if (t) { free(t->p); free(t); } }
int oracle_msm_decoded(const uint8_t* scalars, const oracle_points* t, size_t offset, size_t n, uint8_t out[32],
                       int threads) {
  if (offset + n > t->n) return -4;
  ge r;
  if (n == 0) ge_identity(&r);
  else if (n < 190) straus(&r, scalars, t->p + offset, n);
  else pippenger(&r, scalars, t->p + offset, n, threads);
  ge_encode(out, &r);
  return 0;
}

/* k * B for many scalars (fixed-base, used to make synthetic point sets quickly) */
static const uint8_t BASE_ENC[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9,
                                     0x61, 0xc5, 0x00, 0x51, 0x5f, 0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82,
                                     0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};
void oracle_basepoint_mul(const uint8_t* scalars, size_t n, uint8_t* out, int threads) {
  k_init();
  ge B;
  ge_decode(&B, BASE_ENC);
  /* table[j][d] = (d+1) * 16^j * B, j < 64, d < 8 */
  static ge_cached table[64][8];
  static int have = 0;
  if (!have) {
    ge cur = B;
    for (int j = 0; j < 64; j++) {
      ge m = cur;
      ge_cached c1;
      ge_to_cached(&c1, &cur);
      for (int d = 0; d < 8; d++) {
        ge_to_cached(&table[j][d], &m);
        ge_add_cached(&m, &m, &c1, 0);
      }
      for (int k = 0; k < 4; k++) ge_dbl(&cur, &cur);
    }
    have = 1;
  }
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads > 1 ? threads : 1)
#endif
  for (size_t i = 0; i < n; i++) {
    int8_t dg[64];
    to_radix_2w(dg, scalars + 32 * i, 4);
    ge r;
    ge_identity(&r);
    for (int j = 0; j < 64; j++) {
      int d = dg[j];
      if (d > 0) ge_add_cached(&r, &r, &table[j][d - 1], 0);
      else if (d < 0) ge_add_cached(&r, &r, &table[j][-d - 1], 1);
    }
    ge_encode(out + 32 * i, &r);
  }
}

/* ---------------------------------------------------------------- scalars mod l (4x64 Montgomery)
 * For the CPU restatement of InnerProductProof::create below (the reference's scalars are a
 * Montgomery 4x64 field too: ark-ff Fp256). */
typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } scm;
static const uint64_t SC_L[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0x0ULL, 0x1000000000000000ULL};
static const uint64_t SC_RR[4] = {0xa40611e3449c0f01ULL, 0xd00e1ba768859347ULL, 0xceec73d217f5be65ULL, 0x0399411b7c309a3dULL};
static const uint64_t SC_NINV = 0xd2b51da312547e1bULL; /* -l^-1 mod 2^64 */
static int scm_geq_l(const uint64_t x[4]) {
  for (int i = 3; i >= 0; i--) { if (x[i] > SC_L[i]) return 1; if (x[i] < SC_L[i]) return 0; }
  return 1;
}
static void scm_sub_l(uint64_t x[4]) {
  u128 bw = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)x[i] - SC_L[i] - bw; x[i] = (uint64_t)t; bw = (t >> 64) & 1; }
}
static void scm_montmul(scm* r, const scm* a, const scm* b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { u128 s = (u128)a->v[j] * b->v[i] + t[j] + c; t[j] = (uint64_t)s; c = s >> 64; }
    u128 s = (u128)t[4] + c; t[4] = (uint64_t)s; t[5] = (uint64_t)(s >> 64);
    uint64_t m = t[0] * SC_NINV;
    c = ((u128)m * SC_L[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) { u128 s2 = (u128)m * SC_L[j] + t[j] + c; t[j - 1] = (uint64_t)s2; c = s2 >> 64; }
    s = (u128)t[4] + c; t[3] = (uint64_t)s; t[4] = t[5] + (uint64_t)(s >> 64); t[5] = 0;
  }
  memcpy(r->v, t, 32);
  if (t[4] || scm_geq_l(r->v)) scm_sub_l(r->v);
}
static void scm_add(scm* r, const scm* a, const scm* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { u128 s = (u128)a->v[i] + b->v[i] + c; r->v[i] = (uint64_t)s; c = s >> 64; }
  if (scm_geq_l(r->v)) scm_sub_l(r->v);
}
static void scm_from_bytes(scm* r, const uint8_t b[32]) { scm a, rr; memcpy(a.v, b, 32); memcpy(rr.v, SC_RR, 32); scm_montmul(r, &a, &rr); }
static void scm_to_bytes(uint8_t out[32], const scm* a) { scm one = {{1, 0, 0, 0}}, n; scm_montmul(&n, a, &one); memcpy(out, n.v, 32); }
static void scm_invert(scm* r, const scm* a) { /* a^(l-2) */
  static const uint64_t e[4] = {0x5812631a5cf5d3ebULL, 0x14def9dea2f79cd6ULL, 0x0ULL, 0x1000000000000000ULL};
  scm one_b = {{1, 0, 0, 0}}, rr, acc, base = *a;
  memcpy(rr.v, SC_RR, 32);
  scm_montmul(&acc, &one_b, &rr); /* Montgomery one */
  for (int i = 0; i < 256; i++) {
    if ((e[i >> 6] >> (i & 63)) & 1) scm_montmul(&acc, &acc, &base);
    scm_montmul(&base, &base, &base);
  }
  *r = acc;
}

/* ---------------------------------------------------------------- InnerProductProof::create on the CPU
 * Restates reference src/inner_product_proof.rs:49-193 and fold_witness :202-248 as the fork runs
 * them on the CPU: cross terms serially, L and R as variable-time MSMs, in round 0 the 2n
 * generators multiplied by their factors one after the other (:125-134), every fold a two-term
 * MSM per generator, thread-parallel (rayon there, OpenMP here) only at or above
 * PARALLELISM_THRESHOLD = 10 (:26, :217).  The challenges are inputs (u_j, canonical), so that the
 * result can be compared with oracle/protocol.py and the run timed without a transcript.
 * out_LR: lg n x 64 bytes (L_j | R_j); out_ab: a | b.  Returns 0, -3 (not a power of two), -5. */
static void msm_any(ge* r, const uint8_t* scalars, const ge* pts, size_t n, int threads) {
  if (n == 0) ge_identity(r);
  else if (n < 190) straus(r, scalars, pts, n);
  else pippenger(r, scalars, pts, n, threads);
}
int oracle_ipp_create(size_t n, const uint8_t Q_enc[32], const uint8_t* G_factors, const uint8_t* H_factors,
                      const uint8_t* G_enc, const uint8_t* H_enc, const uint8_t* a_in, const uint8_t* b_in,
                      const uint8_t* challenges, uint8_t* out_LR, uint8_t out_ab[64], int threads) {
  k_init();
  if (n == 0 || (n & (n - 1))) return -3;
  ge* G = malloc(n * sizeof(ge));
  ge* H = malloc(n * sizeof(ge));
  ge Q;
  scm* a = malloc(n * sizeof(scm));
  scm* b = malloc(n * sizeof(scm));
  int bad = !ge_decode(&Q, Q_enc);
  for (size_t i = 0; i < n; i++) {
    bad |= !ge_decode(&G[i], G_enc + 32 * i);
    bad |= !ge_decode(&H[i], H_enc + 32 * i);
    scm_from_bytes(&a[i], a_in + 32 * i);
    scm_from_bytes(&b[i], b_in + 32 * i);
  }
  if (bad) { free(G); free(H); free(a); free(b); return -5; }
  uint8_t* sc = malloc((n + 1) * 32);
  ge* pts = malloc((n + 1) * sizeof(ge));
  size_t m = n;
  int round = 0;
  while (m != 1) {
    size_t h = m / 2;
    /* c_L = <a_lo, b_hi>, c_R = <a_hi, b_lo>  (:87-88, :156-157) */
    scm cl = {{0, 0, 0, 0}}, cr = {{0, 0, 0, 0}}, t;
    for (size_t i = 0; i < h; i++) {
      scm_montmul(&t, &a[i], &b[h + i]); scm_add(&cl, &cl, &t);
      scm_montmul(&t, &a[h + i], &b[i]); scm_add(&cr, &cr, &t);
    }
    for (int side = 0; side < 2; side++) { /* L then R (:90-114, :159-172) */
      for (size_t i = 0; i < h; i++) {
        scm x = side == 0 ? a[i] : a[h + i];       /* a_lo with G_hi | a_hi with G_lo */
        scm y = side == 0 ? b[h + i] : b[i];       /* b_hi with H_lo | b_lo with H_hi */
        if (round == 0 && G_factors) { scm f; scm_from_bytes(&f, G_factors + 32 * (side == 0 ? h + i : i)); scm_montmul(&x, &x, &f); }
        if (round == 0 && H_factors) { scm f; scm_from_bytes(&f, H_factors + 32 * (side == 0 ? i : h + i)); scm_montmul(&y, &y, &f); }
        scm_to_bytes(sc + 32 * i, &x);
        scm_to_bytes(sc + 32 * (h + i), &y);
        pts[i] = side == 0 ? G[h + i] : G[i];
        pts[h + i] = side == 0 ? H[i] : H[h + i];
      }
      scm_to_bytes(sc + 32 * m, side == 0 ? &cl : &cr);
      pts[m] = Q;
      ge r;
      msm_any(&r, sc, pts, m + 1, threads);
      ge_encode(out_LR + 64 * round + 32 * side, &r);
    }
    scm u, ui;
    scm_from_bytes(&u, challenges + 32 * round);
    scm_invert(&ui, &u);
    if (round == 0) { /* G_i <- g_i G_i, H_i <- h_i H_i, one after the other (:125-134) */
      for (size_t i = 0; i < m; i++) {
        if (G_factors) { ge r; straus(&r, G_factors + 32 * i, &G[i], 1); G[i] = r; }
        if (H_factors) { ge r; straus(&r, H_factors + 32 * i, &H[i], 1); H[i] = r; }
      }
    }
    uint8_t ub[32], uib[32];
    scm_to_bytes(ub, &u);
    scm_to_bytes(uib, &ui);
    /* fold_witness (:202-248): serial below the threshold, thread-parallel at or above it */
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads > 1 ? threads : 1) if (h >= 10 && threads > 1)
#endif
    for (size_t i = 0; i < h; i++) {
      scm t1, t2;
      scm_montmul(&t1, &a[i], &u); scm_montmul(&t2, &a[h + i], &ui); scm_add(&a[i], &t1, &t2);   /* a_lo u + u^-1 a_hi */
      scm_montmul(&t1, &b[i], &ui); scm_montmul(&t2, &b[h + i], &u); scm_add(&b[i], &t1, &t2);   /* b_lo u^-1 + u b_hi */
      uint8_t s2[64];
      ge p2[2], r;
      memcpy(s2, uib, 32); memcpy(s2 + 32, ub, 32);
      p2[0] = G[i]; p2[1] = G[h + i];
      straus(&r, s2, p2, 2); G[i] = r;                                                              /* u^-1 G_lo + u G_hi */
      memcpy(s2, ub, 32); memcpy(s2 + 32, uib, 32);
      p2[0] = H[i]; p2[1] = H[h + i];
      straus(&r, s2, p2, 2); H[i] = r;                                                              /* u H_lo + u^-1 H_hi */
    }
    m = h;
    round++;
  }
  scm_to_bytes(out_ab, &a[0]);
  scm_to_bytes(out_ab + 32, &b[0]);
  free(G); free(H); free(a); free(b); free(sc); free(pts);
  return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
