// TEST INFRASTRUCTURE — host build of the __host__ __device__ arithmetic in
// mpc_bulletproof_b200/csrc/{fe,sc,ge}.cuh so that the limb schedules and the
// Pippenger bookkeeping can be checked against oracle/ on a machine without a GPU.
// Nothing in the product loads this library.
#include <vector>
#include <cstring>
#define BPG_HOSTSIM 1
#include <thread>
#include <pthread.h>
#include "../../mpc_bulletproof_b200/csrc/ge.cuh"
#include "../../mpc_bulletproof_b200/csrc/fe16.cuh"
#include "../../mpc_bulletproof_b200/csrc/sc.cuh"
#include "../../mpc_bulletproof_b200/csrc/stark_pt.cuh"
#include "../../mpc_bulletproof_b200/csrc/stark_sc.cuh"
using namespace bpg;

static fe ld(const uint32_t* p) { fe a; for (int i = 0; i < 8; i++) a.v[i] = p[i]; return a; }
static void st(uint32_t* p, const fe& a) { for (int i = 0; i < 8; i++) p[i] = a.v[i]; }
static ge_ext lde(const uint32_t* p) { ge_ext e; e.X = ld(p); e.Y = ld(p + 8); e.Z = ld(p + 16); e.T = ld(p + 24); return e; }
static void ste(uint32_t* p, const ge_ext& e) { st(p, e.X); st(p + 8, e.Y); st(p + 16, e.Z); st(p + 24, e.T); }
static ge_niels ldn(const uint32_t* p) { ge_niels e; e.ypx = ld(p); e.ymx = ld(p + 8); e.t2d = ld(p + 16); return e; }
static void stn(uint32_t* p, const ge_niels& e) { st(p, e.ypx); st(p + 8, e.ymx); st(p + 16, e.t2d); }

static fp ldp(const uint32_t* p) { fp a; for (int i = 0; i < 8; i++) a.v[i] = p[i]; return a; }
static void stp(uint32_t* p, const fp& a) { for (int i = 0; i < 8; i++) p[i] = a.v[i]; }
static sp_xyzz ldx(const uint32_t* p) { sp_xyzz e; e.X = ldp(p); e.Y = ldp(p + 8); e.ZZ = ldp(p + 16); e.ZZZ = ldp(p + 24); return e; }
static void stx(uint32_t* p, const sp_xyzz& e) { stp(p, e.X); stp(p + 8, e.Y); stp(p + 16, e.ZZ); stp(p + 24, e.ZZZ); }

extern "C" {
// ---- Stark-curve policy (stark_fp.cuh, stark_pt.cuh) ----
void hs_fp_mul(const uint32_t* a, const uint32_t* b, uint32_t* o) { stp(o, fp_mul(ldp(a), ldp(b))); }
void hs_fp_add(const uint32_t* a, const uint32_t* b, uint32_t* o) { stp(o, fp_add(ldp(a), ldp(b))); }
void hs_fp_sub(const uint32_t* a, const uint32_t* b, uint32_t* o) { stp(o, fp_sub(ldp(a), ldp(b))); }
void hs_fp_invert(const uint32_t* a, uint32_t* o) { stp(o, fp_invert(ldp(a))); }
void hs_fp_to_mont(const uint32_t* a, uint32_t* o) { stp(o, fp_to_mont(ldp(a))); }
void hs_fp_from_mont(const uint32_t* a, uint32_t* o) { stp(o, fp_from_mont(ldp(a))); }
int hs_sp_from_affine(const uint32_t* xy, uint32_t* aff) { sp_aff q; bool ok = sp_from_affine_words(q, xy); stp(aff, q.x); stp(aff + 8, q.y); return ok; }
void hs_sp_to_affine(const uint32_t* x, uint32_t* xy) { sp_to_affine_words(xy, ldx(x)); }
void hs_sp_from_aff(const uint32_t* aff, uint32_t* o) { sp_aff q; q.x = ldp(aff); q.y = ldp(aff + 8); stx(o, sp_from_aff(q)); }
void hs_sp_madd(const uint32_t* x, const uint32_t* aff, int neg, uint32_t* o) { sp_aff q; q.x = ldp(aff); q.y = ldp(aff + 8); stx(o, sp_madd(ldx(x), q, neg)); }
void hs_sp_add(const uint32_t* a, const uint32_t* b, uint32_t* o) { stx(o, sp_add(ldx(a), ldx(b))); }
void hs_sp_dbl(const uint32_t* a, uint32_t* o) { stx(o, sp_dbl(ldx(a))); }
void hs_scs_montmul(const uint32_t* a, const uint32_t* b, uint32_t* o) {
  sc x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); sc r = scs_montmul(x, y); memcpy(o, r.v, 32);
}
void hs_scs_add(const uint32_t* a, const uint32_t* b, uint32_t* o) {
  sc x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); sc r = scs_add(x, y); memcpy(o, r.v, 32);
}
void hs_scs_sub(const uint32_t* a, const uint32_t* b, uint32_t* o) {
  sc x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); sc r = scs_sub(x, y); memcpy(o, r.v, 32);
}
void hs_fe_mul(const uint32_t* a, const uint32_t* b, uint32_t* o) { st(o, fe_mul(ld(a), ld(b))); }
void hs_fe_sq(const uint32_t* a, uint32_t* o) { st(o, fe_sq(ld(a))); }
void hs_fe_add(const uint32_t* a, const uint32_t* b, uint32_t* o) { st(o, fe_add(ld(a), ld(b))); }
void hs_fe_sub(const uint32_t* a, const uint32_t* b, uint32_t* o) { st(o, fe_sub(ld(a), ld(b))); }
void hs_fe_canon(const uint32_t* a, uint32_t* o) { st(o, fe_canon(ld(a))); }
void hs_fe_invert(const uint32_t* a, uint32_t* o) { st(o, fe_invert(ld(a))); }
int hs_decode(const uint8_t* b, uint32_t* ext) { ge_ext e; bool ok = ge_decode(e, b); ste(ext, e); return ok; }
void hs_encode(const uint32_t* ext, uint8_t* b) { ge_encode(b, lde(ext)); }
void hs_to_niels(const uint32_t* ext, uint32_t* n) { stn(n, ge_to_niels(lde(ext))); }
void hs_madd(const uint32_t* ext, const uint32_t* n, int neg, uint32_t* o) { ste(o, ge_madd(lde(ext), ldn(n), neg)); }
void hs_add(const uint32_t* a, const uint32_t* b, uint32_t* o) { ste(o, ge_add(lde(a), lde(b))); }
void hs_dbl(const uint32_t* a, uint32_t* o) { ste(o, ge_dbl(lde(a))); }
void hs_neg(const uint32_t* a, uint32_t* o) { ste(o, ge_neg(lde(a))); }
void hs_sc_montmul(const uint32_t* a, const uint32_t* b, uint32_t* o) {
  sc x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); sc r = sc_montmul(x, y); memcpy(o, r.v, 32);
}
void hs_sc_mul(const uint32_t* a, const uint32_t* b, uint32_t* o) {
  sc x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); sc r = sc_mul(x, y); memcpy(o, r.v, 32);
}
void hs_sc_add(const uint32_t* a, const uint32_t* b, uint32_t* o) {
  sc x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); sc r = sc_add(x, y); memcpy(o, r.v, 32);
}
void hs_sc_sub(const uint32_t* a, const uint32_t* b, uint32_t* o) {
  sc x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); sc r = sc_sub(x, y); memcpy(o, r.v, 32);
}
// Montgomery form of (ChaCha20 block `block` under `key`, nonce "bpg sLsR v01") mod l
void hs_chacha_scalar(const uint32_t* key, uint32_t block, uint32_t* o) {
  ChaKey k; memcpy(k.k, key, 32); sc r = sc_from_mont(sc_from_chacha_block(k, block)); memcpy(o, r.v, 32);
}
static sc_bias mk_bias(int c, int W) {
  sc_bias b; memset(&b, 0, sizeof b);
  for (int w = 0; w < W; w++) { int bit = c * w + c - 1; b.v[bit >> 5] |= 1u << (bit & 31); }
  return b;
}
// digits[W] for scalar k with window c
int hs_digits(const uint32_t* k, int c, int* digits) {
  int W = (255 + c - 1) / c;
  sc_recoded r = sc_recode(k, mk_bias(c, W));
  for (int w = 0; w < W; w++) digits[w] = sc_digit(r, w, c);
  return W;
}
// Serial walk through the same stages as msm_kernels.cuh (buckets -> chunked
// running sums -> window combine -> Horner), with the same device functions.
int hs_msm(const uint8_t* scalars, const uint8_t* points, int n, int c, int chunk, uint8_t* out) {
  int W = (255 + c - 1) / c;
  uint32_t nb = 1u << (c - 1);
  if ((uint32_t)chunk > nb) chunk = nb;
  sc_bias bias = mk_bias(c, W);
  std::vector<ge_niels> tab(n);
  for (int i = 0; i < n; i++) {
    ge_ext e;
    if (!ge_decode(e, points + 32 * i)) return -1;
    tab[i] = ge_affine_to_niels(e.X, e.Y);
  }
  std::vector<ge_ext> wins(W);
  for (int w = 0; w < W; w++) {
    std::vector<ge_ext> B(nb, ge_identity());
    for (int i = 0; i < n; i++) {
      uint32_t k[8]; memcpy(k, scalars + 32 * i, 32);
      int d = sc_digit(sc_recode(k, bias), w, c);
      if (d) { uint32_t m = d < 0 ? -d : d; B[m - 1] = ge_madd(B[m - 1], tab[i], d < 0); }
    }
    ge_ext wsum = ge_identity();
    for (uint32_t q = 0; q < nb / chunk; q++) {
      ge_ext run = ge_identity(), acc = ge_identity();
      for (int j = chunk - 1; j >= 0; j--) { run = ge_add(run, B[q * chunk + j]); acc = ge_add(acc, run); }
      uint32_t m = q * chunk;
      if (m) {
        ge_ext t = run; int top = 31 - __builtin_clz(m);
        for (int bit = top - 1; bit >= 0; bit--) { t = ge_dbl(t); if ((m >> bit) & 1) t = ge_add(t, run); }
        acc = ge_add(acc, t);
      }
      wsum = ge_add(wsum, acc);
    }
    wins[w] = wsum;
  }
  ge_ext acc = wins[W - 1];
  for (int w = W - 2; w >= 0; w--) { for (int i = 0; i < c; i++) acc = ge_dbl(acc); acc = ge_add(acc, wins[w]); }
  ge_encode(out, acc);
  return 0;
}
}  // extern "C"

// ---- sixteen-lane field layer (fe16.cuh): one host thread per lane ----
struct Team16 {
  pthread_barrier_t bar;
  uint32_t slots[64];
  alignas(16) uint32_t sm[G16_WORDS];
};
template <class F>
static void run_lanes(int lanes, F f) {  // 16: one half-warp; 32: a whole warp (the WIDE form)
  Team16 t;
  pthread_barrier_init(&t.bar, nullptr, lanes);
  std::vector<std::thread> th;
  for (uint32_t k = 0; k < (uint32_t)lanes; k++)
    th.emplace_back([&t, &f, k] {
      grp16 g;
      g.sm = t.sm; g.k = k & 15; g.half = k >> 4; g.par = 0; g.bar = &t.bar; g.slots = t.slots;
      f(g);
    });
  for (auto& x : th) x.join();
  pthread_barrier_destroy(&t.bar);
}
template <class F>
static void run16(F f) { run_lanes(16, f); }
extern "C" {
// op 0: a*b, 1: a+b, 2: a-b, 3: (a*b)^(2^n) by n squarings of the product; limbs = the raw lazy limbs
void hs_fe16_op(const uint32_t* a, const uint32_t* b, int op, int n, uint32_t* o, uint32_t* limbs) {
  fe A = ld(a), B = ld(b);
  run16([&](grp16& g) {
    fe16 x = fe16_from_fe(g, A), y = fe16_from_fe(g, B), r;
    if (op == 0) r = fe16_mul<false>(g, x, y);
    else if (op == 1) r = fe16_add(g, x, y);
    else if (op == 2) r = fe16_sub(g, x, y);
    else { r = fe16_mul<false>(g, x, y); for (int i = 0; i < n; i++) { r = fe16_sq<false>(g, r); if (limbs[g.k] < r.l) limbs[g.k] = r.l; } }
    if (op != 3) limbs[g.k] = r.l;
    fe c = fe16_to_fe(g, r);
    if (g.k == 5) st(o, c);
  });
}
// ((a - b) * (a + b) - a)^2 with every intermediate lazy: exercises products of carried sums and differences
void hs_fe16_mix(const uint32_t* a, const uint32_t* b, uint32_t* o) {
  fe A = ld(a), B = ld(b);
  run16([&](grp16& g) {
    fe16 x = fe16_from_fe(g, A), y = fe16_from_fe(g, B);
    fe16 m = fe16_mul<false>(g, fe16_sub(g, x, y), fe16_add(g, x, y));
    fe16 d = fe16_sub(g, m, x);
    fe16 zero; zero.l = 0;
    fe16 r = fe16_mul<false>(g, fe16_sub(g, zero, d), fe16_sub(g, zero, d));
    fe c = fe16_to_fe(g, r);
    if (g.k == 0) st(o, c);
  });
}
void hs_fe16_pow22523(const uint32_t* a, uint32_t* o) {
  fe A = ld(a);
  run16([&](grp16& g) {
    fe c = fe16_to_fe(g, fe16_pow22523<false>(g, fe16_from_fe(g, A)));
    if (g.k == 15) st(o, c);
  });
}
void hs_from_uniform(const uint8_t* in64, uint8_t* out32) {
  ge_ext p = ge_add(ge_elligator_map(fe_from_bytes_255(in64)), ge_elligator_map(fe_from_bytes_255(in64 + 32)));
  ge_encode(out32, p);
}
int hs_decode16(const uint8_t* b, int wide, uint32_t* ext) {
  int ok = 0;
  run_lanes(wide ? 32 : 16, [&](grp16& g) {
    ge_ext e;
    bool o = wide ? ge_decode16<true>(g, b, e) : ge_decode16<false>(g, b, e);
    if (g.k == 7 && g.half == 0) {
      ste(ext, e);
      ok = o;
    }
  });
  return ok;
}
void hs_encode16(const uint32_t* ext, uint8_t* b) {
  ge_ext e = lde(ext);
  run16([&](grp16& g) {
    fe s = ge_encode16<false>(g, e);
    if (g.k == 3) fe_to_bytes(b, s);
  });
}
// the whole-warp form: 32 lanes, the half-warps split every product
void hs_encode32(const uint32_t* ext, uint8_t* b) {
  ge_ext e = lde(ext);
  run_lanes(32, [&](grp16& g) {
    fe s = ge_encode16<true>(g, e);
    if (g.k == 3 && g.half == 1) fe_to_bytes(b, s);
  });
}
void hs_fe16_wide_chain(const uint32_t* a, const uint32_t* b, int n, uint32_t* o, uint32_t* limbs) {
  fe A = ld(a), B = ld(b);
  run_lanes(32, [&](grp16& g) {
    fe16 r = fe16_mul<true>(g, fe16_from_fe(g, A), fe16_from_fe(g, B));
    for (int i = 0; i < n; i++) { r = fe16_sq<true>(g, r); if (g.half == 0 && limbs[g.k] < r.l) limbs[g.k] = r.l; }
    fe c = fe16_to_fe(g, r);
    if (g.k == 0 && g.half == 0) st(o, c);
  });
}
}
