// libbpgpu: inner-product argument, device-resident state, one MSM per round.
#define BPG_GE_OUTLINE 1  // latency-bound kernels: one out-of-line copy of each point operation (comb_kernels.cuh)
#include "internal.cuh"
#include "ipp_kernels.cuh"
#include "comb_kernels.cuh"
#include "comb_msm_kernels.cuh"

using namespace bpg;

int launch_comb_msm(bpg_ctx* ctx, cudaStream_t s, const uint32_t* comb_cached, const uint32_t* d_scalars, size_t n,
                    uint32_t* parts, uint32_t* ticket, uint32_t* out_ext) {
  CombMsm M;
  M.comb = comb_cached;
  M.scalars = d_scalars;
  M.n = (uint32_t)n;
  M.wsplit = 16;  // four windows per thread: a few hundred threads per term set, a handful of blocks
  while (M.wsplit > 1 && (n * M.wsplit + CB_THREADS - 1) / CB_THREADS > ADHOC_PARTS) M.wsplit >>= 1;
  M.bias4 = bias_for(4);
  M.parts = parts;
  M.ticket = ticket;
  const unsigned grid = (unsigned)((n * M.wsplit + CB_THREADS - 1) / CB_THREADS);
  if (grid > ADHOC_PARTS) return BPG_ERR_ARG;
  k_comb_msm<<<grid, CB_THREADS, 0, s>>>(M, out_ext);
  LAUNCH_CHECK();
  return BPG_OK;
}

// sets of indexed terms over the affine combs of a table, encoded: `single`, `lo`, `hi` per set (k_comb_terms)
int launch_comb_terms(bpg_ctx* ctx, cudaStream_t s, const uint32_t* comb_affine, const uint32_t* d_scalars,
                      const uint32_t* d_point_ids, const uint32_t single[4], const uint32_t lo[4], const uint32_t hi[4],
                      int nsets, uint8_t* d_out_bytes, uint32_t* d_out_ext) {
  if (nsets < 1 || nsets > 4 || (!d_out_bytes && !d_out_ext)) return BPG_ERR_ARG;
  CombTerms M;
  memset(&M, 0, sizeof M);
  M.comb = comb_affine;
  M.scalars = d_scalars;
  M.point_ids = d_point_ids;
  size_t max_units = 0, total = 0;
  for (int i = 0; i < nsets; i++) {
    M.single[i] = single[i];
    M.lo[i] = lo[i];
    M.hi[i] = hi[i];
    max_units = std::max<size_t>(max_units, 1 + hi[i] - lo[i]);
    total += 1 + hi[i] - lo[i];
  }
  static const size_t target_mul = env_size("BPG_COMB_THREADS_PER_SM", 192);
  M.wsplit = 1;
  while (M.wsplit < COMB_WINDOWS && total * M.wsplit < (size_t)ctx->sm_count * target_mul) M.wsplit <<= 1;
  M.bias4 = bias_for(4);
  const unsigned bx = (unsigned)((max_units * M.wsplit + CB_THREADS - 1) / CB_THREADS);
  int rc = ensure_ws(ctx, (size_t)nsets * bx * 128);
  if (rc) return rc;
  uint32_t* parts = (uint32_t*)ctx->ws;
  prof_mark(ctx, BPG_PROF_ACCUM);
  k_comb_terms<<<dim3(bx, nsets), CB_THREADS, 0, s>>>(M, parts);
  LAUNCH_CHECK();
  prof_mark(ctx, BPG_PROF_ENCODE);
  k_parts_encode<<<nsets, CBQ_THREADS, 0, s>>>(parts, bx, d_out_bytes, d_out_ext);
  LAUNCH_CHECK();
  prof_mark(ctx, -1);
  return BPG_OK;
}

// ---------------------------------------------------------------------------
// inner-product argument: device-resident state, one MSM per round
// ---------------------------------------------------------------------------
struct bpg_ipp {
  bpg_ctx* ctx;
  size_t n;        // original length (power of two)
  size_t m;        // current length
  const bpg_table* tab;  // windowed table the round MSMs run over
  bpg_table* own_tab;    // non-null when the state built its own [G | H | Q] table
  bool has_qmul;         // cross terms are multiplied by q_mul (Q = q_mul * table[q_id])
  bool q_sep;            // Q is outside the (caller's windowed) table: c_L Q, c_R Q come from a comb of Q
  uint32_t *q_comb, *q_side;
  uint8_t* buf;    // one allocation for everything below
  uint32_t *a, *b, *wG, *wH, *scalars, *point_ids, *partials, *u_pair, *out_ext;
  uint32_t* q_mul;
  uint8_t *set_ids, *out_bytes;
  bool lr_done;
  int lanes;        // (a, b) pairs folding together: 1, or a party's share + MAC vectors (r1cs_mpc)
  uint32_t* c_ext;  // [lanes][2] caller-supplied cross terms of the current round (shares path)
  // comb rounds (comb_kernels.cuh)
  int mode;               // 0: bucket MSM over the windowed table (no-fold form); 1: comb rounds
  size_t m0;              // mode 0: materialise the folded generators when the vectors have shrunk to m0 (0 = never)
  size_t n_eff;           // generators per vector in the current representation (n, or m0 once materialised)
  bool comb_affine;       // comb rounds read the table's generator combs (true) or the proof's own combs of folded generators
  const uint32_t* comb;   // the combs the rounds read
  uint32_t cg_id, ch_id;  // comb index of G_0 / H_0
  const uint32_t* cq_comb;  // comb of Q's base point (affine)
  uint32_t* mat_buf;      // folded generators | doubling chains | their combs (per proof)
  uint32_t* parts;        // partial sums of a comb round's blocks
  size_t parts_cap;
  bool cross_ready;       // the fold kernel already left this round's cross-term partials
  uint32_t ncross;
  // single-prover comb rounds take the previous fold on the fly (CombRound::fold): the challenge waits here and the
  // folded vectors land in the alternate buffers, which then become the current ones
  bool shares;            // created by bpg_ipp_begin_shares: the caller reads the folded vectors between rounds
  bool fold_pending;
  IppPair fold_up;
  uint32_t* alt;          // a2 | b2 | wG2 | wH2, n_eff scalars each
  size_t alt_n;
  uint32_t* oth[4];       // the vectors (a, b, wG, wH) that are not current: targets of the next fused fold
};

// BPG_IPP_DIRECT_MAX: vectors up to this length run every round on the generators' own combs;
// BPG_IPP_M0: longer ones switch to combs of the folded generators once they have shrunk to this length


// d_* pointers are device pointers; factors may be null (all ones).
// Either (G, H, Q_host) are given and the state builds its own windowed [G | H | Q] table,
// or `shared` is a windowed table that already holds the generators at g_base/h_base and a
// base point at q_id with Q = q_mul * shared[q_id] (the R1CS prover's Q = w*B).
int ipp_begin_dev(bpg_ctx* ctx, const bpg_table* G, size_t g_off, const bpg_table* H, size_t h_off, size_t n,
                         const uint8_t* Q_host, const bpg_table* shared, size_t g_base, size_t h_base, size_t q_id,
                         const uint8_t* q_mul_host, const uint32_t* d_gf, const uint32_t* d_hf, const uint32_t* d_a,
                         const uint32_t* d_b, bpg_ipp** out, int lanes) {
  if (n == 0 || (n & (n - 1))) return BPG_ERR_POW2;
  if (n >= (1u << 28)) return BPG_ERR_ARG;
  if (lanes < 1 || lanes > IPP_MAX_LANES || (lanes > 1 && !shared)) return BPG_ERR_ARG;
  if (shared) {
    if (g_base + n > shared->n || h_base + n > shared->n || q_id >= shared->n) return BPG_ERR_CAPACITY;
    if (!shared->win_c && n > 1) return BPG_ERR_ARG;
  } else if (g_off + n > G->n || h_off + n > H->n) {
    return BPG_ERR_CAPACITY;
  }
  bpg_ipp* st = new (std::nothrow) bpg_ipp();
  if (!st) return BPG_ERR_NOMEM;
  memset(st, 0, sizeof *st);
  st->ctx = ctx;
  st->n = st->m = n;
  st->lanes = lanes;
  int rc = BPG_OK;
  size_t T = 2 * n + 2;
  size_t nparts = 256;
  size_t off = 0;
  const size_t NL = (size_t)lanes;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  size_t o_a = take(NL * n * 32), o_b = take(NL * n * 32), o_wG = take(n * 32), o_wH = take(n * 32);
  size_t o_sc = take(NL * T * 32), o_pid = take(NL * T * 4), o_set = take(NL * T), o_part = take(nparts * 64);
  size_t o_u = take(64), o_ext = take(std::max<size_t>(4, 2 * NL) * 128), o_bytes = take(64 * NL), o_q = take(32),
         o_qm = take(32);
  size_t o_qside = take(64), o_qcomb = take((size_t)COMB_ENTRIES * 96), o_cext = take(NL * 64);
  do {
    cudaError_t e = dev_alloc(ctx, &st->buf, off);
    if (e != cudaSuccess) { ctx->last_cuda = (int)e; rc = BPG_ERR_NOMEM; break; }
    st->a = (uint32_t*)(st->buf + o_a); st->b = (uint32_t*)(st->buf + o_b);
    st->wG = (uint32_t*)(st->buf + o_wG); st->wH = (uint32_t*)(st->buf + o_wH);
    st->scalars = (uint32_t*)(st->buf + o_sc); st->point_ids = (uint32_t*)(st->buf + o_pid);
    st->set_ids = st->buf + o_set; st->partials = (uint32_t*)(st->buf + o_part);
    st->u_pair = (uint32_t*)(st->buf + o_u); st->out_ext = (uint32_t*)(st->buf + o_ext);
    st->out_bytes = st->buf + o_bytes;
    st->q_mul = (uint32_t*)(st->buf + o_qm);
    st->q_side = (uint32_t*)(st->buf + o_qside);
    st->q_comb = (uint32_t*)(st->buf + o_qcomb);
    st->c_ext = (uint32_t*)(st->buf + o_cext);
    uint8_t* d_q = st->buf + o_q;
    cudaStream_t s = ctx->stream;
    if (cudaMemcpyAsync(st->a, d_a, NL * n * 32, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(st->b, d_b, NL * n * 32, cudaMemcpyDeviceToDevice, s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
    unsigned gn = (unsigned)((n + 255) / 256);
    k_ipp_init_weights<<<gn, 256, 0, s>>>(d_gf, d_hf, (uint32_t)n, st->wG, st->wH);
    ctx->launches++;
    if (shared) {
      st->tab = shared;
      st->has_qmul = q_mul_host != nullptr;
      if (q_mul_host) {
        memcpy(ctx->h_pinned + 512, q_mul_host, 32);
        if (cudaMemcpyAsync(st->q_mul, ctx->h_pinned + 512, 32, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
      }
      k_ipp_point_ids<<<dim3(gn, (unsigned)lanes), 256, 0, s>>>(st->point_ids, (uint32_t)n, (uint32_t)g_base, (uint32_t)h_base,
                                                                (uint32_t)q_id);
      ctx->launches++;
      if (cudaStreamSynchronize(s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
    } else if (G == H && G->win_c && n > 1) {
      // The caller's generators already live in ONE windowed table (a resident BulletproofGens):
      // run the round MSMs over it as they are and form c_L Q, c_R Q from a fixed-base comb of Q
      // built here once, on the auxiliary stream beside each round's MSM.
      st->tab = G;
      st->q_sep = true;
      k_ipp_point_ids<<<gn, 256, 0, s>>>(st->point_ids, (uint32_t)n, (uint32_t)g_off, (uint32_t)h_off, (uint32_t)g_off);
      ctx->launches++;
      const size_t qc_bytes = (size_t)COMB_ENTRIES * COMB_AFFINE_WORDS * 4;
      if (ctx->q_cache_valid && memcmp(ctx->q_cache_key, Q_host, 32) == 0) {
        // the same Q as the last call: its comb is resident
        if (cudaMemcpyAsync(st->q_comb, ctx->q_cache_comb, qc_bytes, cudaMemcpyDeviceToDevice, s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
      } else {
        memcpy(ctx->h_pinned + 512, Q_host, 32);
        uint32_t* bad = reinterpret_cast<uint32_t*>(ctx->d_small);
        if (cudaMemsetAsync(bad, 0, 4, s) != cudaSuccess ||
            cudaMemcpyAsync(d_q, ctx->h_pinned + 512, 32, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
        launch_comb_build(ctx, s, d_q, st->q_comb, bad);
        uint32_t* hbad = reinterpret_cast<uint32_t*>(ctx->h_pinned);
        if (cudaMemcpyAsync(hbad, bad, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
        if (*hbad) { rc = BPG_ERR_DECODE; break; }
        ctx->q_cache_valid = false;
        if (!ctx->q_cache_comb && cudaMalloc(&ctx->q_cache_comb, qc_bytes) != cudaSuccess) {
          ctx->q_cache_comb = nullptr;
          cudaGetLastError();  // no cache: not an error
        }
        if (ctx->q_cache_comb &&
            cudaMemcpyAsync(ctx->q_cache_comb, st->q_comb, qc_bytes, cudaMemcpyDeviceToDevice, s) == cudaSuccess) {
          memcpy(ctx->q_cache_key, Q_host, 32);
          ctx->q_cache_valid = true;
        }
      }
    } else {
      k_ipp_point_ids<<<gn, 256, 0, s>>>(st->point_ids, (uint32_t)n, 0u, (uint32_t)n, (uint32_t)(2 * n));
      ctx->launches++;
      // combined table [G | H | Q]
      rc = table_alloc_plain(ctx, 2 * n + 1, &st->own_tab);
      if (rc) break;
      st->tab = st->own_tab;
      if (cudaMemcpyAsync(st->own_tab->niels, G->niels + g_off * NIELS_WORDS, n * NIELS_BYTES, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
          cudaMemcpyAsync(st->own_tab->niels + n * NIELS_WORDS, H->niels + h_off * NIELS_WORDS, n * NIELS_BYTES, cudaMemcpyDeviceToDevice, s) !=
              cudaSuccess) { rc = BPG_ERR_CUDA; break; }
      memcpy(ctx->h_pinned + 512, Q_host, 32);
      uint32_t* bad = reinterpret_cast<uint32_t*>(ctx->d_small);
      if (cudaMemsetAsync(bad, 0, 4, s) != cudaSuccess ||
          cudaMemcpyAsync(d_q, ctx->h_pinned + 512, 32, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
      launch_decode_to_niels(ctx, s, d_q, 1, st->own_tab->niels + 2 * n * NIELS_WORDS, bad);
      uint32_t* hbad = reinterpret_cast<uint32_t*>(ctx->h_pinned);
      if (cudaMemcpyAsync(hbad, bad, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
          cudaStreamSynchronize(s) != cudaSuccess) { rc = BPG_ERR_CUDA; break; }
      if (*hbad) { rc = BPG_ERR_DECODE; break; }
      if (n > 1) {
        rc = bpg_table_set_windows(ctx, st->own_tab, pick_window(n + 1, ctx->forced_c));
        if (rc) break;
      }
    }
  } while (0);
  if (rc != BPG_OK) {
    if (st->own_tab) bpg_table_free(st->own_tab);
    dev_free(ctx, st->buf);
    delete st;
    return rc;
  }
  // Round strategy.  A table that carries combs (bpg_table_build_comb; bpg_gens_new builds them) lets short
  // vectors run every round on the combs, and long ones switch to combs of the folded generators.
  st->n_eff = n;
  if (st->tab->comb && !st->own_tab && n > 1) {
    // measured on a B200 (profiles/r2_ipp_strategy_tuning.json): every round on the generators' combs up to 4096
    // entries; beyond, the folded generators are formed at 1024 entries (2048 from 2^16 on, where the bucket
    // rounds before it are the larger share)
    const size_t direct_max = std::min<size_t>(env_size("BPG_IPP_DIRECT_MAX", 4096), 65536);
    const size_t m0 = env_size("BPG_IPP_M0", n >= 65536 ? 2048 : 1024);
    st->cq_comb = st->q_sep ? st->q_comb : st->tab->comb + (size_t)q_id * COMB_ENTRIES * COMB_AFFINE_WORDS;
    if (n <= direct_max || n <= m0) {
      st->mode = 1;
      st->comb_affine = true;
      st->comb = st->tab->comb;
      st->cg_id = (uint32_t)(shared ? g_base : g_off);
      st->ch_id = (uint32_t)(shared ? h_base : h_off);
    } else if (m0 >= 2 && (m0 & (m0 - 1)) == 0) {
      st->m0 = m0;
      st->cg_id = (uint32_t)(shared ? g_base : g_off);
      st->ch_id = (uint32_t)(shared ? h_base : h_off);
    }
  }
  *out = st;
  return BPG_OK;
}

// comb rounds ----------------------------------------------------------------------------------
static int ipp_ensure_parts(bpg_ipp* st, size_t bytes) {
  if (bytes <= st->parts_cap) return BPG_OK;
  bpg_ctx* ctx = st->ctx;
  if (st->parts) dev_free(ctx, st->parts);
  st->parts = nullptr;
  st->parts_cap = 0;
  if (dev_alloc(ctx, &st->parts, bytes) != cudaSuccess) return BPG_ERR_NOMEM;
  st->parts_cap = bytes;
  return BPG_OK;
}

// The vectors have shrunk to m0: form the folded generators once (k_comb_materialize over the generator combs),
// give them combs of their own, and restart the weights at one.  From here on a round touches 2 m0 points.
static int ipp_materialize(bpg_ipp* st) {
  bpg_ctx* ctx = st->ctx;
  cudaStream_t s = ctx->stream;
  const size_t m0 = st->m0, npts = 2 * m0;
  const size_t w_pts = npts * 32 * COMB_MAT_SPLIT, w_chain = npts * COMB_WINDOWS * 32, w_comb = npts * (size_t)COMB_ENTRIES * COMB_CACHED_WORDS;
  if (dev_alloc(ctx, &st->mat_buf, (w_pts + w_chain + w_comb) * 4) != cudaSuccess) return BPG_ERR_NOMEM;
  uint32_t *folded = st->mat_buf, *chain = folded + w_pts, *comb = chain + w_chain;
  prof_mark(ctx, BPG_PROF_COMBINE);
  int rc = comb_materialize(ctx, s, st->tab->comb, st->cg_id, st->ch_id, st->wG, st->wH, st->n, m0, folded, chain, comb);
  if (rc) return rc;
  k_ipp_init_weights<<<(unsigned)((m0 + 255) / 256), 256, 0, s>>>(nullptr, nullptr, (uint32_t)m0, st->wG, st->wH);
  LAUNCH_CHECK();
  prof_mark(ctx, -1);
  st->mode = 1;
  st->comb_affine = false;
  st->comb = comb;
  st->cg_id = 0;
  st->ch_id = (uint32_t)m0;
  st->n_eff = m0;
  st->cross_ready = false;
  return BPG_OK;
}

// one comb round: accumulation + finish; the encodings of the 2 x lanes sums land in st->out_bytes
static int ipp_comb_round(bpg_ipp* st, bool external_cross, uint8_t* out_bytes /*device-visible: where the encodings go*/) {
  bpg_ctx* ctx = st->ctx;
  cudaStream_t s = ctx->stream;
  const size_t n = st->n_eff, m = st->m, h = m / 2;
  const unsigned sets = 2u * (unsigned)st->lanes;
  prof_mark(ctx, BPG_PROF_OTHER);
  // cross terms: the caller's shares, the fold kernel's partials, or (single prover) formed by the finish itself
  const bool cross_in_finish = !external_cross && !st->cross_ready && st->lanes == 1;
  if (!external_cross && !st->cross_ready && !cross_in_finish) {
    unsigned gcross = (unsigned)std::min<size_t>(256, (h + IPP_THREADS - 1) / IPP_THREADS);
    k_ipp_cross<<<gcross, IPP_THREADS, 0, s>>>(st->a, st->b, (uint32_t)h, st->partials);
    LAUNCH_CHECK();
    st->ncross = gcross;
  }
  if (st->fold_pending && !st->alt) {
    if (dev_alloc(ctx, &st->alt, 4 * n * 32) != cudaSuccess) return BPG_ERR_NOMEM;
    st->alt_n = n;
    for (int k = 0; k < 4; k++) st->oth[k] = st->alt + (size_t)k * n * 8;
  }
  if (st->fold_pending && st->alt_n < n) return BPG_ERR_ARG;  // n_eff only shrinks after the first comb round
  // threads per term: enough (term, window-slice) units to fill the machine, at most one window per thread
  // (each thread's accumulator must then go through the block tree, which costs about three additions of its own:
  // more slices than the machine needs only add tree work)
  static const size_t target_mul = env_size("BPG_COMB_THREADS_PER_SM", 192);
  uint32_t wsplit = 1;
  while (wsplit < COMB_WINDOWS && n * wsplit * sets < (size_t)ctx->sm_count * target_mul) wsplit <<= 1;
  const unsigned bx = (unsigned)((n * wsplit + CB_THREADS - 1) / CB_THREADS);
  int rc = ipp_ensure_parts(st, (size_t)sets * bx * 128);
  if (rc) return rc;
  CombRound R;
  R.comb = st->comb;
  R.g_id = st->cg_id;
  R.h_id = st->ch_id;
  R.a = st->a;
  R.b = st->b;
  R.wG = st->wG;
  R.wH = st->wH;
  R.stride = (uint32_t)st->n;
  R.n = (uint32_t)n;
  R.m = (uint32_t)m;
  R.wsplit = wsplit;
  R.bias4 = bias_for(4);
  R.fold = st->fold_pending ? 1u : 0u;
  R.up = st->fold_up;
  R.a2 = R.b2 = R.wG2 = R.wH2 = nullptr;
  if (st->fold_pending) {  // lanes == 1 here
    R.a2 = st->oth[0];
    R.b2 = st->oth[1];
    R.wG2 = st->oth[2];
    R.wH2 = st->oth[3];
  }
  prof_mark(ctx, BPG_PROF_ACCUM);
  if (st->comb_affine) k_comb_round<true><<<dim3(bx, sets), CB_THREADS, 0, s>>>(R, st->parts);
  else k_comb_round<false><<<dim3(bx, sets), CB_THREADS, 0, s>>>(R, st->parts);
  LAUNCH_CHECK();
  if (st->fold_pending) {
    // the folded state is the current one from here on; the old vectors become the next round's alternates
    st->fold_pending = false;
    uint32_t* cur[4] = {st->a, st->b, st->wG, st->wH};
    st->a = st->oth[0], st->b = st->oth[1], st->wG = st->oth[2], st->wH = st->oth[3];
    for (int k = 0; k < 4; k++) st->oth[k] = cur[k];
  }
  CombFinal F;
  F.parts = st->parts;
  F.nparts = bx;
  F.va = cross_in_finish ? st->a : nullptr;
  F.vb = cross_in_finish ? st->b : nullptr;
  F.vh = (uint32_t)h;
  F.cross = (external_cross || cross_in_finish) ? nullptr : st->partials;
  F.ncross = st->ncross;
  F.c_ext = external_cross ? st->c_ext : nullptr;
  F.q_mul = st->has_qmul ? st->q_mul : nullptr;
  F.q_comb = st->cq_comb;
  F.bias4 = bias_for(4);
  prof_mark(ctx, BPG_PROF_ENCODE);
  k_comb_final<true><<<sets, CBQ_THREADS, 0, s>>>(F, out_bytes, nullptr);
  LAUNCH_CHECK();
  prof_mark(ctx, -1);
  st->cross_ready = false;
  return BPG_OK;
}

extern "C" int bpg_ipp_begin(bpg_ctx* ctx, const bpg_table* G, size_t g_off, const bpg_table* H, size_t h_off,
                             size_t n, const uint8_t Q[32], const uint8_t* G_factors, const uint8_t* H_factors,
                             const uint8_t* a, const uint8_t* b, bpg_ipp** out) {
  if (!ctx || !G || !H || !Q || !a || !b || !out) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_stage(ctx, 4 * n * 32 + 64);
  if (rc) return rc;
  uint8_t* d = ctx->d_stage;
  CK(cudaMemcpyAsync(d, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d + n * 32, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (G_factors) CK(cudaMemcpyAsync(d + 2 * n * 32, G_factors, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (H_factors) CK(cudaMemcpyAsync(d + 3 * n * 32, H_factors, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  return ipp_begin_dev(ctx, G, g_off, H, h_off, n, Q, nullptr, 0, 0, 0, nullptr,
                       G_factors ? (const uint32_t*)(d + 2 * n * 32) : nullptr,
                       H_factors ? (const uint32_t*)(d + 3 * n * 32) : nullptr, (const uint32_t*)d,
                       (const uint32_t*)(d + n * 32), out);
}

extern "C" int bpg_ipp_begin_dev(bpg_ctx* ctx, const bpg_table* G, size_t g_off, const bpg_table* H, size_t h_off,
                                 size_t n, const uint8_t Q[32], const void* d_G_factors, const void* d_H_factors,
                                 const void* d_a, const void* d_b, bpg_ipp** out) {
  if (!ctx || !G || !H || !Q || !d_a || !d_b || !out) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  return ipp_begin_dev(ctx, G, g_off, H, h_off, n, Q, nullptr, 0, 0, 0, nullptr, (const uint32_t*)d_G_factors,
                       (const uint32_t*)d_H_factors, (const uint32_t*)d_a, (const uint32_t*)d_b, out);
}

// Generators and the base of Q live in one windowed table (the R1CS prover: G at g_base, H at
// h_base, Q = q_mul * shared[q_id] with q_id the Pedersen base B and q_mul the challenge w).
extern "C" int bpg_ipp_begin_shared(bpg_ctx* ctx, const bpg_table* shared, size_t g_base, size_t h_base, size_t q_id,
                                    const uint8_t q_mul[32], size_t n, const uint8_t* G_factors,
                                    const uint8_t* H_factors, const uint8_t* a, const uint8_t* b, bpg_ipp** out) {
  if (!ctx || !shared || !a || !b || !out) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_stage(ctx, 4 * n * 32 + 64);
  if (rc) return rc;
  uint8_t* d = ctx->d_stage;
  CK(cudaMemcpyAsync(d, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d + n * 32, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (G_factors) CK(cudaMemcpyAsync(d + 2 * n * 32, G_factors, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (H_factors) CK(cudaMemcpyAsync(d + 3 * n * 32, H_factors, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  return ipp_begin_dev(ctx, nullptr, 0, nullptr, 0, n, nullptr, shared, g_base, h_base, q_id, q_mul,
                       G_factors ? (const uint32_t*)(d + 2 * n * 32) : nullptr,
                       H_factors ? (const uint32_t*)(d + 3 * n * 32) : nullptr, (const uint32_t*)d,
                       (const uint32_t*)(d + n * 32), out);
}

extern "C" size_t bpg_ipp_rounds_left(const bpg_ipp* st) {
  size_t r = 0;
  if (!st) return 0;
  for (size_t m = st->m; m > 1; m >>= 1) r++;
  return r;
}

extern "C" int bpg_ipp_round_LR(bpg_ipp* st, uint8_t L[32], uint8_t R[32]) {
  if (!st || !L || !R) return BPG_ERR_ARG;
  if (st->m <= 1 || st->lr_done || st->lanes != 1) return BPG_ERR_ARG;
  bpg_ctx* ctx = st->ctx;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  size_t n = st->n, m = st->m, h = m / 2;
  if (st->mode == 0 && st->m0 && m <= st->m0) {
    int rc = ipp_materialize(st);
    if (rc) return rc;
  }
  // The two encodings are written by the finishing kernel straight into page-locked host memory (it is mapped into
  // the device's address space): the round trip is the kernel and one stream wait, no copy-engine hop in between.
  uint8_t* host_lr = ctx->h_pinned + 1024;
  if (st->mode == 1) {
    int rc = ipp_comb_round(st, false, host_lr);
    if (rc) return rc;
    CK(cudaStreamSynchronize(s));
    memcpy(L, host_lr, 32);
    memcpy(R, host_lr + 32, 32);
    st->lr_done = true;
    return BPG_OK;
  }
  unsigned gcross = (unsigned)std::min<size_t>(256, (h + IPP_THREADS - 1) / IPP_THREADS);
  prof_mark(ctx, BPG_PROF_OTHER);
  k_ipp_cross<<<gcross, IPP_THREADS, 0, s>>>(st->a, st->b, (uint32_t)h, st->partials);
  LAUNCH_CHECK();
  k_ipp_round_scalars<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(st->a, st->b, st->wG, st->wH, (uint32_t)n,
                                                                  (uint32_t)m, st->scalars, st->set_ids);
  LAUNCH_CHECK();
  k_ipp_cross_finish<<<1, IPP_THREADS, 0, s>>>(st->partials, gcross, (uint32_t)n, st->has_qmul ? st->q_mul : nullptr,
                                               st->scalars, st->set_ids, st->q_sep ? st->q_side : nullptr);
  LAUNCH_CHECK();
  if (st->q_sep) {
    // c_L Q, c_R Q: 64 mixed additions each from the comb of Q, beside the MSM
    CK(cudaEventRecord(ctx->ev_fork, s));
    CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    launch_comb_mul(ctx, ctx->aux_stream, st->q_comb, 1, st->q_side, 2, nullptr, st->out_ext + 64);
    CK(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
  }
  int rc = msm_enqueue(ctx, st->tab->niels, st->tab->n, st->scalars, 2 * n + 2, st->set_ids, st->point_ids, 2,
                       st->out_ext, st->tab->win_c, st->tab->n);
  if (rc) return rc;
  if (st->q_sep) CK(cudaStreamWaitEvent(s, ctx->ev_join, 0));
  rc = bpg_dev_sum_encode(ctx, st->out_ext, st->q_sep ? 2 : 1, 2, host_lr, nullptr);
  if (rc) return rc;
  CK(cudaStreamSynchronize(s));
  memcpy(L, host_lr, 32);
  memcpy(R, host_lr + 32, 32);
  st->lr_done = true;
  return BPG_OK;
}

extern "C" int bpg_ipp_round_fold(bpg_ipp* st, const uint8_t u[32], const uint8_t u_inv[32]) {
  if (!st || !u || !u_inv) return BPG_ERR_ARG;
  if (st->m <= 1 || !st->lr_done) return BPG_ERR_ARG;
  bpg_ctx* ctx = st->ctx;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  // the challenge pair travels as kernel arguments: no staging copy, and no wait here -- the
  // next round's launches queue up behind the fold
  ScPair up;
  memcpy(up.v, u, 32);
  memcpy(up.v + 8, u_inv, 32);
  prof_mark(ctx, BPG_PROF_OTHER);
  if (st->mode == 1 && st->lanes == 1 && !st->shares && st->m > 2 && !st->fold_pending) {
    // single prover, another round follows: that round's accumulation takes this fold on the fly (CombRound::fold)
    memcpy(st->fold_up.v, up.v, sizeof st->fold_up.v);
    st->fold_pending = true;
    st->cross_ready = false;
  } else if (st->mode == 1) {
    // fold + the next round's cross-term partials in one kernel (the shares path gets them from the caller)
    IppPair ip;
    memcpy(ip.v, up.v, sizeof ip.v);
    const size_t cover = std::max<size_t>(st->n_eff, 1);
    const unsigned gx = (unsigned)((cover + IFC_THREADS - 1) / IFC_THREADS);
    const bool want_cross = st->lanes == 1 && st->m >= 4;
    k_ipp_fold_cross<<<dim3(gx, (unsigned)st->lanes), IFC_THREADS, 0, s>>>(st->a, st->b, st->wG, st->wH, (uint32_t)st->n_eff,
                                                                            (uint32_t)st->m, (uint32_t)st->n, ip,
                                                                            want_cross ? st->partials : nullptr);
    LAUNCH_CHECK();
    st->cross_ready = want_cross;
    st->ncross = gx;
  } else {
    k_ipp_fold<<<dim3((unsigned)((st->n + 255) / 256), (unsigned)st->lanes), 256, 0, s>>>(st->a, st->b, st->wG, st->wH,
                                                                                         (uint32_t)st->n, (uint32_t)st->m, up);
    LAUNCH_CHECK();
  }
  prof_mark(ctx, -1);
  st->m /= 2;
  st->lr_done = false;
  return BPG_OK;
}

// ---------------------------------------------------------------------------
// the same rounds on secret shares (r1cs_mpc): SharedInnerProductProof::create, reference
// src/r1cs_mpc/mpc_inner_product.rs:52-228.  A party holds additive shares of a and b -- `lanes` (a, b) pairs:
// the value shares and, in an authenticated fabric, the MAC shares -- while G, H, the factors and Q are
// public.  Everything linear is local: the party's shares of L and R are MSMs of its share vectors
// (:104-126, 172-186) and the folds use the public challenge (:136-137, 196-197).  The cross terms
// c_L = <a_lo, b_hi>, c_R = <a_hi, b_lo> are products of shared values, i.e. the fabric's multiplication
// protocol over the network: the caller reads the current vectors (bpg_ipp_read_ab), runs that protocol,
// and hands this party's shares of c_L, c_R to bpg_ipp_round_LR_shares.
// ---------------------------------------------------------------------------
extern "C" int bpg_ipp_begin_shares(bpg_ctx* ctx, const bpg_table* shared, size_t g_base, size_t h_base, size_t q_id,
                                    const uint8_t q_mul[32], size_t n, int lanes, const uint8_t* G_factors,
                                    const uint8_t* H_factors, const uint8_t* a, const uint8_t* b, bpg_ipp** out) {
  if (!ctx || !shared || !a || !b || !out || lanes < 1 || lanes > IPP_MAX_LANES) return BPG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const size_t NL = (size_t)lanes;
  int rc = ensure_stage(ctx, (2 * NL + 2) * n * 32 + 64);
  if (rc) return rc;
  uint8_t* d = ctx->d_stage;
  uint8_t *d_a = d, *d_b = d + NL * n * 32, *d_gf = d + 2 * NL * n * 32, *d_hf = d_gf + n * 32;
  CK(cudaMemcpyAsync(d_a, a, NL * n * 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_b, b, NL * n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (G_factors) CK(cudaMemcpyAsync(d_gf, G_factors, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (H_factors) CK(cudaMemcpyAsync(d_hf, H_factors, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  rc = ipp_begin_dev(ctx, nullptr, 0, nullptr, 0, n, nullptr, shared, g_base, h_base, q_id, q_mul,
                     G_factors ? (const uint32_t*)d_gf : nullptr, H_factors ? (const uint32_t*)d_hf : nullptr,
                     (const uint32_t*)d_a, (const uint32_t*)d_b, out, lanes);
  if (rc == BPG_OK) (*out)->shares = true;  // the caller reads the folded vectors between rounds: folds stay eager
  return rc;
}
extern "C" int bpg_ipp_lanes(const bpg_ipp* st) { return st ? st->lanes : 0; }
extern "C" size_t bpg_ipp_len(const bpg_ipp* st) { return st ? st->m : 0; }
// current vectors, lane-major: a_out, b_out = lanes x m x 32 bytes (m = bpg_ipp_len)
extern "C" int bpg_ipp_read_ab(bpg_ipp* st, uint8_t* a_out, uint8_t* b_out) {
  if (!st || !a_out || !b_out) return BPG_ERR_ARG;
  bpg_ctx* ctx = st->ctx;
  CK(cudaSetDevice(ctx->device));
  for (int l = 0; l < st->lanes; l++) {
    CK(cudaMemcpyAsync(a_out + (size_t)l * st->m * 32, st->a + (size_t)l * st->n * 8, st->m * 32, cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaMemcpyAsync(b_out + (size_t)l * st->m * 32, st->b + (size_t)l * st->n * 8, st->m * 32, cudaMemcpyDeviceToHost,
                       ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return BPG_OK;
}
// c_L, c_R: lanes x 32 bytes each, this party's shares of the cross terms; L_out, R_out: lanes x 32 bytes,
// the compressed encodings of this party's shares of L and R (per lane)
extern "C" int bpg_ipp_round_LR_shares(bpg_ipp* st, const uint8_t* c_L, const uint8_t* c_R, uint8_t* L_out,
                                       uint8_t* R_out) {
  if (!st || !c_L || !c_R || !L_out || !R_out) return BPG_ERR_ARG;
  if (st->m <= 1 || st->lr_done || st->q_sep || st->own_tab) return BPG_ERR_ARG;
  bpg_ctx* ctx = st->ctx;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t n = st->n, m = st->m;
  const int NL = st->lanes;
  uint8_t* hp = ctx->h_pinned + 4096;
  for (int l = 0; l < NL; l++) {
    memcpy(hp + 64 * l, c_L + 32 * l, 32);
    memcpy(hp + 64 * l + 32, c_R + 32 * l, 32);
  }
  CK(cudaMemcpyAsync(st->c_ext, hp, 64 * (size_t)NL, cudaMemcpyHostToDevice, s));
  if (st->mode == 0 && st->m0 && m <= st->m0) {
    int rc = ipp_materialize(st);
    if (rc) return rc;
  }
  int rc;
  if (st->mode == 1) {
    rc = ipp_comb_round(st, true, st->out_bytes);
    if (rc) return rc;
  } else {
    prof_mark(ctx, BPG_PROF_OTHER);
    k_ipp_round_scalars<<<dim3((unsigned)((n + 255) / 256), (unsigned)NL), 256, 0, s>>>(st->a, st->b, st->wG, st->wH, (uint32_t)n,
                                                                                       (uint32_t)m, st->scalars, st->set_ids);
    LAUNCH_CHECK();
    k_ipp_q_terms_ext<<<1, 32, 0, s>>>(st->c_ext, st->has_qmul ? st->q_mul : nullptr, (uint32_t)n, (uint32_t)NL, st->scalars,
                                       st->set_ids);
    LAUNCH_CHECK();
    rc = msm_enqueue(ctx, st->tab->niels, st->tab->n, st->scalars, (size_t)NL * (2 * n + 2), st->set_ids, st->point_ids,
                     2 * NL, st->out_ext, st->tab->win_c, st->tab->n);
    if (rc) return rc;
    rc = bpg_dev_sum_encode(ctx, st->out_ext, 1, 2 * NL, st->out_bytes, nullptr);
    if (rc) return rc;
  }
  CK(cudaMemcpyAsync(ctx->h_pinned, st->out_bytes, 64 * (size_t)NL, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (int l = 0; l < NL; l++) {
    memcpy(L_out + 32 * l, ctx->h_pinned + 64 * l, 32);
    memcpy(R_out + 32 * l, ctx->h_pinned + 64 * l + 32, 32);
  }
  st->lr_done = true;
  return BPG_OK;
}
// final a, b of every lane (lanes x 32 bytes each): this party's shares (:217-228)
extern "C" int bpg_ipp_finish_shares(bpg_ipp* st, uint8_t* a, uint8_t* b) {
  if (!st || !a || !b) return BPG_ERR_ARG;
  if (st->m != 1) return BPG_ERR_ARG;
  return bpg_ipp_read_ab(st, a, b);
}

extern "C" int bpg_ipp_finish(bpg_ipp* st, uint8_t a[32], uint8_t b[32]) {
  if (!st || !a || !b) return BPG_ERR_ARG;
  if (st->m != 1) return BPG_ERR_ARG;
  bpg_ctx* ctx = st->ctx;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(ctx->h_pinned, st->a, 32, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_pinned + 32, st->b, 32, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  memcpy(a, ctx->h_pinned, 32);
  memcpy(b, ctx->h_pinned + 32, 32);
  return BPG_OK;
}

extern "C" void bpg_ipp_free(bpg_ipp* st) {
  if (!st) return;
  cudaSetDevice(st->ctx->device);
  if (st->q_sep) cudaStreamSynchronize(st->ctx->aux_stream);
  if (st->own_tab) bpg_table_free(st->own_tab);
  dev_free(st->ctx, st->mat_buf);
  dev_free(st->ctx, st->parts);
  dev_free(st->ctx, st->alt);
  dev_free(st->ctx, st->buf);
  delete st;
}

