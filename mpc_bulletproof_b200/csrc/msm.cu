// libbpgpu: one Pippenger launch (sort, schedule, bucket accumulation, reduction) for both curve policies.
#include "msm_launch.cuh"

using namespace bpg;

cudaError_t msm_sort_kernels_init() {
  cudaError_t e = cudaFuncSetAttribute(k_rs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SCATTER_SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_rs_finish, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)(((size_t)4 << RS_MAX_LB) + (size_t)RS_FINISH_CAP * 4));
}


// ---------------------------------------------------------------------------
// MSM launch
// ---------------------------------------------------------------------------
// Plain tables: every window has its own bucket array, reduced separately, then Horner.
int pick_window(size_t n_per_set, int forced) {
  if (forced >= 2) return forced;
  double best = 1e300;
  int best_c = 4;
  for (int c = 3; c <= 20; c++) {
    int W = (255 + c - 1) / c;
    double nb = (double)(1u << (c - 1));
    // mixed adds (7M) for the terms, ~20M per bucket in the reduction tree
    double cost = W * ((double)n_per_set * 7.0 + nb * 20.0);
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}
// Windowed tables: all windows share one bucket array per set.  Lists of ~32 entries keep the
// accumulation efficient; shorter lists only add merge work.
static uint32_t pick_gsub(size_t n_per_set, int nsets, int c, int W) {
  double nb = (double)(1u << (c - 1));
  double lists_for_len = (double)n_per_set * W / (nb * 32.0);       // groups that make the lists ~32 long
  double lists_for_par = (double)(1u << 17) / (nb * (double)nsets);  // groups that give ~2^17 lists (more only add merge work)
  double avg1 = (double)n_per_set * W / nb;                          // list length with one group
  double g = std::max(lists_for_len, std::min(lists_for_par, avg1 / 8.0));  // never below ~8 entries per list
  uint32_t gs = (uint32_t)(g + 0.5);
  return std::min<uint32_t>(std::max<uint32_t>(gs, 1), (uint32_t)W);
}
int pick_window_table(size_t n, int forced) {
  if (forced >= 2) return forced;
  double best = 1e300;
  int best_c = 4;
  for (int c = 3; c <= 20; c++) {
    int W = (255 + c - 1) / c;
    double nb = (double)(1u << (c - 1));
    uint32_t gs = pick_gsub(n, 1, c, W);
    double cost = (double)W * n * 7.0 + (gs > 1 ? gs * nb * 8.0 : 0.0) + nb * 20.0;
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

static void make_cfg(MsmCfg& cfg, size_t n_terms, size_t n_points, int nsets, int c, size_t win_stride, int forced_gsub) {
  cfg.win_stride = (uint32_t)win_stride;
  cfg.c = c;
  cfg.W = (255 + c - 1) / c;
  cfg.nb = 1u << (c - 1);
  cfg.nsets = nsets;
  cfg.n_terms = (uint32_t)n_terms;
  cfg.n_points = (uint32_t)std::max<size_t>(n_points, 1);
  if (win_stride) {
    cfg.gsub = forced_gsub > 0 ? std::min<uint32_t>((uint32_t)forced_gsub, (uint32_t)cfg.W)
                               : pick_gsub((n_terms + nsets - 1) / nsets, nsets, c, cfg.W);
  } else {
    cfg.gsub = (uint32_t)cfg.W;
  }
  cfg.narr = (uint32_t)nsets * cfg.gsub;
  cfg.B = cfg.narr * cfg.nb;
  // segments of over-long buckets (> BIG_SEG entries): at most two per BIG_SEG entries
  cfg.big_cap = (uint32_t)(2 * ((uint64_t)n_terms * cfg.W / BIG_SEG) + 2);
  memset(&cfg.bias, 0, sizeof(cfg.bias));
  for (int w = 0; w < cfg.W; w++) {
    int bit = c * w + c - 1;
    cfg.bias.v[bit >> 5] |= 1u << (bit & 31);
  }
}


// Enqueue one Pippenger launch.  d_scalars: n_terms*32 B; d_set_ids / d_point_ids may
// be null (implicit: term t -> point t % n_points of `table_base`, set t / n_points).
int msm_enqueue(bpg_ctx* ctx, const uint32_t* table_base, size_t n_points, const uint32_t* d_scalars,
                size_t n_terms, const uint8_t* d_set_ids, const uint32_t* d_point_ids, int nsets,
                uint32_t* d_out_ext, int win_c, size_t win_stride, int lane, int curve,
                const uint8_t* h_scalars /*scalars still on the host: uploaded here, in pieces*/) {
  // curve 0: ristretto255 (Niels table, 24 words per entry); curve 1: Stark curve (affine table, 16 words
  // per entry, plain tables only).  Sort and schedule are shared; the bucket arithmetic differs.
  if (nsets <= 0) return BPG_ERR_ARG;
  // lane 1: the auxiliary stream and arena (no phase profiling there)
  struct ProfOff {
    bpg_ctx* c;
    bool saved;
    ProfOff(bpg_ctx* c_, bool off) : c(c_), saved(c_->prof) { if (off) c->prof = false; }
    ~ProfOff() { c->prof = saved; }
  } prof_off(ctx, lane != 0);
  cudaStream_t st = lane ? ctx->aux_stream : ctx->stream;
  uint8_t* const& ws = lane ? ctx->ws_aux : ctx->ws;
  if (n_terms == 0) return msm_identity(ctx, st, curve, nsets, d_out_ext);  // empty sum: identity for every set
  if (n_terms >= (1u << 31)) return BPG_ERR_ARG;
  if (curve == 0 && win_c == 0 && msm_small_applies(n_terms, nsets) && ctx->forced_c < 2) {
    // a handful of terms over a plain table: one quad per term walks the doubling chain (k_msm_small);
    // a forced window width (bpg_set_window) keeps the bucket pipeline, which is how the tests reach it
    if (h_scalars) CK(cudaMemcpyAsync((void*)d_scalars, h_scalars, n_terms * 32, cudaMemcpyHostToDevice, st));
    return msm_small_ristretto(ctx, st, nullptr, lane, table_base, n_points, d_scalars, n_terms, d_set_ids, d_point_ids,
                               nsets, d_out_ext);
  }
  MsmLaunch L;
  L.ctx = ctx;
  L.st = st;
  L.lane = lane;
  L.nsets = nsets;
  L.table = table_base;
  L.out_ext = d_out_ext;
  MsmCfg& cfg = L.cfg;
  int c = win_c ? win_c : pick_window((n_terms + nsets - 1) / nsets, ctx->forced_c);
  make_cfg(cfg, n_terms, n_points, nsets, c, win_c ? win_stride : 0, ctx->forced_gsub);
  if ((uint64_t)cfg.narr * cfg.nb >= (1ull << 31)) return BPG_ERR_ARG;
  const bool windowed = cfg.win_stride != 0;
  L.windowed = windowed;

  // reduction geometry: `rarr` arrays of nb buckets, `tiles0` leaf tiles per array
  L.rarr = windowed ? (uint32_t)nsets : cfg.narr;
  if (curve == 1) {
    L.tiles0 = msm_stark_tiles0(cfg);
    L.thread_leaf = false;
    L.LC = 0;
  } else {
    msm_reduce_geometry(cfg, &L.thread_leaf, &L.LC, &L.tiles0);
  }
  size_t ntiles = (cfg.B + SCAN_TILE - 1) / SCAN_TILE;
  // sort: two-pass radix partition in shared memory (default) or the global-atomic counting sort (BPG_SORT=atomic,
  // and bucket spaces beyond 2^25)
  static const bool sort_atomic = getenv("BPG_SORT") && !strcmp(getenv("BPG_SORT"), "atomic");
  static const uint32_t target_parts = getenv("BPG_RS_PARTS") ? (uint32_t)atoi(getenv("BPG_RS_PARTS")) : RS_TARGET_PARTS;
  static const uint32_t fin_threads = getenv("BPG_RS_FIN_THREADS") ? (uint32_t)atoi(getenv("BPG_RS_FIN_THREADS")) : RS_FINISH_THREADS;
  static const uint32_t fin_cap = getenv("BPG_RS_FIN_CAP") ? (uint32_t)atoi(getenv("BPG_RS_FIN_CAP")) : RS_FINISH_CAP;
  // partitions: about target_parts, more (up to RS_MAX_PARTS) when a partition's pairs would not fit the shared-memory
  // placement of k_rs_finish; launches whose partitions cannot be made small enough (2^24 points and beyond) keep
  // the counting sort
  const uint64_t n_pairs = (uint64_t)n_terms * cfg.W;
  uint32_t want_parts = target_parts;
  while (want_parts < RS_MAX_PARTS && n_pairs / want_parts > (fin_cap * 7) / 8) want_parts <<= 1;
  uint32_t lb = 4;
  while (lb < RS_MAX_LB && ((uint64_t)cfg.B >> lb) > want_parts) lb++;
  uint32_t P = (uint32_t)(((uint64_t)cfg.B + (1u << lb) - 1) >> lb);
  const bool radix = !sort_atomic && P <= RS_MAX_PARTS && n_pairs / P <= (uint64_t)fin_cap * 2 &&
                     (size_t)RS_THREADS * cfg.W * 8 + (size_t)3 * P * 4 <= RS_SCATTER_SMEM;
  size_t off = 0;
  // counts / partition counters and the schedule's control words are adjacent: ONE memset per launch
  size_t o_counts = off;  off += align_up(radix ? (size_t)2 * P * 4 : (size_t)cfg.B * 4);
  size_t o_bins = off;    off += align_up((2 * SIZE_BINS + 4) * 4);  // bins | n_items, part, multi, big_count | cursors
  size_t o_offsets = off; off += align_up(((size_t)cfg.B + 1) * 4);
  size_t o_tiles = off;   off += align_up(radix ? ((size_t)P + 1) * 4 : ntiles * 4);
  size_t o_big = off;     off += align_up(3 * (size_t)cfg.big_cap * 4);
  size_t o_bigpart = off; off += align_up((size_t)cfg.big_cap * 128);
  size_t o_entries = off; off += align_up((size_t)n_terms * cfg.W * 4);
  size_t o_pairs = off;   off += radix ? align_up((size_t)n_terms * cfg.W * 8) : 0;
  size_t o_buckets = off; off += align_up((size_t)cfg.B * 128);
  size_t o_merged = off;  off += (windowed && cfg.gsub > 1) ? align_up((size_t)nsets * cfg.nb * 128) : 0;
  size_t o_rpairs = off;  off += 4 * align_up((size_t)L.rarr * L.tiles0 * 128);  // (A, Y) x ping-pong
  size_t o_wins = off;    off += align_up((size_t)L.rarr * 128);
  // accumulation schedule: at most one item per bucket plus one per ACC_SEG entries
  L.max_items = (size_t)cfg.B + (size_t)n_terms * cfg.W / ACC_SEG + 1;
  L.max_multi = (size_t)n_terms * cfg.W / ACC_SEG + 1;  // buckets longer than ACC_SEG
  size_t o_items = off;   off += align_up(L.max_items * 8);
  size_t o_segslot = off; off += align_up((size_t)cfg.B * 4);
  size_t o_multi = off;   off += align_up(L.max_multi * 4);
  size_t o_segpart = off; off += align_up(2 * L.max_multi * 128);  // sum of nseg over multi-segment buckets <= 2 max_multi
  int rc = ensure_ws(ctx, off, lane);
  if (rc) return rc;
  uint32_t* counts = (uint32_t*)(ws + o_counts);
  uint32_t* tiles = (uint32_t*)(ws + o_tiles);
  uint32_t* bins = (uint32_t*)(ws + o_bins);
  L.offsets = (uint32_t*)(ws + o_offsets);
  L.big_list = (uint32_t*)(ws + o_big);
  L.big_part = (uint32_t*)(ws + o_bigpart);
  L.entries = (uint32_t*)(ws + o_entries);
  L.buckets = (uint32_t*)(ws + o_buckets);
  L.merged = (uint32_t*)(ws + o_merged);
  L.pair_words = align_up((size_t)L.rarr * L.tiles0 * 128) / 4;
  L.pairs = (uint32_t*)(ws + o_rpairs);
  L.wins = (uint32_t*)(ws + o_wins);
  AccSched& sched = L.sched;
  L.big_count = bins + SIZE_BINS + 3;
  sched.bins = bins;
  sched.cursors = bins + SIZE_BINS + 4;
  sched.n_items = bins + SIZE_BINS;
  sched.part_count = bins + SIZE_BINS + 1;
  sched.multi_count = bins + SIZE_BINS + 2;
  sched.items = (uint2*)(ws + o_items);
  sched.seg_slot = (uint32_t*)(ws + o_segslot);
  sched.multi_list = (uint32_t*)(ws + o_multi);
  L.seg_part = (uint32_t*)(ws + o_segpart);
  uint32_t* offsets = L.offsets;

  prof_mark(ctx, BPG_PROF_HIST);
  CK(cudaMemsetAsync(counts, 0, o_bins - o_counts + (2 * SIZE_BINS + 4) * 4, st));
  // radix: a tile's pairs are staged in shared memory (RS_STAGE_PAIRS at most); terms per thread accordingly, but
  // no more than leaves about four tiles per SM
  int tpt = 1;
  while (tpt < RS_TERMS && (size_t)RS_THREADS * (2 * tpt) * cfg.W <= RS_STAGE_PAIRS &&
         n_terms / ((size_t)RS_THREADS * tpt) > (size_t)ctx->sm_count * 4)
    tpt <<= 1;
  const size_t tile_terms = (size_t)RS_THREADS * tpt;
  const uint32_t stage_cap = (uint32_t)(tile_terms * cfg.W);
  const size_t scatter_smem = (size_t)stage_cap * 8 + (size_t)3 * P * 4;
  uint32_t *part_count = counts, *part_cursor = counts + P, *part_base = tiles;
  auto hist = [&](size_t t0, size_t t1) -> int {
    if (radix) {
      k_rs_hist<<<(unsigned)((t1 - t0 + tile_terms - 1) / tile_terms), RS_THREADS, P * 4, st>>>(
          d_scalars, d_set_ids, cfg, lb, P, part_count, (uint32_t)t0, (uint32_t)t1, tpt);
    } else {
      k_hist<<<(unsigned)((t1 - t0 + 255) / 256), 256, 0, st>>>(d_scalars, d_set_ids, cfg, counts, (uint32_t)t0, (uint32_t)t1);
    }
    LAUNCH_CHECK();
    return BPG_OK;
  };
  unsigned gt = (unsigned)((n_terms + 255) / 256);
  if (h_scalars && lane == 0 && n_terms >= (1u << 18)) {
    // Host scalars: the copy runs on the auxiliary stream in pieces and the digit histogram of
    // piece i runs while piece i+1 is still on the bus (hides the histogram of a 2^20-term launch behind the copy).
    const int pieces = 4;
    size_t per = (((n_terms + pieces - 1) / pieces) + tile_terms - 1) / tile_terms * tile_terms;
    CK(cudaEventRecord(ctx->ev_fork, st));  // d_scalars (staging) is free once earlier work is done
    CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    for (int i = 0; i < pieces; i++) {
      size_t t0 = (size_t)i * per, t1 = std::min(n_terms, t0 + per);
      if (t0 >= t1) break;
      if (!ctx->ev_chunk[i]) CK(cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming));
      CK(cudaMemcpyAsync((uint8_t*)d_scalars + t0 * 32, h_scalars + t0 * 32, (t1 - t0) * 32, cudaMemcpyHostToDevice,
                         ctx->aux_stream));
      CK(cudaEventRecord(ctx->ev_chunk[i], ctx->aux_stream));
      CK(cudaStreamWaitEvent(st, ctx->ev_chunk[i], 0));
      rc = hist(t0, t1);
      if (rc) return rc;
    }
  } else {
    if (h_scalars) CK(cudaMemcpyAsync((void*)d_scalars, h_scalars, n_terms * 32, cudaMemcpyHostToDevice, st));
    rc = hist(0, n_terms);
    if (rc) return rc;
  }
  if (radix) {
    prof_mark(ctx, BPG_PROF_SCAN);
    k_rs_scan<<<1, 1024, 0, st>>>(part_count, P, part_base);
    LAUNCH_CHECK();
    prof_mark(ctx, BPG_PROF_SCATTER);
    k_rs_scatter<<<(unsigned)((n_terms + tile_terms - 1) / tile_terms), RS_THREADS, scatter_smem, st>>>(
        d_scalars, d_set_ids, d_point_ids, cfg, lb, P, part_base, part_cursor, (uint2*)(ws + o_pairs), tpt, stage_cap);
    LAUNCH_CHECK();
    k_rs_finish<<<P, fin_threads, ((size_t)1 << lb) * 4 + (size_t)fin_cap * 4, st>>>((const uint2*)(ws + o_pairs), part_base, cfg.B, lb,
                                                                                    P, offsets, L.entries, bins, fin_cap);
    LAUNCH_CHECK();
  } else {
    prof_mark(ctx, BPG_PROF_SCAN);
    k_scan_tiles<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(counts, cfg.B, tiles);
    LAUNCH_CHECK();
    k_scan_spine<<<1, 1024, 0, st>>>(tiles, (uint32_t)ntiles, offsets, cfg.B);
    LAUNCH_CHECK();
    k_scan_apply<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(counts, cfg.B, tiles, offsets, bins);
    LAUNCH_CHECK();
    prof_mark(ctx, BPG_PROF_SCATTER);
    k_scatter<<<gt, 256, 0, st>>>(d_scalars, d_set_ids, d_point_ids, cfg, offsets, counts, L.entries);
    LAUNCH_CHECK();
  }
  // accumulation schedule: (bucket, segment) items by decreasing length; over-long buckets -> big list
  k_size_scatter<<<(cfg.B + 255) / 256, 256, 0, st>>>(offsets, cfg, sched, L.big_count, L.big_list);
  LAUNCH_CHECK();
  if (curve == 1) return msm_accum_reduce_stark(L);
  rc = msm_accum_ristretto(L);
  if (rc) return rc;
  const uint32_t* level0 = (windowed && cfg.gsub > 1) ? L.merged : L.buckets;
  return msm_reduce_ristretto(L, level0);
}
