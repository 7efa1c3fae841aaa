// libbpgpu: one Pippenger launch, ristretto255 bucket reduction, Horner, and the few-term path.
#include "msm_launch.cuh"
#include "msm_reduce_kernels.cuh"

using namespace bpg;

cudaError_t msm_kernels_init() {
  return cudaFuncSetAttribute(k_reduce_pairs_final, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RPB_SMEM);
}

// Small arrays: a leaf block of RT_QUADS quad chunks of LC buckets plus its in-block tree; large arrays: one thread
// per chunk of 16.  (Measured at 2 x 2^16 buckets, an IPP round at n = 2^16: quad chunks of 8 and thread chunks of 4
// tie, thread chunks of 8 / 16 and quad chunks of 4 are 2-11 % of a proof slower; gpurun_out/r1i_tune.jsonl)
void msm_reduce_geometry(const MsmCfg& cfg, bool* thread_leaf, uint32_t* LC, uint32_t* tiles0) {
  *thread_leaf = cfg.nb >= (1u << 17);
  *LC = *thread_leaf ? 16 : (cfg.nb > (1u << 15) ? 8 : 4);
  static const char* force = getenv("BPG_LEAF_LC");  // tuning: 4 or 8 (quad leaf)
  if (force && !*thread_leaf) *LC = atoi(force) == 8 ? 8 : 4;
  *tiles0 = *thread_leaf ? (cfg.nb + *LC - 1) / *LC : (cfg.nb + RT_QUADS * *LC - 1) / (RT_QUADS * *LC);
}

int msm_identity(bpg_ctx* ctx, cudaStream_t st, int curve, int nsets, uint32_t* d_out_ext) {
  // (X, Y, Z, T) = (0, 1, 1, 0); the Stark policy's identity is all zero words
  if (curve == 1) CK(cudaMemsetAsync(d_out_ext, 0, (size_t)nsets * 128, st));
  else {
    k_set_identity<<<nsets, 32, 0, st>>>(d_out_ext);
    LAUNCH_CHECK();
  }
  return BPG_OK;
}

bool msm_small_applies(size_t n_terms, int nsets) { return n_terms <= SMALL_MAX_TERMS && nsets <= SMALL_MAX_SETS; }

int msm_small_ristretto(bpg_ctx* ctx, cudaStream_t st, uint8_t*, int lane, const uint32_t* table_base, size_t n_points,
                        const uint32_t* d_scalars, size_t n_terms, const uint8_t* d_set_ids, const uint32_t* d_point_ids,
                        int nsets, uint32_t* d_out_ext) {
  unsigned nblk = (unsigned)((n_terms + SMALL_QUADS - 1) / SMALL_QUADS);
  prof_mark(ctx, BPG_PROF_ACCUM);
  if (nblk == 1) {
    k_msm_small<<<1, SMALL_THREADS, 0, st>>>(table_base, d_scalars, d_set_ids, d_point_ids, (uint32_t)n_terms,
                                             (uint32_t)std::max<size_t>(n_points, 1), nsets, bias_for(4), d_out_ext);
    LAUNCH_CHECK();
  } else {
    int rc = ensure_ws(ctx, (size_t)nblk * nsets * 128, lane);
    if (rc) return rc;
    uint32_t* parts = (uint32_t*)(lane ? ctx->ws_aux : ctx->ws);
    k_msm_small<<<nblk, SMALL_THREADS, 0, st>>>(table_base, d_scalars, d_set_ids, d_point_ids, (uint32_t)n_terms,
                                                (uint32_t)std::max<size_t>(n_points, 1), nsets, bias_for(4), parts);
    LAUNCH_CHECK();
    k_msm_small_fin<<<1, SMALL_THREADS, 0, st>>>(parts, nblk, nsets, d_out_ext);
    LAUNCH_CHECK();
  }
  prof_mark(ctx, -1);
  return BPG_OK;
}

int msm_reduce_ristretto(MsmLaunch& L, const uint32_t* level0) {
  bpg_ctx* ctx = L.ctx;
  cudaStream_t st = L.st;
  const MsmCfg& cfg = L.cfg;
  const uint32_t rarr = L.rarr;
  const size_t pair_words = L.pair_words;
  prof_mark(ctx, BPG_PROF_REDUCE);
  {
    uint32_t* final_out = L.windowed ? L.out_ext : L.wins;
    uint32_t t = L.tiles0;
    uint32_t* pa[2] = {L.pairs, L.pairs + 2 * pair_words};
    int cur = 0;
    uint32_t* oa = t == 1 ? final_out : pa[cur];
    if (L.thread_leaf) {
      k_reduce_leaf_thread<16><<<(rarr * t + RL_THREADS - 1) / RL_THREADS, RL_THREADS, 0, st>>>(level0, cfg.nb, t, rarr, oa,
                                                                                               pa[cur] + pair_words);
    } else if (L.LC == 8) {
      k_reduce_leaf<8><<<rarr * t, RT_THREADS, 0, st>>>(level0, cfg.nb, t, oa, pa[cur] + pair_words);
    } else {
      k_reduce_leaf<4><<<rarr * t, RT_THREADS, 0, st>>>(level0, cfg.nb, t, oa, pa[cur] + pair_words);
    }
    LAUNCH_CHECK();
    while (t > 1) {
      uint32_t n = t;
      const uint32_t* ia = pa[cur];
      const uint32_t* iy = pa[cur] + pair_words;
      if (n <= RPB_PAIRS) {
        // what is left fits one block per array: finish here
        k_reduce_pairs_final<<<rarr, RPB_THREADS, RPB_SMEM, st>>>(ia, iy, n, final_out);
        LAUNCH_CHECK();
        break;
      }
      t = (n + RP_PAIRS - 1) / RP_PAIRS;
      cur ^= 1;
      oa = t == 1 ? final_out : pa[cur];
      k_reduce_pairs<<<rarr * t, RP_THREADS, 0, st>>>(ia, iy, n, t, oa, pa[cur] + pair_words);
      LAUNCH_CHECK();
    }
  }
  if (!L.windowed) {
    prof_mark(ctx, BPG_PROF_HORNER);
    k_horner<<<L.nsets, 32, 0, st>>>(L.wins, cfg, L.out_ext);
    LAUNCH_CHECK();
  }
  prof_mark(ctx, -1);
  return BPG_OK;
}
