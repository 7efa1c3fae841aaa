// libbpgpu: one Pippenger launch, Stark-curve bucket arithmetic (accumulate, reduce, Horner).
#include "msm_launch.cuh"
#include "stark_msm.cuh"

using namespace bpg;

uint32_t msm_stark_tiles0(const MsmCfg& cfg) { return (cfg.nb + SLEAF_LC - 1) / SLEAF_LC; }

// same stages as the ristretto255 path, short-Weierstrass bucket arithmetic (stark_msm.cuh)
int msm_accum_reduce_stark(MsmLaunch& L) {
  bpg_ctx* ctx = L.ctx;
  cudaStream_t st = L.st;
  const MsmCfg& cfg = L.cfg;
  const int nsets = L.nsets;
  const bool windowed = L.windowed;
  const size_t pair_words = L.pair_words;
  prof_mark(ctx, BPG_PROF_ACCUM);
  k_stark_accum<<<(unsigned)((L.max_items + SACC_THREADS - 1) / SACC_THREADS), SACC_THREADS, 0, st>>>(
      L.table, L.offsets, L.entries, L.sched, L.buckets, L.seg_part);
  LAUNCH_CHECK();
  prof_mark(ctx, BPG_PROF_ACCUM_BIG);
  k_stark_fix<<<(unsigned)std::min<size_t>((L.max_multi + 127) / 128, (size_t)ctx->sm_count * 8), 128, 0, st>>>(
      L.offsets, L.sched, L.seg_part, L.buckets);
  LAUNCH_CHECK();
  unsigned gb = std::min<unsigned>(cfg.big_cap, (unsigned)ctx->sm_count * 4);
  k_stark_big<<<gb, SBIG_THREADS, 0, st>>>(L.table, L.offsets, L.entries, cfg, L.buckets, L.big_count, L.big_list, L.big_part);
  LAUNCH_CHECK();
  k_stark_big_fin<<<std::min<unsigned>((cfg.big_cap + 127) / 128, (unsigned)ctx->sm_count), 128, 0, st>>>(
      cfg, L.buckets, L.big_count, L.big_list, L.big_part);
  LAUNCH_CHECK();
  const uint32_t* lvl0 = L.buckets;
  if (windowed && cfg.gsub > 1) {
    prof_mark(ctx, BPG_PROF_COMBINE);
    k_stark_merge<<<((unsigned)nsets * cfg.nb + 127) / 128, 128, 0, st>>>(L.buckets, cfg, L.merged);
    LAUNCH_CHECK();
    lvl0 = L.merged;
  }
  prof_mark(ctx, BPG_PROF_REDUCE);
  uint32_t arrays = windowed ? (uint32_t)nsets : cfg.narr;
  uint32_t* fin = windowed ? L.out_ext : L.wins;
  // leaf: large arrays one thread per chunk of 8 (throughput), small ones one quad per chunk of 4
  // plus the in-block tree (latency); the pairs levels and Horner are quad-cooperative
  const bool sthread_leaf = cfg.nb >= (1u << 17);
  uint32_t t = sthread_leaf ? (cfg.nb + SLEAF_LC - 1) / SLEAF_LC : (cfg.nb + SRT_QUADS * 4 - 1) / (SRT_QUADS * 4);
  uint32_t* pa[2] = {L.pairs, L.pairs + 2 * pair_words};
  int cur = 0;
  uint32_t* oa = t == 1 ? fin : pa[cur];
  if (sthread_leaf) k_stark_leaf<<<(arrays * t + 127) / 128, 128, 0, st>>>(lvl0, cfg.nb, t, arrays, oa, pa[cur] + pair_words);
  else k_stark_leaf4<4><<<arrays * t, SRT_THREADS, 0, st>>>(lvl0, cfg.nb, t, oa, pa[cur] + pair_words);
  LAUNCH_CHECK();
  while (t > 1) {
    uint32_t n = t;
    t = (n + SRP_PAIRS - 1) / SRP_PAIRS;
    const uint32_t* ia = pa[cur];
    const uint32_t* iy = pa[cur] + pair_words;
    cur ^= 1;
    oa = t == 1 ? fin : pa[cur];
    k_stark_pairs4<<<arrays * t, SRP_THREADS, 0, st>>>(ia, iy, n, t, oa, pa[cur] + pair_words);
    LAUNCH_CHECK();
  }
  if (!windowed) {
    prof_mark(ctx, BPG_PROF_HORNER);
    k_stark_horner4<<<nsets, 32, 0, st>>>(L.wins, cfg, L.out_ext);
    LAUNCH_CHECK();
  }
  prof_mark(ctx, -1);
  return BPG_OK;
}
