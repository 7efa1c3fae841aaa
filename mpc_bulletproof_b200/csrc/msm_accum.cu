// libbpgpu: one Pippenger launch, ristretto255 bucket accumulation (and the merge of bucket groups).
#include "msm_launch.cuh"
#include "msm_accum_kernels.cuh"

using namespace bpg;

int msm_accum_ristretto(MsmLaunch& L) {
  bpg_ctx* ctx = L.ctx;
  cudaStream_t st = L.st;
  const MsmCfg& cfg = L.cfg;
  prof_mark(ctx, BPG_PROF_ACCUM);
  k_accum<<<(unsigned)((L.max_items + ACC_THREADS - 1) / ACC_THREADS), ACC_THREADS, 0, st>>>(L.table, L.offsets, L.entries,
                                                                                            L.sched, L.buckets, L.seg_part);
  LAUNCH_CHECK();
  prof_mark(ctx, BPG_PROF_ACCUM_BIG);
  k_accum_fix<<<(unsigned)std::min<size_t>((L.max_multi * 4 + FIX_THREADS - 1) / FIX_THREADS, (size_t)ctx->sm_count * 8),
                FIX_THREADS, 0, st>>>(L.offsets, L.sched, L.seg_part, L.buckets);
  LAUNCH_CHECK();
  unsigned gbig = std::min<unsigned>(cfg.big_cap, (unsigned)ctx->sm_count * 4);
  k_accum_big<<<gbig, BIG_THREADS, 0, st>>>(L.table, L.offsets, L.entries, cfg, L.buckets, L.big_count, L.big_list, L.big_part);
  LAUNCH_CHECK();
  k_accum_big_fin<<<gbig, BIG_THREADS, 0, st>>>(cfg, L.buckets, L.big_count, L.big_list, L.big_part);
  LAUNCH_CHECK();
  if (L.windowed && cfg.gsub > 1) {
    prof_mark(ctx, BPG_PROF_COMBINE);
    unsigned nq = (unsigned)L.nsets * cfg.nb;
    k_merge<<<(nq * 4 + MERGE_THREADS - 1) / MERGE_THREADS, MERGE_THREADS, 0, st>>>(L.buckets, cfg, L.merged);
    LAUNCH_CHECK();
  }
  return BPG_OK;
}
