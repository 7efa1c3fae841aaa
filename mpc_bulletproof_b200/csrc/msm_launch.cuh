// One Pippenger launch, shared between the translation units of its stages (msm.cu: sort and schedule;
// msm_accum.cu / msm_reduce.cu: ristretto255 bucket arithmetic; msm_stark.cu: Stark-curve bucket arithmetic).
#pragma once
#include "internal.cuh"
#include "msm_sort_kernels.cuh"

struct MsmLaunch {
  bpg_ctx* ctx;
  cudaStream_t st;
  int lane;
  bpg::MsmCfg cfg;
  bpg::AccSched sched;
  int nsets;
  bool windowed;
  const uint32_t* table;
  uint32_t *offsets, *entries, *buckets, *merged, *seg_part;
  uint32_t *big_count, *big_list, *big_part;
  uint32_t *pairs, *wins, *out_ext;
  size_t pair_words, max_items, max_multi;
  uint32_t rarr, tiles0, LC;
  bool thread_leaf;
};
// reduction geometry of the ristretto255 path (msm_reduce.cu) / the Stark path (msm_stark.cu): leaf tiles per array
void msm_reduce_geometry(const bpg::MsmCfg& cfg, bool* thread_leaf, uint32_t* LC, uint32_t* tiles0);
uint32_t msm_stark_tiles0(const bpg::MsmCfg& cfg);
int msm_accum_ristretto(MsmLaunch& L);
int msm_reduce_ristretto(MsmLaunch& L, const uint32_t* level0);
int msm_small_ristretto(bpg_ctx* ctx, cudaStream_t st, uint8_t* ws_or_null, int lane, const uint32_t* table_base, size_t n_points,
                        const uint32_t* d_scalars, size_t n_terms, const uint8_t* d_set_ids, const uint32_t* d_point_ids,
                        int nsets, uint32_t* d_out_ext);
bool msm_small_applies(size_t n_terms, int nsets);
int msm_identity(bpg_ctx* ctx, cudaStream_t st, int curve, int nsets, uint32_t* d_out_ext);
int msm_accum_reduce_stark(MsmLaunch& L);
