// libbpgpu: comb construction (one-time table combs; per-proof combs of folded generators).
#include "internal.cuh"
#include "comb_build_kernels.cuh"

using namespace bpg;

// folded generators (2 m0 points) from the generator combs, their doubling chains, their combs
int comb_materialize(bpg_ctx* ctx, cudaStream_t s, const uint32_t* gen_comb, uint32_t g_id, uint32_t h_id, const uint32_t* wG,
                     const uint32_t* wH, size_t n, size_t m0, uint32_t* folded, uint32_t* chain, uint32_t* comb) {
  const size_t npts = 2 * m0;
  CombMat M;
  M.comb = gen_comb;
  M.g_id = g_id;
  M.h_id = h_id;
  M.q_id = 0;
  M.wG = wG;
  M.wH = wH;
  M.q_mul = nullptr;
  M.n = (uint32_t)n;
  M.m0 = (uint32_t)m0;
  M.bias4 = bias_for(4);
  // warps per output: enough warps for several full waves (each lane walks 64 / (ws msplit) windows of its terms)
  M.msplit = COMB_MAT_SPLIT;
  const size_t warps = npts * M.msplit;
  k_comb_materialize<<<(unsigned)((warps + CB_THREADS / 32 - 1) / (CB_THREADS / 32)), CB_THREADS, 0, s>>>(M, folded);
  LAUNCH_CHECK();
  k_comb_chain<<<(unsigned)((npts * 4 + CB_THREADS - 1) / CB_THREADS), CB_THREADS, 0, s>>>(folded, (uint32_t)npts, M.msplit, chain);
  LAUNCH_CHECK();
  k_comb_multiples<<<(unsigned)((npts * COMB_WINDOWS + CB_THREADS - 1) / CB_THREADS), CB_THREADS, 0, s>>>(
      chain, (uint32_t)(npts * COMB_WINDOWS), comb);
  LAUNCH_CHECK();
  return BPG_OK;
}

// combs of a few ad-hoc points (a proof's points before their scalars exist): the doubling chains do not depend
// on the scalars, so a verifier starts them as soon as it has the proof and its final MSM finds combs.
int comb_from_points(bpg_ctx* ctx, cudaStream_t s, const uint8_t* d_comp, size_t n, uint32_t* ext, uint32_t* chain,
                     uint32_t* comb, uint32_t* bad) {
  k_decode_ext<<<(unsigned)n, 32, 0, s>>>(d_comp, (uint32_t)n, ext, bad);
  LAUNCH_CHECK();
  k_comb_chain<<<(unsigned)((n * 4 + CB_THREADS - 1) / CB_THREADS), CB_THREADS, 0, s>>>(ext, (uint32_t)n, 1, chain);
  LAUNCH_CHECK();
  k_comb_multiples<<<(unsigned)((n * COMB_WINDOWS + CB_THREADS - 1) / CB_THREADS), CB_THREADS, 0, s>>>(
      chain, (uint32_t)(n * COMB_WINDOWS), comb);
  LAUNCH_CHECK();
  return BPG_OK;
}

// ---------------------------------------------------------------------------
// combs of a resident table (one-time, like bpg_table_set_windows): 64 x 8 affine-Niels multiples per point
// ---------------------------------------------------------------------------
extern "C" int bpg_table_build_comb(bpg_ctx* ctx, bpg_table* t) {
  if (!ctx || !t) return BPG_ERR_ARG;
  if (t->comb || t->n == 0) return BPG_OK;
  CK(cudaSetDevice(ctx->device));
  uint32_t* comb = nullptr;
  cudaError_t e = cudaMalloc(&comb, t->n * (size_t)COMB_ENTRIES * COMB_AFFINE_WORDS * 4);
  if (e != cudaSuccess) {
    ctx->last_cuda = (int)e;
    cudaGetLastError();
    return BPG_ERR_NOMEM;
  }
  const size_t CH = 1 << 15;
  for (size_t first = 0; first < t->n; first += CH) {
    size_t cnt = std::min(CH, t->n - first);
    k_table_comb_build<<<(unsigned)cnt, COMB_WINDOWS, 0, ctx->stream>>>(t->niels, (uint32_t)first, comb);
    ctx->launches++;
  }
  cudaError_t se = cudaStreamSynchronize(ctx->stream);
  if (se != cudaSuccess || cudaGetLastError() != cudaSuccess) {
    ctx->last_cuda = (int)se;
    cudaFree(comb);
    return BPG_ERR_CUDA;
  }
  t->comb = comb;
  return BPG_OK;
}
extern "C" int bpg_table_has_comb(const bpg_table* t) { return t && t->comb ? 1 : 0; }
