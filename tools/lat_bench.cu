// Lone-warp latency of the building blocks of the latency-bound kernels (tree reductions, comb rounds,
// encodings): cycles per dependent operation, one warp on one SM.  Built twice: products inline, and out of line
// (-DBPG_FE_OUTLINE=1), to see what the call costs a single warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cudart shared [-DBPG_FE_OUTLINE=1] -o tools/lat_bench tools/lat_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../mpc_bulletproof_b200/csrc/ge.cuh"
#include "../mpc_bulletproof_b200/csrc/ge4.cuh"
#include "../mpc_bulletproof_b200/csrc/fe16.cuh"
using namespace bpg;

#define N_IT 256
__global__ void k_lat(uint32_t* out, const uint32_t* in, long long* cyc) {
  __shared__ __align__(16) uint32_t g16[G16_WORDS];
  fe a, b;
  fe_load(b, in + 8);
  fe_load(a, in + (threadIdx.x & 31) * 8);
  long long t[12];
  int k = 0;
  // 0: fe_mul chain
  t[k++] = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; i++) a = fe_mul(a, b);
  t[k++] = clock64();
  // 1: fe_sq chain
#pragma unroll 1
  for (int i = 0; i < N_IT; i++) a = fe_sq(a);
  t[k++] = clock64();
  // 2: two independent fe_mul chains (ILP 2)
  fe c = b;
#pragma unroll 1
  for (int i = 0; i < N_IT; i++) {
    a = fe_mul(a, b);
    c = fe_mul(c, b);
  }
  t[k++] = clock64();
  a = fe_add(a, c);
  // 3: ge4_add chain (three product levels)
  ge4 p, q;
  p.c = a;
  q.c = b;
#pragma unroll 1
  for (int i = 0; i < N_IT; i++) p = ge4_add(p, q);
  t[k++] = clock64();
  // 4: ge4_add_cached chain (two levels)
#pragma unroll 1
  for (int i = 0; i < N_IT; i++) p = ge4_add_cached(p, q);
  t[k++] = clock64();
  // 5: thread-level mixed addition chain (seven products)
  ge_ext e;
  e.X = p.c;
  e.Y = a;
  e.Z = b;
  e.T = c;
  ge_niels nq;
  nq.ypx = a;
  nq.ymx = b;
  nq.t2d = c;
#pragma unroll 1
  for (int i = 0; i < N_IT; i++) e = ge_madd(e, nq, false);
  t[k++] = clock64();
  // 6: thread-level full addition chain (nine products)
  ge_ext f = e;
#pragma unroll 1
  for (int i = 0; i < N_IT; i++) e = ge_add(e, f);
  t[k++] = clock64();
  // 7: sixteen-lane squaring chain (the encodings' inverse square root)
  grp16 g;
  g.sm = g16;
  g.k = threadIdx.x & 15u;
  g.half = (threadIdx.x >> 4) & 1u;
  g.par = 0;
  fe16 x = fe16_from_fe(g, e.X);
  x = fe16_sqn<true>(g, x, N_IT);
  t[k++] = clock64();
  // 8: the same straight-line (unrolled) ge4_add x 32, executed once: instruction fetch of cold code
#pragma unroll
  for (int i = 0; i < 32; i++) p = ge4_add(p, q);
  t[k++] = clock64();
  fe r = fe16_to_fe(g, x);
  r = fe_add(fe_add(r, p.c), fe_add(e.X, e.Y));
  fe_store(out + threadIdx.x * 8, r);
  if (threadIdx.x == 0)
    for (int i = 0; i + 1 < k; i++) cyc[i] = t[i + 1] - t[i];
}

int main() {
  uint32_t *in, *out;
  long long* cyc;
  cudaMalloc(&in, 4096);
  cudaMalloc(&out, 4096);
  cudaMalloc(&cyc, 128);
  uint32_t h[1024];
  for (int i = 0; i < 1024; i++) h[i] = 0x9e3779b9u * (i + 1) >> 3;
  cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  const char* names[] = {"fe_mul", "fe_sq", "2 x fe_mul (ILP)", "ge4_add (3 levels)", "ge4_add_cached (2 levels)", "ge_madd (7 products)",
                         "ge_add (9 products)", "fe16_sq", "ge4_add straight-line x32 (cold code)"};
  int div[] = {N_IT, N_IT, N_IT, N_IT, N_IT, N_IT, N_IT, N_IT, 32};
  for (int rep = 0; rep < 3; rep++) {
    k_lat<<<1, 32>>>(out, in, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("error %s\n", cudaGetErrorString(e));
      return 1;
    }
  }
  long long c[16];
  cudaMemcpy(c, cyc, 9 * 8, cudaMemcpyDeviceToHost);
  printf("{");
  for (int i = 0; i < 9; i++) printf("\"%s\": %.0f%s", names[i], (double)c[i] / div[i], i < 8 ? ", " : "");
  printf("}\n");
  return 0;
}
