// Throughput of fe_mul variants at the occupancy of the bucket-accumulation kernel (4 warps per sub-partition)
// and at full occupancy: G fe_mul/s over all SMs.  V0 = the shipped fe_mul (fe.cuh); V1..= candidates.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cudart shared -o tools/fe_mul_bench tools/fe_mul_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../mpc_bulletproof_b200/csrc/ge.cuh"
#include "fe_mul_variants.cuh"
using namespace bpg;

template <int V>
__device__ __forceinline__ fe mulv(const fe& a, const fe& b) {
  if (V == 0) return fe_mul(a, b);
  if (V == 1) return fe_mul_v1(a, b);
  if (V == 2) return fe_mul_v2(a, b);
  return fe_mul(a, b);
}

// 4 independent chains per thread (the ILP a mixed addition offers)
template <int V>
__global__ void __launch_bounds__(128) k_tput(uint32_t* out, const uint32_t* in, int iters) {
  fe a[4], b;
  fe_load(b, in + 8);
#pragma unroll
  for (int c = 0; c < 4; c++) {
    fe_load(a[c], in + (threadIdx.x & 31) * 8);
    a[c].v[0] ^= threadIdx.x + blockIdx.x * 131 + c;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < 4; c++) a[c] = mulv<V>(a[c], b);
    b.v[1] ^= a[0].v[3];
  }
  fe r = fe_add(fe_add(a[0], a[1]), fe_add(a[2], a[3]));
  fe_store(out + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 8, r);
}

// ONE dependent chain per thread: the latency regime of the tree / encoding / comb-round kernels
template <int V>
__global__ void __launch_bounds__(128) k_lat(uint32_t* out, const uint32_t* in, int iters) {
  fe a, b;
  fe_load(b, in + 8);
  fe_load(a, in + (threadIdx.x & 31) * 8);
  a.v[0] ^= threadIdx.x + blockIdx.x * 131;
  for (int it = 0; it < iters; it++) a = mulv<V>(a, b);
  fe_store(out + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 8, a);
}

template <typename F>
static double time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

template <int V>
static void check(uint32_t* out, const uint32_t* in, uint32_t ref[8], const char* name, int* ok) {
  k_tput<V><<<1, 32>>>(out, in, 9);
  uint32_t h[8];
  cudaMemcpy(h, out + 8 * 7, 32, cudaMemcpyDeviceToHost);
  if (V == 0) for (int i = 0; i < 8; i++) ref[i] = h[i];
  int same = 1;
  for (int i = 0; i < 8; i++) same &= ref[i] == h[i];
  printf("\"agree_%s\": %d, ", name, same);
  *ok &= same;
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t *out, *in;
  cudaMalloc(&out, (size_t)sms * 16 * 128 * 32 + 4096);
  cudaMalloc(&in, 4096);
  uint32_t hin[1024];
  for (int i = 0; i < 1024; i++) hin[i] = 0x9e3779b9u * (i + 1) ^ (i << 7);
  for (int i = 7; i < 1024; i += 8) hin[i] &= 0x7fffffffu;
  cudaMemcpy(in, hin, 4096, cudaMemcpyHostToDevice);
  uint32_t ref[8];
  int ok = 1;
  printf("{");
  check<0>(out, in, ref, "v0", &ok);
  check<1>(out, in, ref, "v1", &ok);
  check<2>(out, in, ref, "v2", &ok);
  const int iters = 2048;
  struct { int blocks_per_sm; const char* name; } cfgs[] = {{4, "4warps_per_smsp"}, {8, "8warps_per_smsp"}, {12, "12warps_per_smsp"}};
  for (auto& c : cfgs) {
    int blocks = sms * c.blocks_per_sm;
    double muls = (double)blocks * 128 * 4 * iters;
    double t0 = time_ms([&] { k_tput<0><<<blocks, 128>>>(out, in, iters); });
    double t1 = time_ms([&] { k_tput<1><<<blocks, 128>>>(out, in, iters); });
    double t2 = time_ms([&] { k_tput<2><<<blocks, 128>>>(out, in, iters); });
    printf("\"%s_Gmul_per_s\": {\"v0\": %.1f, \"v1\": %.1f, \"v2\": %.1f}, ", c.name, muls / t0 / 1e6, muls / t1 / 1e6, muls / t2 / 1e6);
  }
  // latency: one warp per SM sub-partition (148 blocks x 128 threads), one chain per thread
  {
    const int it2 = 4096;
    double t0 = time_ms([&] { k_lat<0><<<sms, 128>>>(out, in, it2); });
    double t1 = time_ms([&] { k_lat<1><<<sms, 128>>>(out, in, it2); });
    double t2 = time_ms([&] { k_lat<2><<<sms, 128>>>(out, in, it2); });
    printf("\"lone_warp_ns_per_mul\": {\"v0\": %.1f, \"v1\": %.1f, \"v2\": %.1f}, ", t0 * 1e6 / it2, t1 * 1e6 / it2, t2 * 1e6 / it2);
    double u0 = time_ms([&] { k_lat<0><<<sms * 2, 128>>>(out, in, it2); });
    double u2 = time_ms([&] { k_lat<2><<<sms * 2, 128>>>(out, in, it2); });
    printf("\"two_warps_ns_per_mul\": {\"v0\": %.1f, \"v2\": %.1f}, ", u0 * 1e6 / it2, u2 * 1e6 / it2);
  }
  printf("\"ok\": %d}\n", ok);
  return 0;
}
