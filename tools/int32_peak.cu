// INT32 issue-rate microbenchmark for the roofline denominators:
// independent IMAD, IMAD.WIDE.U32, carry-chained IMAD.WIDE.U32.X, IADD3, their mix,
// and whole fe_mul / mixed-add throughput, on all SMs.  Prints one JSON object.
#include <cstdio>
#include <cuda_runtime.h>
#include "../mpc_bulletproof_b200/csrc/ge.cuh"
using namespace bpg;

constexpr int ITERS = 2048;
constexpr int UNR = 8;

__global__ void k_imad(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t x[UNR];
  for (int i = 0; i < UNR; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNR; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
  }
  uint32_t s = 0;
  for (int i = 0; i < UNR; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad_wide(uint32_t* out, uint32_t a, uint32_t b) {
  unsigned long long x[UNR];
  for (int i = 0; i < UNR; i++) x[i] = threadIdx.x + i;
  uint32_t m = a + threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNR; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(m), "r"(b));
  }
  unsigned long long s = 0;
  for (int i = 0; i < UNR; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}
// carry-chained pairs as fe_mul issues them: 4-long chains, two chains interleaved
__global__ void k_imad_wide_x(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t x[16];
  for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
  uint32_t m = a + threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
    mad_wide_cc(x[0], x[1], m, b);
    madc_wide_cc(x[2], x[3], m, b);
    madc_wide_cc(x[4], x[5], m, b);
    madc_wide_cc(x[6], x[7], m, b);
    mad_wide_cc(x[8], x[9], m, b);
    madc_wide_cc(x[10], x[11], m, b);
    madc_wide_cc(x[12], x[13], m, b);
    madc_wide_cc(x[14], x[15], m, b);
  }
  uint32_t s = 0;
  for (int i = 0; i < 16; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_iadd3(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t x[UNR];
  for (int i = 0; i < UNR; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNR; i++) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x[i]) : "r"(a), "r"(b));
  }
  uint32_t s = 0;
  for (int i = 0; i < UNR; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 1 wide MAD : 1 add, independent streams
__global__ void k_mix(uint32_t* out, uint32_t a, uint32_t b) {
  unsigned long long x[UNR];
  uint32_t y[UNR];
  for (int i = 0; i < UNR; i++) { x[i] = threadIdx.x + i; y[i] = i; }
  uint32_t m = a + threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNR; i++) {
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(m), "r"(b));
      asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(y[i]) : "r"(a), "r"(b));
    }
  }
  unsigned long long s = 0;
  for (int i = 0; i < UNR; i++) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}
constexpr int FE_ITERS = 512;
__global__ void k_fe_mul(uint32_t* out, const uint32_t* in) {
  fe a, b, c, d;
  fe_load(a, in + (threadIdx.x & 31) * 8);
  b = a; b.v[0] ^= threadIdx.x; c = a; c.v[1] += blockIdx.x; d = b; d.v[2] ^= 77;
  for (int it = 0; it < FE_ITERS; it++) {
    a = fe_mul(a, b);
    c = fe_mul(c, d);
    b = fe_mul(b, a);
    d = fe_mul(d, c);
  }
  fe r = fe_add(fe_add(a, b), fe_add(c, d));
  fe_store(out + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 8, r);
}
constexpr int MADD_ITERS = 256;
__global__ void __launch_bounds__(128) k_madd(uint32_t* out, const uint32_t* in) {
  ge_niels q;
  ge_load_niels(q, in + (threadIdx.x & 7) * 24);
  ge_ext acc = ge_identity();
  for (int it = 0; it < MADD_ITERS; it++) acc = ge_madd(acc, q, it & 1);
  ge_store_ext(out + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 32, acc);
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

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  int blocks = sms * 8, threads = 256;
  uint32_t *out, *in;
  cudaMalloc(&out, (size_t)blocks * threads * 128 + 4096);
  cudaMalloc(&in, 4096);
  cudaMemset(in, 0x5a, 4096);
  double tot = (double)blocks * threads;
  double t;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  t = time_ms([&] { k_imad<<<blocks, threads>>>(out, 3, 5); });
  printf(", \"imad_Tops\": %.3f", tot * ITERS * UNR / t / 1e9);
  t = time_ms([&] { k_imad_wide<<<blocks, threads>>>(out, 3, 5); });
  printf(", \"imad_wide_Tops\": %.3f", tot * ITERS * UNR / t / 1e9);
  t = time_ms([&] { k_imad_wide_x<<<blocks, threads>>>(out, 3, 5); });
  printf(", \"imad_wide_x_Tops\": %.3f", tot * ITERS * 8 / t / 1e9);
  t = time_ms([&] { k_iadd3<<<blocks, threads>>>(out, 3, 5); });
  printf(", \"iadd3_Tops\": %.3f", tot * ITERS * UNR / t / 1e9);
  t = time_ms([&] { k_mix<<<blocks, threads>>>(out, 3, 5); });
  printf(", \"mix_wide_plus_add_Tops_each\": %.3f", tot * ITERS * UNR / t / 1e9);
  t = time_ms([&] { k_fe_mul<<<blocks, threads>>>(out, in); });
  printf(", \"fe_mul_G_per_s\": %.2f", tot * FE_ITERS * 4 / t / 1e6);
  t = time_ms([&] { k_madd<<<sms * 4, 128>>>(out, in); });
  printf(", \"madd_G_per_s_128x4\": %.3f", (double)sms * 4 * 128 * MADD_ITERS / t / 1e6);
  t = time_ms([&] { k_madd<<<sms * 16, 128>>>(out, in); });
  printf(", \"madd_G_per_s_128x16\": %.3f", (double)sms * 16 * 128 * MADD_ITERS / t / 1e6);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf(", \"clock_khz_max\": %d}\n", clk);
  return 0;
}
