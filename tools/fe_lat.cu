// latency of dependent squarings / multiplications for ONE warp (the encode / inversion regime)
#include <cstdio>
#include <cuda_runtime.h>
#include "../mpc_bulletproof_b200/csrc/ge.cuh"
using namespace bpg;

// (b) dedicated squaring, chained: 28 cross products (doubled once) + 8 squares
__device__ __forceinline__ fe fe_sq_chain(const fe& A) {
  const uint32_t* a = A.v;
  uint32_t e[16], o[16];
#pragma unroll
  for (int i = 0; i < 16; i++) e[i] = o[i] = 0;
  // cross products a_i a_j, i<j, column i+j.  even columns -> e (pairs aligned at even index), odd -> o (o[k] = column k+1)
  // row j=0 has none with i<j... use rows by the smaller index i: products a_i * a_j for j>i
  // column parity = (i+j)&1.
  // i = 0: j=1..7
  // even columns (j even): (0,2)->c2,(0,4)->c4,(0,6)->c6 ; odd (j odd): (0,1)->c1,(0,3)->c3,(0,5)->c5,(0,7)->c7
  mul_wide(e[2], e[3], a[0], a[2]); mul_wide(e[4], e[5], a[0], a[4]); mul_wide(e[6], e[7], a[0], a[6]);
  mul_wide(o[0], o[1], a[0], a[1]); mul_wide(o[2], o[3], a[0], a[3]); mul_wide(o[4], o[5], a[0], a[5]); mul_wide(o[6], o[7], a[0], a[7]);
  // i = 1: j=2..7: columns 3..8.  odd: (1,2)->c3,(1,4)->c5,(1,6)->c7 ; even: (1,3)->c4,(1,5)->c6,(1,7)->c8
  mad_wide_cc(o[2], o[3], a[1], a[2]); madc_wide_cc(o[4], o[5], a[1], a[4]); madc_wide_cc(o[6], o[7], a[1], a[6]); o[8] = addc(0u, 0u);
  mad_wide_cc(e[4], e[5], a[1], a[3]); madc_wide_cc(e[6], e[7], a[1], a[5]); madc_wide_top(e[8], e[9], a[1], a[7]);
  // i = 2: j=3..7: columns 5..9. odd: (2,3)->c5,(2,5)->c7,(2,7)->c9 ; even: (2,4)->c6,(2,6)->c8
  mad_wide_cc(o[4], o[5], a[2], a[3]); madc_wide_cc(o[6], o[7], a[2], a[5]); madc_wide_cc(o[8], o[9], a[2], a[7]); o[10] = addc(0u, 0u);
  mad_wide_cc(e[6], e[7], a[2], a[4]); madc_wide_cc(e[8], e[9], a[2], a[6]); e[10] = addc(0u, 0u);
  // i = 3: j=4..7: columns 7..10. odd: (3,4)->c7,(3,6)->c9 ; even: (3,5)->c8,(3,7)->c10
  mad_wide_cc(o[6], o[7], a[3], a[4]); madc_wide_cc(o[8], o[9], a[3], a[6]); o[10] = addc(o[10], 0u);
  mad_wide_cc(e[8], e[9], a[3], a[5]); madc_wide_cc(e[10], e[11], a[3], a[7]); e[12] = addc(0u, 0u);
  // i = 4: j=5..7: columns 9..11. odd: (4,5)->c9,(4,7)->c11 ; even: (4,6)->c10
  mad_wide_cc(o[8], o[9], a[4], a[5]); madc_wide_cc(o[10], o[11], a[4], a[7]); o[12] = addc(0u, 0u);
  mad_wide_cc(e[10], e[11], a[4], a[6]); e[12] = addc(e[12], 0u);
  // i = 5: j=6,7: columns 11,12. odd: (5,6)->c11 ; even: (5,7)->c12
  mad_wide_cc(o[10], o[11], a[5], a[6]); o[12] = addc(o[12], 0u);
  mad_wide_cc(e[12], e[13], a[5], a[7]); e[14] = addc(0u, 0u);
  // i = 6: j=7: column 13 (odd)
  mad_wide_cc(o[12], o[13], a[6], a[7]); o[14] = addc(0u, 0u);
  // cross = e + (o << 32)
  uint32_t r[16];
  r[0] = e[0];
  r[1] = add_cc(e[1], o[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) r[k] = addc_cc(e[k], o[k - 1]);
  r[15] = addc(e[15], o[14]);
  // double
#pragma unroll
  for (int k = 15; k >= 1; k--) r[k] = (r[k] << 1) | (r[k - 1] >> 31);
  r[0] <<= 1;
  // + squares
  uint32_t lo, hi;
  mul_wide(lo, hi, a[0], a[0]); r[0] = add_cc(r[0], lo); r[1] = addc_cc(r[1], hi);
#pragma unroll
  for (int i = 1; i < 8; i++) {
    mul_wide(lo, hi, a[i], a[i]);
    r[2 * i] = addc_cc(r[2 * i], lo);
    r[2 * i + 1] = addc_cc(r[2 * i + 1], hi);
  }
  return fe_reduce512(r);
}

// (c) ILP form: independent 64-bit products, column sums with 64-bit adds
__device__ __forceinline__ fe fe_sq_ilp(const fe& A) {
  const uint32_t* a = A.v;
  // column k accumulators: lo 64 bits + hi count
  unsigned long long cl[15];
  uint32_t ch[15];
#pragma unroll
  for (int k = 0; k < 15; k++) { cl[k] = 0; ch[k] = 0; }
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = i + 1; j < 8; j++) {
      unsigned long long p = (unsigned long long)a[i] * a[j];
      unsigned long long s = cl[i + j] + p;
      ch[i + j] += s < p;
      cl[i + j] = s;
    }
  // double cross sums, add squares
#pragma unroll
  for (int k = 0; k < 15; k++) {
    ch[k] = (ch[k] << 1) | (uint32_t)(cl[k] >> 63);
    cl[k] <<= 1;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    unsigned long long p = (unsigned long long)a[i] * a[i];
    unsigned long long s = cl[2 * i] + p;
    ch[2 * i] += s < p;
    cl[2 * i] = s;
  }
  // carry propagate: value = sum_k (cl[k] + ch[k] 2^64) 2^(32k)
  uint32_t r[16];
  unsigned long long carry = 0;  // carry into column k (64-bit is enough: < 2^40)
  uint32_t carry_hi = 0;
#pragma unroll
  for (int k = 0; k < 15; k++) {
    unsigned long long s = cl[k] + carry;
    uint32_t c2 = s < carry;
    r[k] = (uint32_t)s;
    // next carry = (s >> 32) + ((ch[k] + c2 + carry_hi... ) << 32)
    unsigned long long hi = (unsigned long long)(ch[k] + c2) + carry_hi;
    carry = (s >> 32) + (hi << 32);
    carry_hi = (uint32_t)(hi >> 32);
  }
  r[15] = (uint32_t)carry;
  return fe_reduce512(r);
}

template <int V>
__global__ void k_chain(uint32_t* out, const uint32_t* in, int iters) {
  fe a;
  fe_load(a, in + (threadIdx.x & 31) * 8);
  a.v[0] ^= threadIdx.x + blockIdx.x;
  for (int it = 0; it < iters; it++) {
    if (V == 0) a = fe_mul(a, a);
    if (V == 1) a = fe_sq_chain(a);
    if (V == 2) a = fe_sq_ilp(a);
  }
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

int main() {
  uint32_t *out, *in;
  cudaMalloc(&out, 148 * 16 * 256 * 32 + 4096);
  cudaMalloc(&in, 4096);
  cudaMemset(in, 0x5a, 4096);
  const int iters = 4096;
  // correctness: all variants must agree
  uint32_t h[3][8];
  k_chain<0><<<1, 32>>>(out, in, 7); cudaMemcpy(h[0], out + 8 * 5, 32, cudaMemcpyDeviceToHost);
  k_chain<1><<<1, 32>>>(out, in, 7); cudaMemcpy(h[1], out + 8 * 5, 32, cudaMemcpyDeviceToHost);
  k_chain<2><<<1, 32>>>(out, in, 7); cudaMemcpy(h[2], out + 8 * 5, 32, cudaMemcpyDeviceToHost);
  int ok1 = 1, ok2 = 1;
  for (int i = 0; i < 8; i++) { ok1 &= h[0][i] == h[1][i]; ok2 &= h[0][i] == h[2][i]; }
  printf("{\"agree_chain\": %d, \"agree_ilp\": %d", ok1, ok2);
  struct { int blocks, threads; const char* name; } cfgs[] = {{1, 32, "1warp"}, {148, 128, "1warp_per_smsp"}, {148 * 4, 128, "4warp_per_smsp"}, {148 * 8, 256, "16warp_per_smsp"}};
  for (auto& c : cfgs) {
    double t0 = time_ms([&] { k_chain<0><<<c.blocks, c.threads>>>(out, in, iters); });
    double t1 = time_ms([&] { k_chain<1><<<c.blocks, c.threads>>>(out, in, iters); });
    double t2 = time_ms([&] { k_chain<2><<<c.blocks, c.threads>>>(out, in, iters); });
    printf(", \"%s_ns_per_op\": {\"mul\": %.1f, \"sq_chain\": %.1f, \"sq_ilp\": %.1f}", c.name, t0 * 1e6 / iters, t1 * 1e6 / iters, t2 * 1e6 / iters);
  }
  printf("}\n");
  return 0;
}
