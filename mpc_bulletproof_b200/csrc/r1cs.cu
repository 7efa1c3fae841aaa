// libbpgpu: R1CS scalar preparation and constraint flattening on the device.
#define BPG_FE_OUTLINE 1  // latency-bound kernels: products are calls, not 1.5 KB of inline code each
#include "internal.cuh"
#include "svec_kernels.cuh"

using namespace bpg;

#include "r1cs_dev.inc"
