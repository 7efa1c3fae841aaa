// libbpgpu: R1CS scalar preparation and constraint flattening on the device.
#include "internal.cuh"
#include "svec_kernels.cuh"

using namespace bpg;

#include "r1cs_dev.inc"
