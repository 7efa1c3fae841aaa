// libbpgpu: Stark-curve policy (tables, MSM entry points, inner-product rounds).
#include "internal.cuh"
#include "stark_point_kernels.cuh"
#include "ipp_kernels.cuh"
#include "ipp_kernels_t.cuh"

using namespace bpg;

#include "stark_msm.inc"
#include "stark_ipp.inc"
