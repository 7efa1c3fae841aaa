// Umbrella: the whole ristretto255 Pippenger pipeline.
#pragma once
#include "msm_sort_kernels.cuh"
#include "msm_accum_kernels.cuh"
#include "msm_reduce_kernels.cuh"
