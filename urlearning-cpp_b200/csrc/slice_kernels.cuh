// slice_kernels.cuh — small helpers shared by the shared-memory counting kernels (tree_kernels.cuh).
#pragma once
#include "bic_kernels.cuh"

namespace urlgpu {

// floor(j / d) by multiplication: m = floor(2^32 / d) (d >= 2), at most one correction step
__device__ __forceinline__ uint32_t fast_div(uint32_t j, uint32_t d, uint32_t m) {
    if (d == 1) return j;
    uint32_t q = __umulhi(j, m);
    if (j - q * d >= d) q++;
    return q;
}

} // namespace urlgpu
