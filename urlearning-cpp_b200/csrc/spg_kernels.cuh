// spg_kernels.cuh — the sparse parent graph of one variable as a DEVICE query structure (sm_100a).
//
// Replaces bestscorecalculators::SparseParentBitwise (score_cache/sparse_parent_bitwise.cpp:24-110), the search side's
// "best cached subset" look-up: the variable's cache entries are sorted by score (best first); for every variable p a
// bitset over the sorted entries marks the entries that do NOT use p as a parent (:60-78, after the flip).  The best
// score among the subsets of an allowed set U is the FIRST set bit of the AND of the bitsets of all variables outside U
// (:90-110).  Here one warp answers one query: lane l owns 64 entries of the current 2048-entry chunk, ANDs the
// words of the excluded variables (coalesced 256-byte rows), a ballot finds the first lane with a surviving entry.  The
// best entries come first, so a query usually ends in the first chunk.  Natural batch producers are the pattern
// database construction (heuristic/static_pattern_database.cpp:224-248: one look-up per (sub-network, leaf)) and the
// expansion of an A* frontier.
#pragma once
#include "common.cuh"

namespace urlgpu {

// not_used[p * bw + w], bit i = entry 64 w + i does not contain variable p (bits beyond n are 0)
__global__ void __launch_bounds__(256) spg_build_kernel(const uint64_t *__restrict__ masks, uint64_t n, int words, int variable_count, uint64_t bw,
                                                        uint64_t *__restrict__ not_used) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= bw * (uint64_t)variable_count) return;
    const int p = (int)(t / bw);
    const uint64_t w = t - (uint64_t)p * bw;
    uint64_t x = 0;
    const uint64_t e0 = w * 64;
    for (int i = 0; i < 64 && e0 + i < n; i++)
        if (!((masks[(e0 + i) * (uint64_t)words + (p >> 6)] >> (p & 63)) & 1)) x |= (uint64_t)1 << i;
    not_used[t] = x;
}

constexpr int kSpgMaxWords = 4;

__global__ void __launch_bounds__(256) spg_query_kernel(const uint64_t *__restrict__ not_used, const float *__restrict__ scores, uint64_t n, uint64_t bw,
                                                        int variable_count, int words, const uint64_t *__restrict__ allowed, uint64_t nq,
                                                        float *__restrict__ best, long long *__restrict__ index) {
    const uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    uint64_t U[kSpgMaxWords];
#pragma unroll
    for (int w = 0; w < kSpgMaxWords; w++) U[w] = w < words ? __ldg(allowed + q * (uint64_t)words + w) : 0;
    long long found = -1;
    for (uint64_t w0 = 0; w0 < bw && found < 0; w0 += 32) {
        const uint64_t w = w0 + lane;
        uint64_t x = w < bw ? ~(uint64_t)0 : 0;
        for (int p = 0; p < variable_count; p++) {
            if ((U[p >> 6] >> (p & 63)) & 1) continue;            // allowed as a parent: uniform across the warp
            if (w < bw) x &= __ldg(not_used + (uint64_t)p * bw + w);
        }
        const unsigned hit = __ballot_sync(0xffffffffu, x != 0);
        if (hit) {
            const int l = __ffs(hit) - 1;
            const uint64_t xx = __shfl_sync(0xffffffffu, x, l);
            found = (long long)((w0 + l) * 64 + (uint64_t)(__ffsll((long long)xx) - 1));
        }
    }
    if (lane == 0) {
        best[q] = found >= 0 ? scores[found] : 3.402823466e+38f;   // std::numeric_limits<float>::max() (:104-106)
        if (index) index[q] = found;
    }
}

} // namespace urlgpu
