// slice_kernels.cuh — row bucketing and the slice COUNT kernel used by the cube path for its big root tables.
//
// Rows are bucketed ONCE per variable by the joint value of the top `dmax` cube digits (slice_key/scatter kernels) so
// that all rows with a given prefix are contiguous.  A root table is cut into slices along its HIGHEST digits; one CTA
// per (root, slice) streams only the matching row segments, histograms the remaining digits with shared-memory
// atomics and writes the slice to its place in the root's dense global table (bic_slice_count_kernel): every cell
// is written exactly once, no global atomics and no memset.  The tree path (tree_kernels.cuh) keeps the tables on
// chip altogether.
#pragma once
#include "bic_kernels.cuh"

namespace urlgpu {

constexpr int kSliceMaxDepth = 16;     // top digits usable for slicing

struct SliceVar {                      // per-variable constants of the slice path
    int c, rv, max_parents, dmax;
    int card[kMaxDenseCand];           // cube order
    const uint8_t *cols[kMaxDenseCand + 1];   // bucketed columns: [0] = child, [1+i] = cube bit i   (each n_stride bytes)
    const uint32_t *prefix_off;        // [P_dmax + 1] row offsets of the full-depth prefixes
    uint32_t P_dmax;
};

// ---- bucketing -------------------------------------------------------------------------------------------
__global__ void slice_key_kernel(BicData d, CandInfo ci_cube, int dmax, uint32_t *__restrict__ keys, uint32_t *__restrict__ hist) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= d.n) return;
    uint32_t key = 0;
    for (int b = ci_cube.c - 1; b >= ci_cube.c - dmax; b--) key = key * (uint32_t)ci_cube.card[b] + d.codes[(int64_t)ci_cube.var[b] * d.n_stride + r];
    keys[r] = key;
    atomicAdd(&hist[key], 1u);
}
// exclusive scan of hist[0..m) into off[0..m], single block
__global__ void slice_scan_kernel(const uint32_t *__restrict__ hist, uint32_t m, uint32_t *__restrict__ off, uint32_t *__restrict__ cursor) {
    __shared__ uint32_t part[1024];
    const uint32_t per = (m + blockDim.x - 1) / blockDim.x;
    const uint32_t b = threadIdx.x * per, e = min(b + per, m);
    uint32_t s = 0;
    for (uint32_t i = b; i < e; i++) s += hist[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (uint32_t i = 0; i < blockDim.x; i++) { const uint32_t t = part[i]; part[i] = run; run += t; }
        off[m] = run;
    }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (uint32_t i = b; i < e; i++) { off[i] = run; cursor[i] = run; run += hist[i]; }
}
__global__ void slice_scatter_kernel(BicData d, CandInfo ci_cube, const uint32_t *__restrict__ keys, uint32_t *__restrict__ cursor,
                                     uint8_t *__restrict__ out /*[(c+1)][n_stride]*/) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= d.n) return;
    const uint32_t pos = atomicAdd(&cursor[keys[r]], 1u);
    out[pos] = d.codes[(int64_t)ci_cube.v * d.n_stride + r];
    for (int i = 0; i < ci_cube.c; i++) out[(int64_t)(i + 1) * d.n_stride + pos] = d.codes[(int64_t)ci_cube.var[i] * d.n_stride + r];
}

// ---- the slice kernel -------------------------------------------------------------------------------------
constexpr int kSliceMaxSeg = 512;

struct SliceShared {
    const uint8_t *col[kMaxCols];   // local (non-slicing) columns, child first
    uint32_t stride[kMaxCols];
    int ncols;
    uint32_t root_cells;            // cells of this slice's root table
    uint32_t hi_cells;              // root_cells / prod_{i<z} card[i]
    uint32_t nseg;
    uint32_t absent_card[kSliceMaxDepth], absent_weight[kSliceMaxDepth];
    int nabsent;
    uint32_t q_base;                // prefix index contribution of the slicing digits
    uint32_t q_stride;              // P_dmax / P_depth
    int root_size;                  // |P|
    uint32_t total_rows;
};

// floor(j / d) by multiplication: m = floor(2^32 / d) (d >= 2), at most one correction step
__device__ __forceinline__ uint32_t fast_div(uint32_t j, uint32_t d, uint32_t m) {
    if (d == 1) return j;
    uint32_t q = __umulhi(j, m);
    if (j - q * d >= d) q++;
    return q;
}

// ---- slice COUNT kernel (cube path roots): same slicing and row walk as bic_slice_kernel, but the slice is simply
// written to its place in the root's dense global table (the slicing digits are the table's most significant digits,
// so slice `si` is the contiguous range [si*slice_cells, (si+1)*slice_cells)).  Every cell is written exactly once:
// no global atomics and no memset, unlike bic_count_global_kernel.
struct SliceCountRoot {
    uint32_t mask;
    uint32_t chunk0;
    uint64_t table_off;   // int32 elements
    uint16_t nslices;
    uint8_t depth;
    uint8_t pad;
};

constexpr int kSliceCountThreads = 1024;

__global__ void __launch_bounds__(kSliceCountThreads) bic_slice_count_kernel(SliceVar sv, const SliceCountRoot *__restrict__ roots, int nroots, int *__restrict__ tables) {
    extern __shared__ __align__(16) int s_tab[];
    __shared__ SliceShared ss;
    __shared__ int s_root;
    __shared__ uint32_t s_segbeg[kSliceMaxSeg + 1];
    __shared__ uint32_t s_segoff[kSliceMaxSeg + 1];
    constexpr int NT = kSliceCountThreads;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 32) { // 32-ary search for the last root with chunk0 <= blockIdx.x
        int lo = 0, hi = nroots; // answer in [lo, hi)
        while (hi - lo > 1) {
            const int step = (hi - lo + 31) / 32;
            const int pos = lo + lane * step;
            const bool le = pos < hi && roots[pos].chunk0 <= blockIdx.x;
            const unsigned m = __ballot_sync(0xffffffffu, le);
            const int last = 31 - __clz(m); // lane 0 always qualifies (roots[lo].chunk0 <= blockIdx.x)
            const int nlo = lo + last * step;
            hi = min(hi, nlo + step);
            lo = nlo;
        }
        if (lane == 0) s_root = lo;
    }
    __syncthreads();
    const SliceCountRoot rt = roots[s_root];
    const int rv = sv.rv;
    const uint32_t si = blockIdx.x - rt.chunk0;
    if (tid == 0) {
        const int lowbit = sv.c - rt.depth;
        uint32_t q = 0, w = 1, rem = si;
        ss.nabsent = 0; ss.nseg = 1;
        for (int b = lowbit; b < sv.c; b++) {
            const uint32_t cb = (uint32_t)sv.card[b];
            if ((rt.mask >> b) & 1) { const uint32_t val = rem % cb; rem /= cb; q += val * w; }
            else { ss.absent_card[ss.nabsent] = cb; ss.absent_weight[ss.nabsent] = w; ss.nabsent++; ss.nseg *= cb; }
            w *= cb;
        }
        ss.q_base = q;
        ss.q_stride = sv.P_dmax / w;
        ss.col[0] = sv.cols[0]; ss.stride[0] = 1;
        uint32_t base = (uint32_t)rv;
        int nc = 1;
        for (int b = 0; b < lowbit; b++)
            if ((rt.mask >> b) & 1) { ss.col[nc] = sv.cols[b + 1]; ss.stride[nc] = base; base *= (uint32_t)sv.card[b]; nc++; }
        ss.ncols = nc; ss.root_cells = base;
    }
    __syncthreads();
    for (uint32_t seg = tid; seg < ss.nseg; seg += NT) {
        uint32_t q = ss.q_base, rem = seg;
        for (int a = 0; a < ss.nabsent; a++) { q += (rem % ss.absent_card[a]) * ss.absent_weight[a]; rem /= ss.absent_card[a]; }
        const uint32_t r0 = sv.prefix_off[(size_t)q * ss.q_stride], r1 = sv.prefix_off[(size_t)(q + 1) * ss.q_stride];
        s_segbeg[seg] = r0;
        s_segoff[seg] = r1 - r0;
    }
    for (uint32_t i = tid; i < ss.root_cells; i += NT) s_tab[i] = 0;
    __syncthreads();
    if (tid < 32) { // exclusive scan of the segment lengths by one warp
        uint32_t carry = 0;
        for (uint32_t b0 = 0; b0 < ss.nseg; b0 += 32) {
            const uint32_t i = b0 + lane;
            const uint32_t len = i < ss.nseg ? s_segoff[i] : 0;
            uint32_t x = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (i < ss.nseg) s_segoff[i] = carry + x - len;
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) { s_segoff[ss.nseg] = carry; ss.total_rows = carry; }
    }
    __syncthreads();
    if (ss.nseg == 1) {
        const uint32_t r0 = s_segbeg[0], r1 = r0 + ss.total_rows;
        const uint32_t a0 = min(r1, (r0 + 15u) & ~15u), a1 = max(a0, r1 & ~15u);
        for (uint32_t r = r0 + tid; r < a0; r += NT) {
            uint32_t idx = 0;
            for (int c = 0; c < ss.ncols; c++) idx += (uint32_t)ss.col[c][r] * ss.stride[c];
            atomicAdd(&s_tab[idx], 1);
        }
        for (uint32_t r = a0 + (uint32_t)tid * 16u; r < a1; r += NT * 16u) {
            uint32_t idx[16];
#pragma unroll
            for (int i = 0; i < 16; i++) idx[i] = 0;
            for (int c = 0; c < ss.ncols; c++) accum16(ld_stream_u4(reinterpret_cast<const uint4 *>(ss.col[c] + r)), ss.stride[c], idx);
#pragma unroll
            for (int i = 0; i < 16; i++) atomicAdd(&s_tab[idx[i]], 1);
        }
        for (uint32_t r = a1 + tid; r < r1; r += NT) {
            uint32_t idx = 0;
            for (int c = 0; c < ss.ncols; c++) idx += (uint32_t)ss.col[c][r] * ss.stride[c];
            atomicAdd(&s_tab[idx], 1);
        }
    } else {
        // fragmented slice: the segments are walked as one concatenated row range; 4 rows per thread are in flight
        uint32_t seg = 0;
        for (uint32_t g0 = tid; g0 < ss.total_rows; g0 += NT * 4) {
            uint32_t rr[4], idx[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t g = g0 + u * NT;
                idx[u] = 0;
                if (g < ss.total_rows) {
                    while (s_segoff[seg + 1] <= g) seg++;
                    rr[u] = s_segbeg[seg] + (g - s_segoff[seg]);
                } else rr[u] = 0xffffffffu;
            }
            for (int c = 0; c < ss.ncols; c++) {
                const uint8_t *__restrict__ col = ss.col[c];
                const uint32_t st = ss.stride[c];
#pragma unroll
                for (int u = 0; u < 4; u++) if (rr[u] != 0xffffffffu) idx[u] += (uint32_t)col[rr[u]] * st;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) if (rr[u] != 0xffffffffu) atomicAdd(&s_tab[idx[u]], 1);
        }
    }
    __syncthreads();
    int4 *__restrict__ dst = reinterpret_cast<int4 *>(tables + rt.table_off + (uint64_t)si * ss.root_cells);
    const int4 *src4 = reinterpret_cast<const int4 *>(s_tab);
    if ((ss.root_cells & 3u) == 0 && (((rt.table_off + (uint64_t)si * ss.root_cells) & 3ull) == 0)) {
        for (uint32_t i = tid; i < ss.root_cells / 4; i += NT) dst[i] = src4[i];
    } else {
        int *d1 = tables + rt.table_off + (uint64_t)si * ss.root_cells;
        for (uint32_t i = tid; i < ss.root_cells; i += NT) d1[i] = s_tab[i];
    }
}

} // namespace urlgpu
