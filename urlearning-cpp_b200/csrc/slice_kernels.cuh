// slice_kernels.cuh — K1 "slice" path: root tables counted in shared-memory slices, whole subtrees derived on chip.
//
// Same algebra as the cube path (bic_kernels.cuh): the candidate family is subset-closed, candidates are ordered by
// ascending arity ("cube bits"), the parent of a set is the set plus its lowest missing cube bit, and the ROOTS are
// the sets of layer L* (= max_parents+1 when it exists).  The subtree of a root P is { P ^ D : D subset of the
// trailing-ones run of P } (z = lowest missing bit of P, run = bits 0..z-1).  Every descendant is the marginal of
// P's table over some of the z LOWEST digits, so if P's table is cut into slices along its HIGHEST digits, every
// slice can produce its share of every descendant table independently, entirely in shared memory:
//
//   1. per variable, rows are bucketed ONCE by the joint value of the top `dmax` cube digits (slice_key/scatter
//      kernels) so that all rows with a given prefix are contiguous;
//   2. one CTA per (root, slice): the slice fixes the values of the root's top present digits; the CTA streams only
//      the matching row segments (one segment per combination of the ABSENT top digits), histograms the remaining
//      digits with shared-memory atomics, then sums the run digits out one at a time (level by level: after
//      digit y every table for D subset {y..z-1} exists), scoring each table as it appears;
//   3. per-(root, D) exact int64 log-likelihood accumulators are added to global memory once per CTA.
//
// No contingency table is ever written to HBM; traffic is the bucketed rows (L2 resident) and a few accumulators.
// Arithmetic is identical to the other paths (exact int64 sums of the float table), so results are bit-identical.
#pragma once
#include "bic_kernels.cuh"

namespace urlgpu {

constexpr int kSliceMaxRun = 9;        // z <= 9: at most 512 tables per CTA
constexpr int kSliceMaxDepth = 16;     // top digits usable for slicing
constexpr int kSliceThreads = 512;

struct SliceVar {                      // per-variable constants of the slice path
    int c, rv, max_parents, dmax;
    int card[kMaxDenseCand];           // cube order
    const uint8_t *cols[kMaxDenseCand + 1];   // bucketed columns: [0] = child, [1+i] = cube bit i   (each n_stride bytes)
    const uint32_t *prefix_off;        // [P_dmax + 1] row offsets of the full-depth prefixes
    uint32_t P_dmax;
};

struct SliceRoot {
    uint32_t mask;        // cube mask of the root
    uint32_t chunk0;      // first CTA of this root
    uint32_t acc_off;     // first accumulator (2^z of them, indexed by D)
    uint16_t nslices;
    uint8_t depth;        // top digits (bits c-1 .. c-depth) spanned by the slicing digits
    uint8_t z;            // run length = lowest missing bit
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

__global__ void __launch_bounds__(kSliceThreads) bic_slice_kernel(SliceVar sv, const SliceRoot *__restrict__ roots, int nroots, const long long *__restrict__ qlog,
                                                                 long long *__restrict__ acc_out, uint32_t table_cells_budget) {
    extern __shared__ __align__(16) int s_tab[];             // [budget] tables, then toff, tsize, acc
    __shared__ SliceShared ss;
    __shared__ int s_root;
    __shared__ uint32_t s_segbeg[kSliceMaxSeg + 1];         // first row of each segment (bucketed row index)
    __shared__ uint32_t s_segoff[kSliceMaxSeg + 1];         // exclusive prefix of the segment lengths
    uint32_t *s_toff = reinterpret_cast<uint32_t *>(s_tab + table_cells_budget);
    uint32_t *s_tsize = s_toff + (1u << kSliceMaxRun);
    long long *s_acc = reinterpret_cast<long long *>(s_tsize + (1u << kSliceMaxRun));
    const int tid = threadIdx.x;
    if (tid == 0) {
        int lo = 0, hi = nroots - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (roots[mid].chunk0 <= blockIdx.x) lo = mid; else hi = mid - 1;
        }
        s_root = lo;
    }
    __syncthreads();
    const SliceRoot rt = roots[s_root];
    const int z = rt.z, rv = sv.rv;
    const uint32_t ntab = 1u << z;
    if (tid == 0) {
        // decode the slice: digits c-1 .. c-depth; present ones take their value from the slice index
        const uint32_t si = blockIdx.x - rt.chunk0;
        const int lowbit = sv.c - rt.depth; // lowest top digit
        uint32_t q = 0, w = 1, rem = si;    // w = weight of a digit inside the depth-`depth` prefix index
        ss.nabsent = 0; ss.nseg = 1;
        for (int b = lowbit; b < sv.c; b++) { // lowest present digit = least significant slice digit
            const uint32_t cb = (uint32_t)sv.card[b];
            if ((rt.mask >> b) & 1) { const uint32_t val = rem % cb; rem /= cb; q += val * w; }
            else { ss.absent_card[ss.nabsent] = cb; ss.absent_weight[ss.nabsent] = w; ss.nabsent++; ss.nseg *= cb; }
            w *= cb;
        }
        ss.q_base = q;
        ss.q_stride = sv.P_dmax / w; // w = P_depth
        // local columns: child, then the present digits below the slicing zone, ascending
        ss.col[0] = sv.cols[0]; ss.stride[0] = 1;
        uint32_t base = (uint32_t)rv, hi_cells = 1;
        int nc = 1, size = 0;
        for (int b = 0; b < sv.c; b++)
            if ((rt.mask >> b) & 1) {
                size++;
                if (b < lowbit) { ss.col[nc] = sv.cols[b + 1]; ss.stride[nc] = base; base *= (uint32_t)sv.card[b]; nc++; if (b >= z) hi_cells *= (uint32_t)sv.card[b]; }
            }
        ss.ncols = nc; ss.root_cells = base; ss.hi_cells = hi_cells * (uint32_t)rv; ss.root_size = size;
    }
    __syncthreads();
    // table sizes (products only) in parallel, segment bounds in parallel
    for (uint32_t D = tid; D < ntab; D += kSliceThreads) {
        uint32_t cells = ss.hi_cells;
        for (int i = 0; i < z; i++) if (!((D >> i) & 1)) cells *= (uint32_t)sv.card[i];
        s_tsize[D] = cells;
        s_acc[D] = 0;
    }
    for (uint32_t seg = tid; seg < ss.nseg; seg += kSliceThreads) {
        uint32_t q = ss.q_base, rem = seg;
        for (int a = 0; a < ss.nabsent; a++) { q += (rem % ss.absent_card[a]) * ss.absent_weight[a]; rem /= ss.absent_card[a]; }
        const uint32_t r0 = sv.prefix_off[(size_t)q * ss.q_stride], r1 = sv.prefix_off[(size_t)(q + 1) * ss.q_stride];
        s_segbeg[seg] = r0;
        s_segoff[seg] = r1 - r0; // length for now
    }
    for (uint32_t i = tid; i < ss.root_cells; i += kSliceThreads) s_tab[i] = 0;
    __syncthreads();
    if (tid == 0) {
        // table offsets in creation order: digit y = z-1 .. 0, within a digit increasing D
        s_toff[0] = 0;
        uint32_t next = ss.root_cells;
        for (int y = z - 1; y >= 0; y--)
            for (uint32_t e = 0; e < (1u << (z - 1 - y)); e++) {
                const uint32_t D = (e << (y + 1)) | (1u << y);
                s_toff[D] = next;
                next += s_tsize[D];
            }
    }
    if (tid == 32) { // exclusive scan of the segment lengths
        uint32_t run = 0;
        for (uint32_t sgi = 0; sgi < ss.nseg; sgi++) { const uint32_t len = s_segoff[sgi]; s_segoff[sgi] = run; run += len; }
        s_segoff[ss.nseg] = run;
        ss.total_rows = run;
    }
    __syncthreads();
    // ---- count this slice's rows: the segments are walked as one concatenated range so all threads stay busy ----
    if (ss.nseg == 1) {
        const uint32_t r0 = s_segbeg[0], r1 = r0 + ss.total_rows;
        const uint32_t a0 = min(r1, (r0 + 15u) & ~15u), a1 = max(a0, r1 & ~15u);
        for (uint32_t r = r0 + tid; r < a0; r += kSliceThreads) {
            uint32_t idx = 0;
            for (int c = 0; c < ss.ncols; c++) idx += (uint32_t)ss.col[c][r] * ss.stride[c];
            atomicAdd(&s_tab[idx], 1);
        }
        for (uint32_t r = a0 + (uint32_t)tid * 16u; r < a1; r += kSliceThreads * 16u) {
            uint32_t idx[16];
#pragma unroll
            for (int i = 0; i < 16; i++) idx[i] = 0;
            for (int c = 0; c < ss.ncols; c++) accum16(ld_stream_u4(reinterpret_cast<const uint4 *>(ss.col[c] + r)), ss.stride[c], idx);
#pragma unroll
            for (int i = 0; i < 16; i++) atomicAdd(&s_tab[idx[i]], 1);
        }
        for (uint32_t r = a1 + tid; r < r1; r += kSliceThreads) {
            uint32_t idx = 0;
            for (int c = 0; c < ss.ncols; c++) idx += (uint32_t)ss.col[c][r] * ss.stride[c];
            atomicAdd(&s_tab[idx], 1);
        }
    } else {
        uint32_t seg = 0;
        for (uint32_t g = tid; g < ss.total_rows; g += kSliceThreads) {
            while (s_segoff[seg + 1] <= g) seg++; // g only grows: amortised O(1)
            const uint32_t r = s_segbeg[seg] + (g - s_segoff[seg]);
            uint32_t idx = 0;
            for (int c = 0; c < ss.ncols; c++) idx += (uint32_t)ss.col[c][r] * ss.stride[c];
            atomicAdd(&s_tab[idx], 1);
        }
    }
    __syncthreads();
    const int lane = tid & 31;
    // ---- score the root slice (only when the root itself is a scored set) ----
    if (ss.root_size <= sv.max_parents) {
        long long acc = score_configs(s_tab, rv, 0, ss.root_cells / rv, qlog, tid, kSliceThreads);
        acc = warp_sum_ll(acc);
        if (lane == 0 && acc != 0) atomicAdd(reinterpret_cast<unsigned long long *>(&s_acc[0]), (unsigned long long)acc);
    }
    // ---- sum the run digits out, highest first; after digit y every table with D subset {y..z-1} exists ----
    uint32_t Bc = ss.root_cells / ss.hi_cells; // configurations below the run top: prod_{i<z} card
    for (int y = z - 1; y >= 0; y--) {
        const uint32_t ry = (uint32_t)sv.card[y];
        Bc /= ry;                                  // stride (in configurations) of digit y while all lower digits are present
        const uint32_t magic = Bc > 1 ? 0xFFFFFFFFu / Bc : 0;
        for (uint32_t e = 0; e < (1u << (z - 1 - y)); e++) {
            const uint32_t Dsrc = e << (y + 1), Ddst = Dsrc | (1u << y);
            const uint32_t out_cfg = s_tsize[Ddst] / (uint32_t)rv;
            const int *__restrict__ src = s_tab + s_toff[Dsrc];
            int *__restrict__ dst = s_tab + s_toff[Ddst];
            const bool score = (ss.root_size - __popc(Ddst)) <= sv.max_parents;
            long long acc = 0;
            for (uint32_t j = tid; j < out_cfg; j += kSliceThreads) {
                const uint32_t hi = fast_div(j, Bc, magic), lo = j - hi * Bc;
                const uint32_t pc0 = lo + hi * ry * Bc;
                int nij = 0;
                for (int k = 0; k < rv; k++) {
                    int cnt = 0;
                    for (uint32_t a = 0; a < ry; a++) cnt += src[(pc0 + a * Bc) * rv + k];
                    dst[j * rv + k] = cnt;
                    nij += cnt;
                    if (score && cnt > 1) acc += __ldg(&qlog[cnt]);
                }
                if (score && nij > 1) acc -= __ldg(&qlog[nij]);
            }
            if (score && out_cfg > 0) {
                acc = warp_sum_ll(acc);
                if (lane == 0 && acc != 0) atomicAdd(reinterpret_cast<unsigned long long *>(&s_acc[Ddst]), (unsigned long long)acc);
            }
        }
        __syncthreads();
    }
    for (uint32_t D = tid; D < ntab; D += kSliceThreads) {
        const long long a = s_acc[D];
        if (a != 0) atomicAdd(reinterpret_cast<unsigned long long *>(&acc_out[rt.acc_off + D]), (unsigned long long)a);
    }
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

// scores[res_mask(P ^ D)] for every (root, D) whose set is in the scored family
__global__ void slice_finalize_kernel(BicData d, CandInfo ci_res, const SliceRoot *__restrict__ roots, int nroots, const uint8_t *__restrict__ perm /*cube bit -> result bit*/,
                                      const long long *__restrict__ acc, uint32_t total, float *__restrict__ scores, long long *__restrict__ ll_fixed) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int lo = 0, hi = nroots - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (roots[mid].acc_off <= i) lo = mid; else hi = mid - 1;
    }
    const SliceRoot rt = roots[lo];
    const uint32_t D = i - rt.acc_off;
    const uint32_t T = rt.mask ^ D;
    if (__popc(T) > ci_res.max_parents) return;
    uint32_t rm = 0;
    for (int b = 0; b < ci_res.c; b++) if ((T >> b) & 1) rm |= 1u << perm[b];
    float pen = (float)(ci_res.rv - 1);
    for (int b = 0; b < ci_res.c; b++)
        if ((rm >> b) & 1) pen = __fmul_rn(pen, (float)ci_res.card[b]);
    scores[rm] = bic_finalize(acc[i], pen, d.base);
    if (ll_fixed) ll_fixed[rm] = acc[i];
}

} // namespace urlgpu
