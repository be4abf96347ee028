// atomic_probe.cu — microbenchmark: shared-memory and L2 (global RED) histogram throughput on B200.
// Design input for the K1 counting kernels (how many table updates per second the machine sustains).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s line %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ uint32_t rng(uint32_t &s){ s ^= s<<13; s ^= s>>17; s ^= s<<5; return s; }

// each thread performs `iters` updates on a shared table of `cells` ints, indices pseudo-random (precomputed in regs)
__global__ void smem_atomics(int cells, int iters, int *sink) {
    extern __shared__ int h[];
    for (int i = threadIdx.x; i < cells; i += blockDim.x) h[i] = 0;
    __syncthreads();
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int it = 0; it < iters; it++) {
        uint32_t idx[16];
#pragma unroll
        for (int k = 0; k < 16; k++) idx[k] = rng(s) % (uint32_t)cells;
#pragma unroll
        for (int k = 0; k < 16; k++) atomicAdd(&h[idx[k]], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) sink[blockIdx.x] = h[0];
}
__global__ void gmem_red(int *table, uint32_t cells, int iters) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 777u;
    for (int it = 0; it < iters; it++) {
        uint32_t idx[16];
#pragma unroll
        for (int k = 0; k < 16; k++) idx[k] = rng(s) % cells;
#pragma unroll
        for (int k = 0; k < 16; k++) atomicAdd(&table[idx[k]], 1);
    }
}
// same index generation without the atomics: the ALU floor to subtract
__global__ void rng_only(int cells, int iters, int *sink) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int k = 0; k < 16; k++) acc += rng(s) % (uint32_t)cells;
    if (acc == 0xdeadbeef) sink[0] = acc;
}
__global__ void stream_read(const uint4 *p, size_t n, int *sink) {
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { uint4 v = p[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0xdeadbeef) sink[0] = acc;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("device %s, %d SMs, L2 %d MB, smem optin %zu\n", pr.name, pr.multiProcessorCount, pr.l2CacheSize >> 20, pr.sharedMemPerBlockOptin);
    int *sink; CK(cudaMalloc(&sink, 1 << 20));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    CK(cudaFuncSetAttribute(smem_atomics, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    const int sms = pr.multiProcessorCount;
    float ms;
    {
        const int iters = 256, threads = 1024;
        rng_only<<<sms, threads>>>(1000, iters, sink);
        cudaEventRecord(a); rng_only<<<sms, threads>>>(50000, iters, sink); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        cudaEventElapsedTime(&ms, a, b);
        printf("rng only: %.3f ms for %.3e idx -> %.1f G idx/s\n", ms, (double)sms * threads * iters * 16, (double)sms * threads * iters * 16 / ms / 1e6);
    }
    for (int cells : {64, 1024, 12288, 49152}) {
        for (int threads : {256, 1024}) {
            const int iters = 256;
            int ctas_per_sm = cells * 4 <= 48 * 1024 ? (1024 / threads) * 2 : 1;
            if (ctas_per_sm * threads > 2048) ctas_per_sm = 2048 / threads;
            if ((size_t)ctas_per_sm * cells * 4 > 200 * 1024) ctas_per_sm = 1;
            int grid = sms * ctas_per_sm;
            smem_atomics<<<grid, threads, cells * 4>>>(cells, iters, sink);
            cudaEventRecord(a);
            smem_atomics<<<grid, threads, cells * 4>>>(cells, iters, sink);
            cudaEventRecord(b); CK(cudaEventSynchronize(b));
            cudaEventElapsedTime(&ms, a, b);
            double ops = (double)grid * threads * iters * 16;
            printf("smem atomics cells=%6d threads=%4d ctas/sm=%d: %.3f ms, %.1f G updates/s (%.2f per clk per SM @1.9GHz)\n", cells, threads, ctas_per_sm, ms,
                   ops / ms / 1e6, ops / ms / 1e6 / sms / 1.9);
        }
    }
    for (size_t mb : {1, 2, 8, 32, 64, 256, 1024}) {
        uint32_t cells = (uint32_t)(mb << 20) / 4;
        int *t; CK(cudaMalloc(&t, (size_t)cells * 4)); CK(cudaMemset(t, 0, (size_t)cells * 4));
        const int iters = 64, threads = 256, grid = sms * 8;
        gmem_red<<<grid, threads>>>(t, cells, iters);
        cudaEventRecord(a);
        gmem_red<<<grid, threads>>>(t, cells, iters);
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        cudaEventElapsedTime(&ms, a, b);
        double ops = (double)grid * threads * iters * 16;
        printf("global RED table=%5zu MB: %.3f ms, %.1f G updates/s\n", mb, ms, ops / ms / 1e6);
        cudaFree(t);
    }
    for (size_t mb : {16, 60, 512, 4096}) {
        size_t n = (mb << 20) / 16;
        uint4 *p; CK(cudaMalloc(&p, n * 16)); CK(cudaMemset(p, 1, n * 16));
        int reps = mb < 100 ? 50 : 5;
        stream_read<<<sms * 8, 512>>>(p, n, sink);
        cudaEventRecord(a);
        for (int r = 0; r < reps; r++) stream_read<<<sms * 8, 512>>>(p, n, sink);
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        cudaEventElapsedTime(&ms, a, b);
        printf("stream read %5zu MB x%d: %.1f GB/s\n", mb, reps, (double)n * 16 * reps / ms / 1e6);
        cudaFree(p);
    }
    return 0;
}
