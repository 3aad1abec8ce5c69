// Dev microbenchmark: tcgen05.ld / tcgen05.st throughput per SM vs number of warps.
#include <cstdio>
#include "../mppi_tf_b200/csrc/mppi_mlp.cuh"
using namespace mppi;

__global__ void __launch_bounds__(512) bw(long long *out, int iters, int mode)
{
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int warp = threadIdx.x >> 5;
    const uint32_t base = slot + ((uint32_t)(32 * (warp & 3)) << 16) + 64u * ((warp >> 2) & 7);
    uint32_t v0[32], v1[32];
#pragma unroll
    for (int i = 0; i < 32; i++) { v0[i] = threadIdx.x + i; v1[i] = i; }
    tmem_st32(base, v0); tmem_st32(base + 32, v1); tc_wait_st();
    __syncthreads();
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; it++) {
        if (mode == 0) {            // loads: 64 columns per iteration
            tmem_ld32(base, v0);
            tmem_ld32(base + 32, v1);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i++) acc += v0[i] ^ v1[i];
        } else {                    // stores
#pragma unroll
            for (int i = 0; i < 32; i++) { v0[i] += it; }
            tmem_st32(base, v0);
            tmem_st32(base + 32, v0);
            tc_wait_st();
        }
    }
    long long t1 = clock64();
    if (acc == 0x12345678) out[7] = acc;
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(slot, 512);
}

int main()
{
    long long *d; cudaMalloc(&d, 64);
    for (int mode = 0; mode < 2; mode++)
        for (int nthreads : {32, 128, 256, 512}) {
            const int iters = 2000;
            bw<<<148, nthreads>>>(d, iters, mode);
            long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            double bytes = (double)iters * (nthreads / 32) * 32 * 64 * 4;
            printf("%s warps=%2d: %lld cycles, %.1f B/clk/SM, %.1f cycles per 64-col warp op [%s]\n", mode ? "ST" : "LD", nthreads / 32, h,
                   bytes / h, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
