// Dev microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, A in TMEM, B in smem) vs N and issue count.
#include <cstdio>
#include "../mppi_tf_b200/csrc/mppi_mlp.cuh"
using namespace mppi;

template <int N>
__global__ void __launch_bounds__(128) bench(long long *out, int n_mma, int ncta_sync)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    for (int i = threadIdx.x; i < 40960 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc(128, N);
        const uint32_t lbo = (uint32_t)(N >> 3) * 128u, sbo = 128u;
        uint32_t ph = 0;
        for (int rep = 0; rep < 4; rep++) {
            long long t0 = clock64();
            if (elect_one()) {
#pragma unroll 1
                for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const uint64_t bdesc = make_smem_desc(smem_u32(smem) + (uint32_t)k * 2u * lbo, lbo, sbo);
                        umma_ts(tmem, tmem + 256 + (uint32_t)k * 8u, bdesc, idesc, 1u);
                    }
                }
                umma_commit(&bar);
            }
            __syncwarp();
            long long t1 = clock64();
            mbar_wait(&bar, ph);
            ph ^= 1;
            long long t2 = clock64();
            if (threadIdx.x == 0 && blockIdx.x == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N>
void run(long long *d, int n_mma)
{
    cudaFuncSetAttribute(bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    bench<N><<<148, 128, 65536>>>(d, n_mma, 0);
    long long h[8];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("N=%3d n=%3d issue %lld  total %lld cycles -> %.1f cyc/mma (issue %.1f)  [%s]\n", N, n_mma, h[6], h[7], (double)h[7] / n_mma,
           (double)h[6] / n_mma, cudaGetErrorString(e));
}

int main()
{
    long long *d;
    cudaMalloc(&d, 64);
    for (int n : {8, 32, 128}) {
        run<16>(d, n);
        run<32>(d, n);
        run<64>(d, n);
        run<128>(d, n);
        run<256>(d, n);
    }
    return 0;
}
