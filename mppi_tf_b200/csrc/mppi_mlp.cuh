// Learned-MLP dynamics on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Model (row A13; behaviour of /root/reference/scripts/src/models/nn_model.py:54-60,215-239,289-304
// re-shaped to BASELINE config 4):   X = (concat(x,u) - Xmean)/Xstd  ->  h1 = relu(W1^T X + b1)
//   -> h2 = relu(W2^T h1 + b2) -> d = W3^T h2 + b3 -> x' = x + d*Ystd + Ymean.
//
// Mapping: one CTA tile = 128 samples = the 128 TMEM lanes (UMMA_M = 128, cta_group::1); thread r of
// the CTA owns sample row r for the whole rollout (fp32 state in registers).  Each layer is
//   D[128 x N] (fp32, TMEM) = A[128 x K] (bf16, TMEM, written by the threads with tcgen05.st)
//                             x B[N x K]^T (bf16 weights, K-major canonical layout in shared memory,
//                                           staged once per CTA by a TMA bulk copy)
// issued by ONE thread as K/16 tcgen05.mma instructions and tracked with tcgen05.commit on an
// mbarrier; the epilogue (bias + ReLU + bf16 pack) goes TMEM -> registers -> TMEM as the next A.
#pragma once
#include <cuda_bf16.h>

#include "mppi_device.cuh"

namespace mppi {

constexpr int kMlpH = 128;        // hidden width (both hidden layers)
constexpr int kMlpKin = 16;       // input features padded to one UMMA K step
constexpr int kMlpNout = 16;      // output features padded to the minimum N for M = 128
constexpr int kMlpThreads = 128;  // one thread per TMEM lane / sample row

// shared-memory weight blob (bytes), canonical K-major no-swizzle core-matrix layout:
//   element (n, k) of B[N x K] at ((k/8)*(N/8) + n/8)*128 + (n%8)*16 + (k%8)*2
constexpr int kW1Bytes = kMlpH * kMlpKin * 2;      //  4 KB   N = 128, K = 16
constexpr int kW2Bytes = kMlpH * kMlpH * 2;        // 32 KB   N = 128, K = 128
constexpr int kW3Bytes = kMlpNout * kMlpH * 2;     //  4 KB   N = 16,  K = 128
constexpr int kWBlobBytes = kW1Bytes + kW2Bytes + kW3Bytes;

// TMEM column map of one tile (256 columns allocated)
constexpr uint32_t kColD = 0;      // [0,128)   fp32 accumulator of layers 1 and 2
constexpr uint32_t kColA = 128;    // [128,192) bf16x2 activations (next layer's A, K = 128)
constexpr uint32_t kColX = 192;    // [192,200) bf16x2 network input (K = 16)
constexpr uint32_t kColD3 = 224;   // [224,240) fp32 accumulator of the output layer (N = 16)
constexpr uint32_t kTmemCols = 256;

struct MlpParams {
    const void *wblob;        // device, kWBlobBytes, canonical layouts W1 | W2 | W3
    const float *fvec;        // device: b1[128] b2[128] b3[16] xmean[16] xinvstd[16] ystd[16] ymean[16]
    int s, a;                 // state / action dims (s + a <= 16, s <= 16)
};
constexpr int kFvecFloats = 128 + 128 + 16 * 5;

__host__ __device__ inline int canon_offset_bytes(int n, int k, int N)
{
    return ((k >> 3) * (N >> 3) + (n >> 3)) * 128 + (n & 7) * 16 + (k & 7) * 2;
}

// ---- tcgen05 / TMEM primitives ------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor kind::f16: D = F32, A = B = BF16, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]^T, one K = 16 step (A from tensor memory)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 bit: thread i of the warp <-> TMEM lane (lane_base + i); N consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// pack two floats to bf16x2 with ReLU folded into the conversion; `lo` lands in the low half (lower k)
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// Shared-memory / TMEM context of one CTA tile.
struct MlpTile {
    uint32_t tmem;            // TMEM base address (lane 0, column 0 of the allocation)
    uint32_t lane_addr;       // tmem + (32 * (warp % 4)) << 16 : this warp's lane quadrant
    uint32_t sW;              // shared address of the weight blob
    const float *fvec;        // shared: biases and normalisation vectors
    uint64_t *mma_bar;        // mbarrier the MMA commits arrive on
    uint32_t phase;           // its current parity
};

// Issue one dense layer: D[d_col] = A[a_col, K] * W^T  (N outputs), by ONE thread; completion on mma_bar.
__device__ __forceinline__ void mlp_issue_layer(const MlpTile &t, uint32_t d_col, uint32_t a_col, uint32_t w_off, int N, int K)
{
    const uint32_t idesc = make_idesc(128, N);
    const uint32_t lbo = (uint32_t)(N >> 3) * 128u, sbo = 128u;      // adjacent k-groups / adjacent n-groups
    tc_fence_after();
    for (int k = 0; k < K / 16; k++) {
        const uint64_t bdesc = make_smem_desc(t.sW + w_off + (uint32_t)k * 2u * lbo, lbo, sbo);
        umma_ts(t.tmem + d_col, t.tmem + a_col + (uint32_t)k * 8u, bdesc, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(t.mma_bar);
}

// Hidden-layer epilogue for this thread's row: A[kColA] = bf16(relu(D[kColD] + bias))
__device__ __forceinline__ void mlp_hidden_epilogue(const MlpTile &t, const float *bias)
{
#pragma unroll 1
    for (int c = 0; c < kMlpH / 32; c++) {
        uint32_t v[32];
        tmem_ld32(t.lane_addr + kColD + 32u * c, v);
        tc_wait_ld();
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const float2 b = *reinterpret_cast<const float2 *>(bias + 32 * c + 2 * i);
            o[i] = pack_relu_bf16x2(__uint_as_float(v[2 * i]) + b.x, __uint_as_float(v[2 * i + 1]) + b.y);
        }
        tmem_st16(t.lane_addr + kColA + 16u * c, o);
    }
    tc_wait_st();
}

// One MLP step for this thread's sample row.  All kMlpThreads threads of the CTA must call it together.
// x (fp32 state, s values) is updated in place: x' = x + d * Ystd + Ymean.
template <int S, int A>
__device__ __forceinline__ void mlp_step(MlpTile &t, float (&x)[S], const float (&u)[A])
{
    static_assert(S + A <= kMlpKin && S <= kMlpNout, "MLP tile supports s + a <= 16");
    const float *b1 = t.fvec, *b2 = t.fvec + 128, *b3 = t.fvec + 256, *xmean = t.fvec + 272, *xinv = t.fvec + 288,
                *ystd = t.fvec + 304, *ymean = t.fvec + 320;
    // ---- input row: normalise, pack to bf16, store as the K = 16 A operand -----------------------
    {
        float in[kMlpKin];
#pragma unroll
        for (int i = 0; i < kMlpKin; i++) in[i] = 0.f;
#pragma unroll
        for (int i = 0; i < S; i++) in[i] = (x[i] - xmean[i]) * xinv[i];
#pragma unroll
        for (int i = 0; i < A; i++) in[S + i] = (u[i] - xmean[S + i]) * xinv[S + i];
        uint32_t px[8];
#pragma unroll
        for (int i = 0; i < 8; i++) px[i] = pack_bf16x2(in[2 * i], in[2 * i + 1]);
        tmem_st8(t.lane_addr + kColX, px);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) mlp_issue_layer(t, kColD, kColX, 0, kMlpH, kMlpKin);
    mbar_wait(t.mma_bar, t.phase);
    t.phase ^= 1;
    tc_fence_after();
    mlp_hidden_epilogue(t, b1);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) mlp_issue_layer(t, kColD, kColA, kW1Bytes, kMlpH, kMlpH);
    mbar_wait(t.mma_bar, t.phase);
    t.phase ^= 1;
    tc_fence_after();
    mlp_hidden_epilogue(t, b2);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) mlp_issue_layer(t, kColD3, kColA, kW1Bytes + kW2Bytes, kMlpNout, kMlpH);
    mbar_wait(t.mma_bar, t.phase);
    t.phase ^= 1;
    tc_fence_after();
    {
        uint32_t v[16];
        tmem_ld16(t.lane_addr + kColD3, v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < S; i++) x[i] += fmaf(__uint_as_float(v[i]) + b3[i], ystd[i], ymean[i]);
    }
}

// CTA prologue: TMEM allocation, mbarriers, TMA-staged weights, vectors.  smem_w must be 128-B aligned
// and hold kWBlobBytes; smem_f holds kFvecFloats floats.
__device__ __forceinline__ void mlp_tile_init(MlpTile &t, const MlpParams &mp, uint8_t *smem_w, float *smem_f,
                                              uint64_t *bars /*[2]*/, uint32_t *tmem_slot)
{
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);     // weights landed
        mbar_init(&bars[1], 1);     // MMA commits
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < kFvecFloats; i += blockDim.x) smem_f[i] = mp.fvec[i];
    if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], kWBlobBytes);
        bulk_g2s(smem_w, mp.wblob, kWBlobBytes, &bars[0]);
    }
    t.tmem = *tmem_slot;
    t.lane_addr = t.tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    t.sW = smem_u32(smem_w);
    t.fvec = smem_f;
    t.mma_bar = &bars[1];
    t.phase = 0;
    mbar_wait(&bars[0], 0);         // every thread observes the weights (async-proxy writes) before any MMA
}

__device__ __forceinline__ void mlp_tile_fini(MlpTile &t)
{
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc(t.tmem, kTmemCols);
}

}  // namespace mppi
