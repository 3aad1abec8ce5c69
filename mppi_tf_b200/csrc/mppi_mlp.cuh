// Learned-MLP dynamics on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Model (row A13; behaviour of /root/reference/scripts/src/models/nn_model.py:54-60,215-239,289-304
// re-shaped to BASELINE config 4):   X = (concat(x,u) - Xmean)/Xstd  ->  h1 = relu(W1^T X + b1)
//   -> h2 = relu(W2^T h1 + b2) -> d = W3^T h2 + b3 -> x' = x + d*Ystd + Ymean.
//
// Mapping: one CTA tile = 128 samples = the 128 TMEM lanes (UMMA_M = 128, cta_group::1).  The CTA is
// warp-specialised: warps 0-3 are the ROW warps (thread r owns sample row r for the whole rollout,
// fp32 state in registers), warp 4 is the MMA warp (one lane issues every tcgen05.mma).  Each layer is
//   D[128 x N] (fp32, TMEM) = A[128 x K] (bf16, TMEM, written by the row threads with tcgen05.st)
//                             x B[N x K]^T (bf16 weights, K-major canonical layout in shared memory,
//                                           staged once per CTA by a TMA bulk copy)
// and is issued as two N = 64 halves, each committed to its own mbarrier, so that the row warps
// convert half 0 (TMEM -> F2FP.RELU -> TMEM) while the tensor pipe is still producing half 1, and
// the next layer's first K steps start as soon as half 0 of its A operand is in place.  No
// __syncthreads on the step path: row warps and the MMA warp hand tiles over through mbarriers only.
// Two CTAs (2 x 256 TMEM columns) are resident per SM and fill each other's tensor-pipe gaps.
#pragma once
#include <cuda_bf16.h>

#include "mppi_device.cuh"

namespace mppi {

constexpr int kMlpH = 128;        // hidden width (both hidden layers)
constexpr int kMlpKh = 144;       // hidden K padded by one UMMA K step: column 128 is a constant 1 that
                                  // multiplies the bias row of the next layer's weights (bias folded
                                  // into the GEMM: no per-element bias add in the epilogue)
constexpr int kMlpKin = 16;       // input features (+ constant 1) padded to one UMMA K step
constexpr int kMlpNout = 16;      // output features padded to the minimum N for M = 128
constexpr int kMlpRows = 128;     // samples per tile = TMEM lanes
#ifndef MPPI_MLP_SPLIT_N
#define MPPI_MLP_SPLIT_N 1
#endif
constexpr bool kMlpSplitN = MPPI_MLP_SPLIT_N != 0;   // layer 2 as two N = 64 halves (1) or one N = 128 MMA per K step (0)
constexpr int kMlpRowWarps = 4;    // one per TMEM lane quadrant: thread = sample row (state, noise, cost)
constexpr int kMlpEpiWarps = 8;    // warps 0-3 (rows) and 4-7 (helpers): two threads per row share every conversion,
                                   // each taking 32 of the 64 accumulator columns of an N half
constexpr int kMlpMmaWarp = kMlpEpiWarps;
constexpr int kMlpThreads = 32 * (kMlpEpiWarps + 1);

// shared-memory weight blob (bytes), canonical K-major no-swizzle core-matrix layout:
//   element (n, k) of B[N x K] at ((k/8)*(N/8) + n/8)*128 + (n%8)*16 + (k%8)*2
constexpr int kW1Bytes = kMlpH * kMlpKin * 2;      //  4 KB   N = 128, K = 16
constexpr int kW2Bytes = kMlpH * kMlpKh * 2;       // 36 KB   N = 128, K = 144 (row 128 = b2)
constexpr int kW3Bytes = kMlpNout * kMlpKh * 2;    // 4.5 KB  N = 16,  K = 144 (row 128 = b3)
constexpr int kWBlobBytes = kW1Bytes + kW2Bytes + kW3Bytes;

// TMEM column map of one tile (256 columns allocated)
constexpr uint32_t kColD = 0;      // [0,128)   fp32 accumulator of layers 1 and 2, as two N = 64 halves
constexpr uint32_t kColA = 128;    // [128,200) bf16x2 activations (next layer's A, K = 144; cols 192.. = const 1, 0..)
constexpr uint32_t kColX = 200;    // [200,208) bf16x2 network input (K = 16)
constexpr uint32_t kColD3 = 224;   // [224,240) fp32 accumulator of the output layer (N = 16)
constexpr uint32_t kTmemCols = 256;

struct MlpParams {
    const void *wblob;        // device, kWBlobBytes, canonical layouts W1 | W2 | W3
    const float *fvec;        // device: xmean[16] xinvstd[16] ystd[16] ymean[16] (biases live in the weight blob)
    int s, a;                 // state / action dims (s + a <= 16, s <= 16)
};
constexpr int kFvecFloats = 16 * 4;

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

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#ifdef MPPI_MLP_TRACE
// developer build only: clock stamps of every hand-over of CTA 0 (row warp 0 / MMA warp), first steps.
// The event index lives in a register (fire-and-forget stores: a few issue slots per stamp).
__device__ long long g_mlp_trace[2][16 * 64];
__device__ int g_mlp_trace_n[2];
#define MLP_TRACE(who, ev)                                                                         \
    do {                                                                                           \
        if (t.tr_on && t.tr_n < 16 * 64) {                                                         \
            g_mlp_trace[who][t.tr_n] = ((long long)clock64() << 8) | (ev);                         \
            t.tr_n++;                                                                              \
        }                                                                                          \
    } while (0)
#else
#define MLP_TRACE(who, ev)
#endif

// Register re-partitioning between the roles (setmaxnreg, warp-group granularity): the kernel launches with 96
// registers per thread; during the rollout the row warps take what the helper and MMA warps do not need.
constexpr int kMlpRegsLaunch = 96, kMlpRegsRow = 136, kMlpRegsHelper = 56, kMlpRegsMma = 40;
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// mbarriers of one CTA tile
enum MlpBar : int {
    kBarW = 0,     // weights landed (TMA complete_tx)
    kBarX,         // network input row block stored          row warps -> MMA warp   (count 4)
    kBarA0,        // activation K half 0 stored              row warps -> MMA warp   (count 4)
    kBarA1,        // activation K half 1 stored
    kBarD0,        // accumulator N half 0 complete           MMA commit -> row warps (count 1)
    kBarD1,        // accumulator N half 1 complete
    kBarD3,        // output-layer accumulator complete
    kMlpNumBars
};

// Shared-memory / TMEM context of one CTA tile.
struct MlpTile {
    uint32_t tmem;            // TMEM base address (lane 0, column 0 of the allocation)
    uint32_t lane_addr;       // tmem + (32 * (warp % 4)) << 16 : this warp's lane quadrant
    uint32_t sW;              // shared address of the weight blob
    const float *fvec;        // shared: normalisation vectors
    uint64_t *bars;           // [kMlpNumBars]
    uint32_t ph_d, ph_d3;     // row warps: parity of the next D0/D1 and D3 completion
#ifdef MPPI_MLP_TRACE
    int tr_n;
    bool tr_on;
#endif
};

// ---- MMA warp -------------------------------------------------------------------------------------
// K = 16 steps [k0, k1) of D[d_col, N] (+)= A[a_col ...] * W[w_off ...]^T; weights in canonical layout
// of a matrix with N_total rows (the N half is selected through w_off).
__device__ __forceinline__ void mlp_issue(const MlpTile &t, uint32_t d_col, uint32_t a_col, uint32_t w_off, int N, int N_total,
                                          int k0, int k1, bool fresh)
{
    const uint32_t idesc = make_idesc(128, N);
    const uint32_t lbo = (uint32_t)(N_total >> 3) * 128u, sbo = 128u;   // adjacent k-groups / adjacent n-groups
#pragma unroll
    for (int k = k0; k < k1; k++) {
        const uint64_t bdesc = make_smem_desc(t.sW + w_off + (uint32_t)k * 2u * lbo, lbo, sbo);
        umma_ts(t.tmem + d_col, t.tmem + a_col + (uint32_t)k * 8u, bdesc, idesc, (fresh && k == k0) ? 0u : 1u);
    }
}

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// The whole MMA side of `nsteps` network evaluations of this CTA's tile.  Called by the WHOLE MMA warp
// (converged): every lane follows the mbarriers, one elected lane issues the tcgen05 instructions
// (a warp-uniform region keeps the uniform-datapath operands of UTCHMMA free of per-thread loops).
__device__ __forceinline__ void mlp_mma_loop(MlpTile &t, int nsteps)
{
    constexpr int KH = kMlpKh / 16, KH0 = kMlpH / 32;    // 9 K steps per hidden layer, 4 of them in K half 0
    constexpr uint32_t kHalfB = (kMlpH / 2 / 8) * 128u;  // byte offset of rows 64.. in a canonical N = 128 matrix
    uint32_t ph_x = 0, ph_a = 0;
    for (int s = 0; s < nsteps; s++) {
        // layer 1: X[128 x 16] -> D (two N halves)
        MLP_TRACE(1, 0);
        mbar_wait_spin(&t.bars[kBarX], ph_x);
        MLP_TRACE(1, 1);
        ph_x ^= 1;
        tc_fence_after();
        if (elect_one()) {
            mlp_issue(t, kColD, kColX, 0, 64, kMlpH, 0, 1, true);
            umma_commit(&t.bars[kBarD0]);
            mlp_issue(t, kColD + 64, kColX, kHalfB, 64, kMlpH, 0, 1, true);
            umma_commit(&t.bars[kBarD1]);
        }
        __syncwarp();
        // layer 2: starts on K half 0 of A1 while the row warps still convert half 1
        MLP_TRACE(1, 2);
        mbar_wait_spin(&t.bars[kBarA0], ph_a);
        MLP_TRACE(1, 3);
        tc_fence_after();
        if (elect_one()) mlp_issue(t, kColD, kColA, kW1Bytes, kMlpSplitN ? 64 : 128, kMlpH, 0, KH0, true);
        __syncwarp();
        MLP_TRACE(1, 4);
        mbar_wait_spin(&t.bars[kBarA1], ph_a);
        MLP_TRACE(1, 5);
        ph_a ^= 1;
        tc_fence_after();
        if (elect_one()) {
            if (kMlpSplitN) {
                mlp_issue(t, kColD, kColA, kW1Bytes, 64, kMlpH, KH0, KH, false);
                umma_commit(&t.bars[kBarD0]);
                mlp_issue(t, kColD + 64, kColA, kW1Bytes + kHalfB, 64, kMlpH, 0, KH, true);
                umma_commit(&t.bars[kBarD1]);
            } else {
                mlp_issue(t, kColD, kColA, kW1Bytes, 128, kMlpH, KH0, KH, false);
                umma_commit(&t.bars[kBarD0]);
                umma_commit(&t.bars[kBarD1]);
            }
        }
        __syncwarp();
        // output layer: one batch (nine tiny MMAs) once both K halves of A2 are in place - an early
        // start on half 0 would save 50 tensor cycles and cost a second issue batch (~200 cycles)
        MLP_TRACE(1, 6);
        mbar_wait_spin(&t.bars[kBarA0], ph_a);
        MLP_TRACE(1, 8);
        mbar_wait_spin(&t.bars[kBarA1], ph_a);
        MLP_TRACE(1, 9);
        ph_a ^= 1;
        tc_fence_after();
        if (elect_one()) {
            mlp_issue(t, kColD3, kColA, kW1Bytes + kW2Bytes, kMlpNout, kMlpNout, 0, KH, true);
            umma_commit(&t.bars[kBarD3]);
        }
        __syncwarp();
        MLP_TRACE(1, 10);
    }
}

// ---- row warps ------------------------------------------------------------------------------------
// this thread's 32 columns (`part` 0 / 1) of accumulator N half `h` -> 16 packed bf16x2 words, ReLU folded
// into the conversion (the bias already sits in the accumulator: K-augmented GEMM)
__device__ __forceinline__ void mlp_load_pack(const MlpTile &t, int h, int part, uint32_t (&o)[16])
{
    uint32_t v[32];
    tmem_ld32(t.lane_addr + kColD + 64u * h + 32u * part, v);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; i++) o[i] = pack_relu_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
}
// store this thread's 16 columns of K half `h` of the next A operand and hand them to the MMA warp
__device__ __forceinline__ void mlp_store_signal(const MlpTile &t, int h, int part, const uint32_t (&o)[16])
{
    tmem_st16(t.lane_addr + kColA + 32u * h + 16u * part, o);
    tc_wait_st();
    tc_fence_before();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&t.bars[h ? kBarA1 : kBarA0]);
}

// Step part 1: normalise (x, u), store the bf16 input row, hand it to the MMA warp.
template <int S, int A>
__device__ __forceinline__ void mlp_row_begin(MlpTile &t, const float (&x)[S], const float (&u)[A])
{
    static_assert(S + A + 1 <= kMlpKin && S <= kMlpNout, "MLP tile supports s + a <= 15");
    const float *xmean = t.fvec, *xinv = t.fvec + 16;
    float in[kMlpKin];
#pragma unroll
    for (int i = 0; i < kMlpKin; i++) in[i] = 0.f;
#pragma unroll
    for (int i = 0; i < S; i++) in[i] = (x[i] - xmean[i]) * xinv[i];
#pragma unroll
    for (int i = 0; i < A; i++) in[S + i] = (u[i] - xmean[S + i]) * xinv[S + i];
    in[S + A] = 1.0f;                      // multiplies the b1 row of W1
    uint32_t px[8];
#pragma unroll
    for (int i = 0; i < 8; i++) px[i] = pack_bf16x2(in[2 * i], in[2 * i + 1]);
    tmem_st8(t.lane_addr + kColX, px);
    tc_wait_st();
    tc_fence_before();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&t.bars[kBarX]);
    MLP_TRACE(0, 1);
}
// Step part 2: layer-1 epilogue (the A region is dead here: the previous output layer has completed).
// Called by the row thread (part 0) and by its helper (part 1).
__device__ __forceinline__ void mlp_row_layer1(MlpTile &t, int part)
{
    uint32_t o[16];
    MLP_TRACE(0, 2);
    mbar_wait_spin(&t.bars[kBarD0], t.ph_d);
    MLP_TRACE(0, 3);
    tc_fence_after();
    mlp_load_pack(t, 0, part, o);
    MLP_TRACE(0, 4);
    mlp_store_signal(t, 0, part, o);
    MLP_TRACE(0, 5);
    mbar_wait_spin(&t.bars[kBarD1], t.ph_d);
    MLP_TRACE(0, 6);
    tc_fence_after();
    mlp_load_pack(t, 1, part, o);
    mlp_store_signal(t, 1, part, o);
    MLP_TRACE(0, 7);
    t.ph_d ^= 1;
}
// Step part 3: layer-2 epilogue.  Half 0 is converted while the tensor pipe produces half 1, but it
// is stored only after half 1 has completed: until then the MMAs still read A1 from the same columns.
__device__ __forceinline__ void mlp_row_layer2(MlpTile &t, int part)
{
    uint32_t o[16];
    MLP_TRACE(0, 8);
    mbar_wait_spin(&t.bars[kBarD0], t.ph_d);
    MLP_TRACE(0, 9);
    tc_fence_after();
    mlp_load_pack(t, 0, part, o);
    MLP_TRACE(0, 10);
    mbar_wait_spin(&t.bars[kBarD1], t.ph_d);
    MLP_TRACE(0, 11);
    tc_fence_after();
    mlp_store_signal(t, 0, part, o);
    MLP_TRACE(0, 12);
    mlp_load_pack(t, 1, part, o);
    mlp_store_signal(t, 1, part, o);
    MLP_TRACE(0, 13);
    t.ph_d ^= 1;
}
// The helper thread of a row: its share of every conversion of `nsteps` network evaluations.
__device__ __forceinline__ void mlp_helper_loop(MlpTile &t, int nsteps)
{
    for (int s = 0; s < nsteps; s++) {
        mlp_row_layer1(t, 1);
        mlp_row_layer2(t, 1);
    }
}
// Step part 4: x' = x + d * Ystd + Ymean
template <int S>
__device__ __forceinline__ void mlp_row_finish(MlpTile &t, float (&x)[S])
{
    const float *ystd = t.fvec + 32, *ymean = t.fvec + 48;
    MLP_TRACE(0, 14);
    mbar_wait_spin(&t.bars[kBarD3], t.ph_d3);
    MLP_TRACE(0, 15);
    t.ph_d3 ^= 1;
    tc_fence_after();
    uint32_t v[16];
    tmem_ld16(t.lane_addr + kColD3, v);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < S; i++) x[i] += fmaf(__uint_as_float(v[i]), ystd[i], ymean[i]);
}

// CTA prologue: TMEM allocation (MMA warp), mbarriers, TMA-staged weights, vectors.  smem_w must be
// 128-B aligned and hold kWBlobBytes; smem_f holds kFvecFloats floats.
__device__ __forceinline__ void mlp_tile_init(MlpTile &t, const MlpParams &mp, uint8_t *smem_w, float *smem_f,
                                              uint64_t *bars /*[kMlpNumBars]*/, uint32_t *tmem_slot)
{
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bars[kBarW], 1);
        mbar_init(&bars[kBarX], kMlpRowWarps);
        mbar_init(&bars[kBarA0], kMlpEpiWarps);
        mbar_init(&bars[kBarA1], kMlpEpiWarps);
        mbar_init(&bars[kBarD0], 1);
        mbar_init(&bars[kBarD1], 1);
        mbar_init(&bars[kBarD3], 1);
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < kFvecFloats; i += blockDim.x) smem_f[i] = mp.fvec[i];
    if (warp == kMlpMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[kBarW], kWBlobBytes);
        bulk_g2s(smem_w, mp.wblob, kWBlobBytes, &bars[kBarW]);
    }
    t.tmem = *tmem_slot;
    t.lane_addr = t.tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    t.sW = smem_u32(smem_w);
    t.fvec = smem_f;
    t.bars = bars;
    t.ph_d = 0;
    t.ph_d3 = 0;
#ifdef MPPI_MLP_TRACE
    t.tr_n = 0;
    t.tr_on = blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (warp == 0 || warp == kMlpMmaWarp);
#endif
    if (warp < kMlpRowWarps) {   // K-augmentation columns of the activation operand: k = 128 is the constant 1, k = 129..143 are 0
        uint32_t one[8] = {pack_bf16x2(1.0f, 0.0f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st8(t.lane_addr + kColA + kMlpH / 2, one);
        tc_wait_st();
    }
    mbar_wait(&bars[kBarW], 0);     // every thread observes the weights (async-proxy writes) before any MMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}

// All MMAs of the tile have been consumed (the row warps waited for the last D3) when this is called.
__device__ __forceinline__ void mlp_tile_fini(MlpTile &t)
{
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == kMlpMmaWarp) tmem_dealloc(t.tmem, kTmemCols);
}

}  // namespace mppi
