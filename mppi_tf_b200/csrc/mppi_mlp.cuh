// Learned-MLP dynamics on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Model (row A13; behaviour of /root/reference/scripts/src/models/nn_model.py:54-60,215-239,289-304
// re-shaped to BASELINE config 4):   X = (concat(x,u) - Xmean)/Xstd  ->  h1 = relu(W1^T X + b1)
//   -> h2 = relu(W2^T h1 + b2) -> d = W3^T h2 + b3 -> x' = x + d*Ystd + Ymean.
//
// One evaluation of the network for a tile of 128 samples is a chain of three small GEMMs with a
// TMEM -> registers -> TMEM conversion between them; the chain is latency-bound (about 3000 cycles
// per step for 700 cycles of tensor-pipe time, measured), so the design goal is TILES IN FLIGHT per
// SM, and the scarce resource is TMEM (512 columns).  A tile is squeezed into 128 columns:
//     [0,64)    A operand: bf16x2 activations of the layer being consumed (K = 128), two 32-column
//               slots Q = K steps 0-3, P = K steps 4-7; the K = 16 network input X aliases P[0,8)
//     [64,128)  D: fp32 accumulator of ONE N = 64 half of a layer (the output layer uses D[0,16))
// so one CTA per SM keeps FOUR tiles in flight.  Each layer runs as two N = 64 halves through the
// same D columns: MMA half 0 -> row warps load D (and release it) -> MMA half 1 runs while the row
// warps convert half 0.  Activations are written back only when the MMAs that read the previous
// ones have completed (half 0 of layer 2 waits in registers).
//
// The CTA is warp-specialised: 16 ROW warps (tile i = warps 4i..4i+3, one per TMEM lane quadrant;
// thread r of a tile owns sample row r for the whole rollout, fp32 state in registers) and 4 MMA warps
// (one per tile; all lanes follow the mbarriers, one elected lane issues tcgen05.mma / commit).
// Operands: A from TMEM (written by the row threads with tcgen05.st), B = bf16 weights in the
// canonical K-major layout in shared memory, staged once per CTA by a TMA bulk copy.  b1 is folded
// into the K = 16 input GEMM (constant-1 input column), b2 is added in the epilogue, b3 is folded
// into the output affine map.  No __syncthreads on the step path.
#pragma once
#include <cuda_bf16.h>

#include "mppi_device.cuh"

namespace mppi {

constexpr int kMlpH = 128;        // hidden width (both hidden layers)
constexpr int kMlpKin = 16;       // input features (+ constant 1) padded to one UMMA K step
constexpr int kMlpNout = 16;      // output features padded to the minimum N for M = 128
constexpr int kMlpRows = 128;     // samples per tile = TMEM lanes
constexpr int kMlpTiles = 4;      // tiles in flight per CTA (4 x 128 TMEM columns)
constexpr int kMlpRowWarps = 4 * kMlpTiles;
constexpr int kMlpRowThreads = 32 * kMlpRowWarps;
constexpr int kMlpThreads = kMlpRowThreads + 32 * kMlpTiles;   // + one MMA warp per tile

// shared-memory weight blob (bytes), canonical K-major no-swizzle core-matrix layout:
//   element (n, k) of B[N x K] at ((k/8)*(N/8) + n/8)*128 + (n%8)*16 + (k%8)*2
constexpr int kW1Bytes = kMlpH * kMlpKin * 2;      //  4 KB   N = 128, K = 16 (row s+a = b1)
constexpr int kW2Bytes = kMlpH * kMlpH * 2;        // 32 KB   N = 128, K = 128
constexpr int kW3Bytes = kMlpNout * kMlpH * 2;     //  4 KB   N = 16,  K = 128
constexpr int kWBlobBytes = kW1Bytes + kW2Bytes + kW3Bytes;

// TMEM column map of one tile (tile i at column 128 i of the CTA's 512-column allocation)
constexpr uint32_t kColP = 0;      // [0,32)   A operand, K steps 4-7;  network input X = [0,8)
constexpr uint32_t kColQ = 32;     // [32,64)  A operand, K steps 0-3
constexpr uint32_t kColD = 64;     // [64,128) fp32 accumulator of one N = 64 half / of the output layer
constexpr uint32_t kTileCols = 128;
constexpr uint32_t kTmemCols = kTileCols * kMlpTiles;

struct MlpParams {
    const void *wblob;        // device, kWBlobBytes, canonical layouts W1 | W2 | W3
    const float *fvec;        // device: b2[128] xmean[16] xinvstd[16] ystd[16] yc[16] (yc = b3*ystd + ymean)
    int s, a;                 // state / action dims (s + a <= 15, s <= 16)
};
constexpr int kFvecFloats = 128 + 16 * 4;

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
// developer build only: clock stamps of every hand-over of CTA 0, tile 0 (row warp 0 / MMA warp 16).
__device__ long long g_mlp_trace[2][16 * 64];
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

// Register re-partitioning between the warp groups (setmaxnreg): the kernel launches with 96 registers
// per thread (640 threads); the MMA warps give theirs up, the row warps - which hold a converted
// half layer in registers while the next accumulator half streams in - take them.
constexpr int kMlpRowRegs = 112, kMlpMmaRegs = 32;    // pool = 640 * 96; the MMA warps release 128 * 64 = 512 * 16
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// mbarriers of one tile
enum MlpBar : int {
    kBarXR = 0,    // network input stored (and the output accumulator read)   row warps -> MMA warp (count 4)
    kBarDF,        // accumulator half read: D may be overwritten              row warps -> MMA warp (count 4)
    kBarAR,        // activations of a whole layer stored                      row warps -> MMA warp (count 4)
    kBarDfull,     // accumulator complete                                     MMA commit -> row warps (count 1)
    kMlpBarsPerTile
};
constexpr int kMlpNumBars = 1 + kMlpBarsPerTile * kMlpTiles;   // [0] = weights landed

// Shared-memory / TMEM context of one tile, as seen by one of its threads.
struct MlpTile {
    uint32_t tmem;            // TMEM address of the tile: lane 0, first column of the tile
    uint32_t lane_addr;       // tmem + (32 * (warp % 4)) << 16 : this row warp's lane quadrant
    uint32_t sW;              // shared address of the weight blob
    const float *fvec;        // shared: b2 and the normalisation vectors
    uint64_t *bars;           // this tile's [kMlpBarsPerTile]
    uint32_t ph;              // row warps: parity of the next Dfull completion
#ifdef MPPI_MLP_TRACE
    int tr_n;
    bool tr_on;
#endif
};

// ---- MMA warp -------------------------------------------------------------------------------------
// One N-wide slab of a K = 128 layer: D = A * W[rows n0 .. n0+N)^T as eight K = 16 steps.  A: K steps
// 0-3 in slot Q, 4-7 in slot P; weights: canonical layout of an [N_total x 128] matrix at w_off.
__device__ __forceinline__ void mlp_issue_hidden(const MlpTile &t, uint32_t w_off, int n0, int N, int N_total)
{
    const uint32_t idesc = make_idesc(128, N);
    const uint32_t lbo = (uint32_t)(N_total >> 3) * 128u, sbo = 128u;   // adjacent k-groups / adjacent n-groups
    const uint32_t w = t.sW + w_off + (uint32_t)(n0 >> 3) * 128u;
#pragma unroll
    for (int k = 0; k < kMlpH / 16; k++) {
        const uint64_t bdesc = make_smem_desc(w + (uint32_t)k * 2u * lbo, lbo, sbo);
        const uint32_t a_col = (k < 4 ? kColQ + 8u * k : kColP + 8u * (k - 4));
        umma_ts(t.tmem + kColD, t.tmem + a_col, bdesc, idesc, k > 0 ? 1u : 0u);
    }
}
// One N = 64 half of the input layer (K = 16, A = X in P[0,8))
__device__ __forceinline__ void mlp_issue_input(const MlpTile &t, int n0)
{
    const uint32_t lbo = (uint32_t)(kMlpH >> 3) * 128u, sbo = 128u;
    const uint64_t bdesc = make_smem_desc(t.sW + (uint32_t)(n0 >> 3) * 128u, lbo, sbo);
    umma_ts(t.tmem + kColD, t.tmem + kColP, bdesc, make_idesc(128, 64), 0u);
}

// The whole MMA side of `nsteps` network evaluations of one tile.  Called by the WHOLE MMA warp
// (converged): every lane follows the mbarriers, one elected lane issues the tcgen05 instructions.
__device__ __forceinline__ void mlp_mma_loop(MlpTile &t, int nsteps)
{
    uint32_t ph_x = 0, ph_f = 0, ph_a = 0;
    uint64_t *full = &t.bars[kBarDfull];
    for (int s = 0; s < nsteps; s++) {
        MLP_TRACE(1, 0);
        mbar_wait(&t.bars[kBarXR], ph_x);            // X stored; D (previous output) read
        ph_x ^= 1;
        tc_fence_after();
        MLP_TRACE(1, 1);
        if (elect_one()) { mlp_issue_input(t, 0); umma_commit(full); }
        __syncwarp();
        MLP_TRACE(1, 2);
        mbar_wait(&t.bars[kBarDF], ph_f);            // layer-1 half 0 read
        ph_f ^= 1;
        tc_fence_after();
        MLP_TRACE(1, 3);
        if (elect_one()) { mlp_issue_input(t, 64); umma_commit(full); }
        __syncwarp();
        MLP_TRACE(1, 4);
        mbar_wait(&t.bars[kBarAR], ph_a);            // A1 stored (implies layer-1 half 1 read)
        ph_a ^= 1;
        tc_fence_after();
        MLP_TRACE(1, 5);
        if (elect_one()) { mlp_issue_hidden(t, kW1Bytes, 0, 64, kMlpH); umma_commit(full); }
        __syncwarp();
        MLP_TRACE(1, 6);
        mbar_wait(&t.bars[kBarDF], ph_f);            // layer-2 half 0 read
        ph_f ^= 1;
        tc_fence_after();
        MLP_TRACE(1, 7);
        if (elect_one()) { mlp_issue_hidden(t, kW1Bytes, 64, 64, kMlpH); umma_commit(full); }
        __syncwarp();
        MLP_TRACE(1, 8);
        mbar_wait(&t.bars[kBarAR], ph_a);            // A2 stored (implies layer-2 half 1 read)
        ph_a ^= 1;
        tc_fence_after();
        MLP_TRACE(1, 9);
        if (elect_one()) { mlp_issue_hidden(t, kW1Bytes + kW2Bytes, 0, kMlpNout, kMlpNout); umma_commit(full); }
        __syncwarp();
        MLP_TRACE(1, 10);
    }
}

// ---- row warps ------------------------------------------------------------------------------------
__device__ __forceinline__ void mlp_row_signal(const MlpTile &t, int bar)
{
    tc_fence_before();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&t.bars[bar]);
}
__device__ __forceinline__ void mlp_row_wait_full(MlpTile &t)
{
    mbar_wait(&t.bars[kBarDfull], t.ph);
    t.ph ^= 1;
    tc_fence_after();
}
// the 64 fp32 accumulator columns -> registers (two 32-column loads in flight)
__device__ __forceinline__ void mlp_load_d(const MlpTile &t, uint32_t (&v0)[32], uint32_t (&v1)[32])
{
    tmem_ld32(t.lane_addr + kColD, v0);
    tmem_ld32(t.lane_addr + kColD + 32u, v1);
    tc_wait_ld();
}
// (+ bias) -> ReLU -> bf16x2: 64 accumulator values to 32 packed words
template <bool BIAS>
__device__ __forceinline__ void mlp_pack(const uint32_t (&v0)[32], const uint32_t (&v1)[32], const float *bias, uint32_t (&o)[32])
{
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float2 p0 = make_float2(__uint_as_float(v0[4 * i]), __uint_as_float(v0[4 * i + 1]));
        float2 p1 = make_float2(__uint_as_float(v0[4 * i + 2]), __uint_as_float(v0[4 * i + 3]));
        float2 p2 = make_float2(__uint_as_float(v1[4 * i]), __uint_as_float(v1[4 * i + 1]));
        float2 p3 = make_float2(__uint_as_float(v1[4 * i + 2]), __uint_as_float(v1[4 * i + 3]));
        if (BIAS) {
            const float4 b0 = *reinterpret_cast<const float4 *>(bias + 4 * i);        // warp-uniform address: broadcast
            const float4 b1 = *reinterpret_cast<const float4 *>(bias + 32 + 4 * i);
            p0 = __fadd2_rn(p0, make_float2(b0.x, b0.y));
            p1 = __fadd2_rn(p1, make_float2(b0.z, b0.w));
            p2 = __fadd2_rn(p2, make_float2(b1.x, b1.y));
            p3 = __fadd2_rn(p3, make_float2(b1.z, b1.w));
        }
        o[2 * i] = pack_relu_bf16x2(p0.x, p0.y);
        o[2 * i + 1] = pack_relu_bf16x2(p1.x, p1.y);
        o[16 + 2 * i] = pack_relu_bf16x2(p2.x, p2.y);
        o[16 + 2 * i + 1] = pack_relu_bf16x2(p3.x, p3.y);
    }
}

// Step part 1: normalise (x, u), store the bf16 input row into P[0,8), hand it to the MMA warp.
// Precondition: the previous output accumulator has been read (mlp_row_finish) - XR also releases D.
template <int S, int A>
__device__ __forceinline__ void mlp_row_begin(MlpTile &t, const float (&x)[S], const float (&u)[A])
{
    static_assert(S + A + 1 <= kMlpKin && S <= kMlpNout, "MLP tile supports s + a <= 15");
    const float *xmean = t.fvec + 128, *xinv = t.fvec + 144;
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
    tmem_st8(t.lane_addr + kColP, px);
    tc_wait_st();
    mlp_row_signal(t, kBarXR);
    MLP_TRACE(0, 1);
}
// Step part 2: layer-1 epilogue.  Half 0 goes to Q at once (layer 1 reads only X = P[0,8)); half 1
// overwrites P once the second input MMA has completed.
__device__ __forceinline__ void mlp_row_layer1(MlpTile &t)
{
    uint32_t v0[32], v1[32], o[32];
    MLP_TRACE(0, 2);
    mlp_row_wait_full(t);
    MLP_TRACE(0, 3);
    mlp_load_d(t, v0, v1);
    mlp_row_signal(t, kBarDF);
    MLP_TRACE(0, 4);
    mlp_pack<false>(v0, v1, nullptr, o);
    tmem_st32(t.lane_addr + kColQ, o);
    MLP_TRACE(0, 5);
    mlp_row_wait_full(t);
    MLP_TRACE(0, 6);
    mlp_load_d(t, v0, v1);
    mlp_pack<false>(v0, v1, nullptr, o);
    tmem_st32(t.lane_addr + kColP, o);
    tc_wait_st();
    mlp_row_signal(t, kBarAR);
    MLP_TRACE(0, 7);
}
// Step part 3: layer-2 epilogue.  Half 0 waits in registers: the MMAs of half 1 still read A1.
__device__ __forceinline__ void mlp_row_layer2(MlpTile &t)
{
    uint32_t v0[32], v1[32], o0[32], o1[32];
    const float *b2 = t.fvec;
    MLP_TRACE(0, 8);
    mlp_row_wait_full(t);
    MLP_TRACE(0, 9);
    mlp_load_d(t, v0, v1);
    mlp_row_signal(t, kBarDF);
    MLP_TRACE(0, 10);
    mlp_pack<true>(v0, v1, b2, o0);
    MLP_TRACE(0, 11);
    mlp_row_wait_full(t);
    MLP_TRACE(0, 12);
    tmem_st32(t.lane_addr + kColQ, o0);
    mlp_load_d(t, v0, v1);
    mlp_pack<true>(v0, v1, b2 + 64, o1);
    tmem_st32(t.lane_addr + kColP, o1);
    tc_wait_st();
    mlp_row_signal(t, kBarAR);
    MLP_TRACE(0, 13);
}
// Step part 4: x' = x + d * Ystd + (b3 * Ystd + Ymean)
template <int S>
__device__ __forceinline__ void mlp_row_finish(MlpTile &t, float (&x)[S])
{
    const float *ystd = t.fvec + 160, *yc = t.fvec + 176;
    MLP_TRACE(0, 14);
    mlp_row_wait_full(t);
    MLP_TRACE(0, 15);
    uint32_t v[16];
    tmem_ld16(t.lane_addr + kColD, v);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < S; i++) x[i] += fmaf(__uint_as_float(v[i]), ystd[i], yc[i]);
}

// CTA prologue: TMEM allocation, mbarriers, TMA-staged weights, vectors.  smem_w must be 128-B aligned
// and hold kWBlobBytes; smem_f holds kFvecFloats floats.  Ends with a CTA barrier.
__device__ __forceinline__ void mlp_tile_init(MlpTile &t, const MlpParams &mp, uint8_t *smem_w, float *smem_f,
                                              uint64_t *bars /*[kMlpNumBars]*/, uint32_t *tmem_slot)
{
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        for (int i = 0; i < kMlpTiles; i++) {
            uint64_t *b = bars + 1 + i * kMlpBarsPerTile;
            mbar_init(&b[kBarXR], 4);
            mbar_init(&b[kBarDF], 4);
            mbar_init(&b[kBarAR], 4);
            mbar_init(&b[kBarDfull], 1);
        }
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < kFvecFloats; i += blockDim.x) smem_f[i] = mp.fvec[i];
    if (warp == kMlpRowWarps) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], kWBlobBytes);
        bulk_g2s(smem_w, mp.wblob, kWBlobBytes, &bars[0]);
    }
    const int tile = warp < kMlpRowWarps ? (warp >> 2) : (warp - kMlpRowWarps);
    t.tmem = *tmem_slot + kTileCols * (uint32_t)tile;
    t.lane_addr = t.tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    t.sW = smem_u32(smem_w);
    t.fvec = smem_f;
    t.bars = bars + 1 + tile * kMlpBarsPerTile;
    t.ph = 0;
#ifdef MPPI_MLP_TRACE
    t.tr_n = 0;
    t.tr_on = blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (warp == 0 || warp == kMlpRowWarps);
#endif
    mbar_wait(&bars[0], 0);         // every thread observes the weights (async-proxy writes) before any MMA
    __syncthreads();
}

// All MMAs have been consumed (the row warps waited for their last output accumulator) when this is called.
__device__ __forceinline__ void mlp_tile_fini(uint32_t tmem_base)
{
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == kMlpRowWarps) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace mppi
