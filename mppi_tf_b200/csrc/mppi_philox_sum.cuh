// Phase 2 of the Philox-mode update kernels that regenerate the noise: sum_k e_k n_k over the CTA's samples, re-walking
// the Philox counters of phase 1, then the CTA partial and the last-CTA merge (mppi_update.cuh).  Shared by
// rollout_philox_kernel (mppi_rollout.cu) and rollout_philox_fast_kernel (mppi_rollout_fast.cu); `Gen::normals4`
// returns the four normals of one call as the kernel's phase 1 saw them.
#pragma once
#include "mppi_device.cuh"
#include "mppi_update.cuh"

namespace mppi {

constexpr int kPhiloxThreads = 512;
constexpr int kPhiloxCtasPerSm = 2;
constexpr int kListCap = 256;          // list entries per warp: the CTA's non-zero-weight list holds 8 iterations of samples

struct GenPlain {        // ten (or p.rounds) rounds behind a uniform branch, true standard normals
    static __device__ __forceinline__ float4 normals4(const uint4 *tab, uint32_t call, const PhiloxSample s, const RolloutParams &p)
    {
        float z[4];
        normals4_tab(tab, call, s, p, z);
        return make_float4(z[0], z[1], z[2], z[3]);
    }
};
template <int R>
struct GenFast {         // compile-time rounds, n = z / z_scale
    static __device__ __forceinline__ float4 normals4(const uint4 *tab, uint32_t call, const PhiloxSample s, const RolloutParams &p)
    {
        return normals4_fast<R>(tab, call, s, p);
    }
};

// Shared-memory carve-up of the two regenerating kernels (dynamic shared memory, 16-byte aligned base).
struct PhiloxSmem {
    uint2 *sList;       // [NW * kListCap] non-zero-weight samples of the CTA
    float *sUV;         // [T][RS] per-step table (U_t, or L_t in the superposition kernel)
    float *sAcc;        // [NW][TAp] per-warp sums
    float *sN, *sWork;  // [TAp] each
    float *sScale;      // [kMaxParts]
    float *sRed;        // [64]
    float4 *sScratch;   // [kPhiloxThreads] merge scratch
    uint4 *sTab;        // [ceil(TA/4)] per-call uniform Philox words
    __device__ __forceinline__ void carve(float *smem, int T, int RS, int TAp)
    {
        constexpr int NW = kPhiloxThreads / 32;
        sList = reinterpret_cast<uint2 *>(smem);
        sUV = smem + 2 * NW * kListCap;
        sAcc = sUV + T * RS;
        sN = sAcc + NW * TAp;
        sWork = sN + TAp;
        sScale = sWork + TAp;
        sRed = sScale + kMaxParts;
        sScratch = reinterpret_cast<float4 *>(sRed + 64);
        sTab = reinterpret_cast<uint4 *>(sScratch + kPhiloxThreads);
    }
};
inline size_t philox_smem_bytes(int A, int T, int TA)
{
    const int RS = (A + 3) & ~3, TAp = (TA + 31) & ~31, NW = kPhiloxThreads / 32;
    return sizeof(float) * ((size_t)T * RS + (size_t)NW * TAp + 2 * TAp + kMaxParts + 64) + sizeof(float4) * kPhiloxThreads +
           sizeof(uint2) * NW * kListCap + sizeof(uint4) * (size_t)((TA + 3) >> 2);
}

// Weighted noise sum + partial + finish.  CTA-wide (contains barriers); the CTA owns samples [32 w_lo + .., kend) with
// thread tid starting at kfirst = 32 w_lo + tid, stride blockDim.  beta_c / max_c: the CTA's cost range (or the fixed
// beta of a normalised update).
//
// Only samples whose weight is non-zero in fp32 are revisited (with lambda of the order of the cost spread most weights
// underflow: e_k = 0 contributes exactly nothing).  Dense CTAs (no weight can underflow) walk their own samples chunk by
// chunk; otherwise the CTA compacts its samples, in order (warp-major, then iteration), into one shared-memory list of
// (global sample index, weight) and the (32-entry group, 32-normal chunk) work items of the list are dealt to the warps
// in contiguous ranges, so that all 16 warps work even when the list is short (K/8 per rank: about 90 entries per CTA).
// Deterministic: list order and item ranges depend only on the data.
template <int A, class Gen>
__device__ __forceinline__ void philox_weighted_sum_and_finish(const RolloutParams &p, int ctrl, const float *costs, int w_lo, int kfirst,
                                                               int kend, float beta_c, float max_c, float nil, uint32_t phA,
                                                               const PhiloxSmem &sm, int vtid = -1)
{
    constexpr int NW = kPhiloxThreads / 32;
    constexpr int kstride = kPhiloxThreads;
    // vtid: the thread's index under the caller's (virtual) warp numbering, kfirst = 32 w_lo + vtid; every shared-memory row,
    // list position and work range below is a function of it, so the result does not depend on the placement
    const int tid = vtid >= 0 ? vtid : (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int TA = p.TA, TAp = (TA + 31) & ~31;
    const int ncall = (TA + 3) >> 2;
    const int nchunk = (ncall + 7) >> 3;
    float *sAcc = sm.sAcc, *sRed = sm.sRed;
    const uint4 *sTab = sm.sTab;
    float eta = 0.f;
    // CTA-uniform: can any weight of this CTA underflow at all?  (false also for NaN and for the weight
    // pass of a normalised update, whose exponents are bounded by 1/lambda)
    const bool sparse = (max_c - beta_c) * fabsf(nil) > kWeightCutLog2;
    if (!sparse) {
        for (int ch = 0; ch < nchunk; ch++) {
            float2 acc2[16];
#pragma unroll
            for (int i = 0; i < 16; i++) acc2[i] = make_float2(0.f, 0.f);
            const bool full = (ch * 8 + 8 <= ncall);    // warp-uniform: all 8 calls of the chunk exist
            for (int k = kfirst; k < kend; k += kstride) {
                const PhiloxSample ps = philox_sample(phA, (uint32_t)(p.k_offset + k));
                const float e = sample_weight(costs[k], beta_c, nil);
                const float2 e2 = make_float2(e, e);
                if (ch == 0) eta += e;
#pragma unroll
                for (int c8 = 0; c8 < 8; c8++) {
                    if (full || ch * 8 + c8 < ncall) {
                        const float4 z = Gen::normals4(sTab, (uint32_t)(ch * 8 + c8), ps, p);
                        acc2[2 * c8] = __ffma2_rn(e2, make_float2(z.x, z.y), acc2[2 * c8]);
                        acc2[2 * c8 + 1] = __ffma2_rn(e2, make_float2(z.z, z.w), acc2[2 * c8 + 1]);
                    }
                }
            }
            float acc[32];
#pragma unroll
            for (int i = 0; i < 16; i++) { acc[2 * i] = acc2[i].x; acc[2 * i + 1] = acc2[i].y; }
            const float r = warp_transpose_sum32(acc, lane);
            MPPI_CHECK(ch * 32 + lane < TAp);
            sAcc[warp * TAp + ch * 32 + lane] = r;
        }
    } else {
        float *myacc = sAcc + warp * TAp;
        for (int j = lane; j < TAp; j += 32) myacc[j] = 0.f;
        int *sCnt = reinterpret_cast<int *>(sRed);                                   // [NW] per-warp counts
        constexpr int BI = kListCap / 32;                                            // iterations per batch
        const int k_cta = 32 * w_lo;
        const int n_it_cta = (kend - k_cta + kstride - 1) / kstride;                 // CTA-uniform
        const unsigned lt = (1u << lane) - 1u;
        for (int b0 = 0; b0 < n_it_cta; b0 += BI) {
            const int kb = k_cta + b0 * kstride + 32 * warp;                         // this warp's first sample of the batch
            const int nit = min(BI, n_it_cta - b0);
            int cnt_w = 0;
#pragma unroll 1
            for (int it = 0; it < nit; it++) {
                const int k = kb + it * kstride + lane;
                float e = 0.f;
                if (k < kend) e = sample_weight(costs[k], beta_c, nil);
                eta += e;
                cnt_w += __popc(__ballot_sync(0xffffffffu, e != 0.f));
            }
            if (lane == 0) sCnt[warp] = cnt_w;
            __syncthreads();
            int off = 0, total = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const int c = sCnt[w];
                if (w < warp) off += c;
                total += c;
            }
#pragma unroll 1
            for (int it = 0; it < nit; it++) {
                const int k = kb + it * kstride + lane;
                float e = 0.f;
                if (k < kend) e = sample_weight(costs[k], beta_c, nil);
                const unsigned m = __ballot_sync(0xffffffffu, e != 0.f);
                MPPI_CHECK(e == 0.f || (off + __popc(m & lt) >= 0 && off + __popc(m & lt) < NW * kListCap));
                if (e != 0.f) sm.sList[off + __popc(m & lt)] = make_uint2((uint32_t)(p.k_offset + k), __float_as_uint(e));
                off += __popc(m);
            }
            __syncthreads();
            trace_stamp(p, ctrl, 8);
            // work items (chunk, entry group), chunk-major; warp w takes the contiguous range [n w / NW, n (w+1) / NW)
            const int n_groups = (total + 31) >> 5;
            const int n_items = n_groups * nchunk;
            const int it_lo = (int)((long long)n_items * warp / NW), it_hi = (int)((long long)n_items * (warp + 1) / NW);
            int cur_ch = -1;
            float2 acc2[16];
            auto flush = [&](int ch) {
                float acc[32];
#pragma unroll
                for (int i = 0; i < 16; i++) { acc[2 * i] = acc2[i].x; acc[2 * i + 1] = acc2[i].y; }
                const float r = warp_transpose_sum32(acc, lane);
                MPPI_CHECK(ch >= 0 && ch * 32 + lane < TAp);
                myacc[ch * 32 + lane] += r;
            };
#pragma unroll 1
            for (int it = it_lo; it < it_hi; it++) {
                const int ch = it / n_groups, i0 = 32 * (it - ch * n_groups);
                if (ch != cur_ch) {
                    if (cur_ch >= 0) flush(cur_ch);
#pragma unroll
                    for (int i = 0; i < 16; i++) acc2[i] = make_float2(0.f, 0.f);
                    cur_ch = ch;
                }
                MPPI_CHECK(i0 >= 0 && i0 < total && total <= NW * kListCap && ch < nchunk);
                const uint2 ent = (i0 + lane < total) ? sm.sList[i0 + lane] : make_uint2(0u, 0u);   // padding lanes: weight 0
                const PhiloxSample ps = philox_sample(phA, ent.x);
                const float e = __uint_as_float(ent.y);
                const float2 e2 = make_float2(e, e);
                const bool full = (ch * 8 + 8 <= ncall);
#pragma unroll
                for (int c8 = 0; c8 < 8; c8++) {
                    if (full || ch * 8 + c8 < ncall) {
                        const float4 z = Gen::normals4(sTab, (uint32_t)(ch * 8 + c8), ps, p);
                        acc2[2 * c8] = __ffma2_rn(e2, make_float2(z.x, z.y), acc2[2 * c8]);
                        acc2[2 * c8 + 1] = __ffma2_rn(e2, make_float2(z.z, z.w), acc2[2 * c8 + 1]);
                    }
                }
            }
            if (cur_ch >= 0) flush(cur_ch);
            __syncthreads();                         // list and counts are reused by the next batch / the reductions below
            trace_stamp(p, ctrl, 9);
        }
    }
    eta = warp_sum(eta);
    if (lane == 0) sRed[warp] = eta;
    __syncthreads();
    float eta_c = 0.f;
#pragma unroll
    for (int w = 0; w < NW; w++) eta_c += sRed[w];
    for (int j = tid; j < TA; j += kPhiloxThreads) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; w++) s += sAcc[w * TAp + j];
        sm.sN[j] = s;
    }
    __syncthreads();
    publish_and_finish<A, true>(p, ctrl, beta_c, eta_c, sm.sN, sm.sWork, sm.sScale, sRed, sm.sScratch, kPhiloxThreads);
}

}  // namespace mppi
