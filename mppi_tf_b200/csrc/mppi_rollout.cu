// Fused MPPI update kernels for the point-mass model (sm_100a).
//
//   rollout_philox_kernel    the hot path: noise regenerated in registers from a counter-based
//                            Philox stream, never stored.  Phase 1 rolls every sample of the CTA
//                            through the horizon and keeps only its cost; phase 2 re-walks the same
//                            Philox counters to accumulate sum_k e_k z_k.  One launch per update:
//                            the last CTA to finish merges all CTA partials, applies the update
//                            and shifts the sequence.
//   rollout_injected_kernel  parity/debug mode: eps is read from HBM exactly once.  Each warp owns
//                            a 32-sample tile that the TMA unit (cp.async.bulk) lands in shared
//                            memory; the tile stays resident between the rollout (one lane per
//                            sample) and the weighted sum (one lane per column) with an online
//                            max-shifted rescale.
//   finish_kernel            multi-rank only: merges the all-gathered rank payloads.
//
// Reference maths: /root/reference/src/controller_base.cpp:166-329, src/model_base.cpp:53-82,
// src/cost_base.cpp:37-68 (restated in oracle/mppi_oracle_impl.h, which tests compare against).
#include "mppi_device.cuh"
#include "mppi_internal.h"

namespace mppi {

constexpr int kPhiloxThreads = 256;
constexpr int kMaxParts = 2048;   // CTA partials (or ranks) one merge can take

template <int A>
struct Row {
    static constexpr int RS = (2 * A + 3) & ~3;   // floats per staged step: U_t[A], w_t[A], pad
};

// Stage the per-step uniforms of controller `ctrl` into shared memory:
//   sUV[t][0..A)  = U_t          (mean action, src/controller_base.cpp:205-208)
//   sUV[t][A..2A) = w_t          action-cost vector: Philox mode lambda*U_t (since eps = Sigma z,
//                                 lambda U^T Sigma^-1 eps = lambda U^T z); injected mode
//                                 lambda*Sigma^-T U_t (src/cost_base.cpp:63-68)
template <int A, bool PHILOX>
__device__ __forceinline__ void stage_sequence(const RolloutParams &p, int ctrl, float *sUV)
{
    constexpr int RS = Row<A>::RS;
    const float *U = p.U + (size_t)ctrl * p.TA;
    for (int i = threadIdx.x; i < p.T * RS; i += blockDim.x) {
        const int t = i / RS, j = i - t * RS;
        float v = 0.f;
        if (j < A) {
            v = U[t * A + j];
        } else if (j < 2 * A) {
            const int r = j - A;
            if (PHILOX) {
                v = p.lambda * U[t * A + r];
            } else {
#pragma unroll
                for (int l = 0; l < A; l++) v = fmaf(p.lam_inv_sigma_T[r * A + l], U[t * A + l], v);
            }
        }
        sUV[i] = v;
    }
}

// One rollout step for one sample (src/controller_base.cpp:251-269):
//   u = U_t + eps_t ; x <- A x + (B/m) u ; S += q(x) + lambda U_t^T Sigma^-1 eps_t
// `n` is z_t in Philox mode (eps_t = Sigma z_t formed here) and eps_t in injected mode.
template <int A, bool PHILOX, bool DIAG>
__device__ __forceinline__ void rollout_step(PointMass<A> &x, float &S, const float *uv_row,
                                             const float *n, const RolloutParams &p,
                                             const float (&g)[2 * A], const float (&q)[2 * A])
{
    constexpr int RS = Row<A>::RS;
    float uv[RS];
#pragma unroll
    for (int i = 0; i < RS / 4; i++) {
        const float4 v = reinterpret_cast<const float4 *>(uv_row)[i];
        uv[4 * i] = v.x; uv[4 * i + 1] = v.y; uv[4 * i + 2] = v.z; uv[4 * i + 3] = v.w;
    }
    float u[A];
    float ac = 0.f;
#pragma unroll
    for (int j = 0; j < A; j++) {
        float e;
        if (PHILOX) {
            if (DIAG) {
                e = p.sigma[j * A + j] * n[j];
            } else {
                e = 0.f;
#pragma unroll
                for (int l = 0; l < A; l++) e = fmaf(p.sigma[j * A + l], n[l], e);
            }
        } else {
            e = n[j];
        }
        u[j] = uv[j] + e;
        ac = fmaf(uv[A + j], n[j], ac);
    }
    x.step(u, p.dt, p.c_pu, p.c_vu);
    S += x.state_cost(g, q) + ac;
}

// -------------------------------------------------------------------------------------------------
// Merge of partial records {beta, eta, -, -, N[TA]} (CTA partials or rank payloads), block-wide and
// in a fixed order (deterministic).  On return sN[0..TA) holds sum_c scale_c N_c and the returned
// beta/eta are the merged values: beta = min_c beta_c, scale_c = exp(-(beta_c - beta)/lambda).
// -------------------------------------------------------------------------------------------------
struct Merged { float beta, eta; };

__device__ Merged merge_parts(const float *parts, size_t part_stride, int nparts, int TA,
                              float neg_inv_lambda_log2e, float *sN, float *sScale, float *sRed)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    float b = kInf;
    for (int c = tid; c < nparts; c += blockDim.x) b = fminf(b, __ldcg(parts + c * part_stride));
    b = warp_min(b);
    if (lane == 0) sRed[warp] = b;
    __syncthreads();
    float beta = sRed[0];
    for (int w = 1; w < nw; w++) beta = fminf(beta, sRed[w]);
    __syncthreads();
    float e = 0.f;
    for (int c = tid; c < nparts; c += blockDim.x) {
        const float bc = __ldcg(parts + c * part_stride);
        const float sc = (bc == kInf) ? 0.f : weight_exp(bc, beta, neg_inv_lambda_log2e);
        sScale[c] = sc;
        e = fmaf(sc, __ldcg(parts + c * part_stride + 1), e);
    }
    e = warp_sum(e);
    if (lane == 0) sRed[warp] = e;
    __syncthreads();
    float eta = 0.f;
    for (int w = 0; w < nw; w++) eta += sRed[w];
    for (int j = tid; j < TA; j += blockDim.x) {
        float acc = 0.f;
        for (int c = 0; c < nparts; c++) acc = fmaf(sScale[c], __ldcg(parts + c * part_stride + 4 + j), acc);
        sN[j] = acc;
    }
    __syncthreads();
    return Merged{beta, eta};
}

// U' = U + Delta (src/controller_base.cpp:223), next = U'[0] (:327-329), U <- [U'[1:], 0] (:310-324).
// In Philox mode sN holds sum e z, so Delta_t = Sigma (sN_t) / eta; injected mode sums eps directly.
template <int A, bool PHILOX>
__device__ void apply_update(const RolloutParams &p, int ctrl, Merged m, float *sN, float *sOut)
{
    const int TA = p.TA;
    const float inv_eta = 1.0f / m.eta;
    float *U = p.U + (size_t)ctrl * TA;
    for (int i = threadIdx.x; i < TA; i += blockDim.x) {
        const int t = i / A, j = i - t * A;
        float d;
        if (PHILOX) {
            d = 0.f;
#pragma unroll
            for (int l = 0; l < A; l++) d = fmaf(p.sigma[j * A + l], sN[t * A + l], d);
        } else {
            d = sN[i];
        }
        const float un = U[i] + d * inv_eta;
        sOut[i] = un;
        p.U_new[(size_t)ctrl * TA + i] = un;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TA; i += blockDim.x) {
        U[i] = (i + A < TA) ? sOut[i + A] : 0.f;
        if (i < A) p.next[ctrl * A + i] = sOut[i];
    }
    if (threadIdx.x == 0) {
        p.stats[2 * ctrl] = m.beta;
        p.stats[2 * ctrl + 1] = m.eta;
    }
}

// Publish this CTA's partial; the last CTA of the controller merges and finishes the update.
// sN: CTA sums [TA]; sWork: >= TA floats scratch; sScale: kMaxParts floats; sRed: 32 floats.
template <int A, bool PHILOX>
__device__ void publish_and_finish(const RolloutParams &p, int ctrl, float beta_c, float eta_c, float *sN,
                                   float *sWork, float *sScale, float *sRed)
{
    __shared__ int s_is_last;
    const int TA = p.TA, stride = partial_stride(TA), nparts = gridDim.x;
    float *mine = p.partials + ((size_t)ctrl * nparts + blockIdx.x) * stride;
    if (threadIdx.x == 0) { mine[0] = beta_c; mine[1] = eta_c; }
    for (int j = threadIdx.x; j < TA; j += blockDim.x) mine[4 + j] = sN[j];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&p.counters[ctrl], 1u);
        s_is_last = (prev == (unsigned)nparts - 1u);
    }
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();
    if (threadIdx.x == 0) p.counters[ctrl] = 0u;
    Merged m = merge_parts(p.partials + (size_t)ctrl * nparts * stride, stride, nparts, TA,
                           p.neg_inv_lambda_log2e, sN, sScale, sRed);
    if (p.world > 1) {
        float *pay = p.payload + (size_t)ctrl * stride;
        if (threadIdx.x == 0) { pay[0] = m.beta; pay[1] = m.eta; pay[2] = 0.f; pay[3] = 0.f; }
        for (int j = threadIdx.x; j < stride - 4; j += blockDim.x) pay[4 + j] = (j < TA) ? sN[j] : 0.f;
        return;
    }
    apply_update<A, PHILOX>(p, ctrl, m, sN, sWork);
}

// -------------------------------------------------------------------------------------------------
// Philox mode
// -------------------------------------------------------------------------------------------------
template <int A, bool DIAG>
__global__ void __launch_bounds__(kPhiloxThreads, 4)
rollout_philox_kernel(const __grid_constant__ RolloutParams p)
{
    constexpr int RS = Row<A>::RS;
    constexpr int NW = kPhiloxThreads / 32;
    extern __shared__ float4 smem_f4[];
    float *smem = reinterpret_cast<float *>(smem_f4);
    const int TA = p.TA, TAp = (TA + 31) & ~31;
    float *sUV = smem;                   // [T][RS]
    float *sAcc = sUV + p.T * RS;        // [NW][TAp]  per-warp chunk sums
    float *sN = sAcc + NW * TAp;         // [TAp]
    float *sWork = sN + TAp;             // [TAp]
    float *sScale = sWork + TAp;         // [kMaxParts]
    float *sRed = sScale + kMaxParts;    // [32]

    const int ctrl = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    stage_sequence<A, true>(p, ctrl, sUV);

    float g[2 * A], q[2 * A], x0[2 * A];
    {
        const float *gp = p.goal + (p.goal_per_ctrl ? (size_t)ctrl * 2 * A : 0);
        const float *xp = p.x + (size_t)ctrl * 2 * A;
#pragma unroll
        for (int i = 0; i < 2 * A; i++) {
            q[i] = p.sqrt_q[i];
            g[i] = q[i] * gp[i];
            x0[i] = p.x_inline ? p.x0[i] : xp[i];
        }
    }
    __syncthreads();

    float *costs = p.costs + (size_t)ctrl * p.K_local;
    const int kstride = gridDim.x * kPhiloxThreads;
    const int kfirst = blockIdx.x * kPhiloxThreads + tid;
    const int nfull = p.T >> 2, trem = p.T & 3;
    const uint32_t stream = (uint32_t)ctrl;

    // ---- phase 1: rollout + cost ---------------------------------------------------------------
    float bmin = kInf;
    for (int it = 0, k = kfirst; it < p.n_iter; it++, k += kstride) {
        if (k >= p.K_local) break;
        const uint32_t kg = (uint32_t)(p.k_offset + k);
        PointMass<A> x;
        x.init(x0);
        float S = 0.f;
        const float *uv = sUV;
        uint32_t call = 0;
        for (int tb = 0; tb < nfull; tb++) {       // full blocks of 4 steps = A Philox calls, no guards
            float z[4 * A];
#pragma unroll
            for (int c = 0; c < A; c++) normals4(call + c, kg, stream, p, &z[4 * c]);
            call += A;
#pragma unroll
            for (int tt = 0; tt < 4; tt++) rollout_step<A, true, DIAG>(x, S, uv + tt * RS, &z[tt * A], p, g, q);
            uv += 4 * RS;
        }
        if (trem) {                                 // tail: T % 4 steps
            float z[4 * A];
#pragma unroll
            for (int c = 0; c < A; c++) normals4(call + c, kg, stream, p, &z[4 * c]);
#pragma unroll
            for (int tt = 0; tt < 3; tt++)
                if (tt < trem) rollout_step<A, true, DIAG>(x, S, uv + tt * RS, &z[tt * A], p, g, q);
        }
        S += x.state_cost(g, q);    // terminal cost on top of step T-1's (src/controller_base.cpp:271-272)
        costs[k] = S;
        bmin = fminf(bmin, S);
    }
    bmin = warp_min(bmin);
    if (lane == 0) sRed[warp] = bmin;
    __syncthreads();
    float beta_c = sRed[0];
#pragma unroll
    for (int w = 1; w < NW; w++) beta_c = fminf(beta_c, sRed[w]);
    __syncthreads();

    // ---- phase 2: sum_k e_k z_k, regenerating z from the same counters ----------------------------
    const int ncall = (TA + 3) >> 2;
    const int nchunk = (ncall + 7) >> 3;
    float eta = 0.f;
    for (int ch = 0; ch < nchunk; ch++) {
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; i++) acc[i] = 0.f;
        const bool full = (ch * 8 + 8 <= ncall);    // warp-uniform: all 8 calls of the chunk exist
        for (int it = 0, k = kfirst; it < p.n_iter; it++, k += kstride) {
            if (k >= p.K_local) break;
            const uint32_t kg = (uint32_t)(p.k_offset + k);
            const float e = weight_exp(costs[k], beta_c, p.neg_inv_lambda_log2e);
            if (ch == 0) eta += e;
            if (full) {
#pragma unroll
                for (int c8 = 0; c8 < 8; c8++) {
                    float z[4];
                    normals4((uint32_t)(ch * 8 + c8), kg, stream, p, z);
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[4 * c8 + j] = fmaf(e, z[j], acc[4 * c8 + j]);
                }
            } else {
#pragma unroll
                for (int c8 = 0; c8 < 8; c8++) {
                    if (ch * 8 + c8 < ncall) {
                        float z[4];
                        normals4((uint32_t)(ch * 8 + c8), kg, stream, p, z);
#pragma unroll
                        for (int j = 0; j < 4; j++) acc[4 * c8 + j] = fmaf(e, z[j], acc[4 * c8 + j]);
                    }
                }
            }
        }
        const float r = warp_transpose_sum32(acc, lane);
        sAcc[warp * TAp + ch * 32 + lane] = r;
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
        sN[j] = s;
    }
    __syncthreads();
    publish_and_finish<A, true>(p, ctrl, beta_c, eta_c, sN, sWork, sScale, sRed);
}

// -------------------------------------------------------------------------------------------------
// Injected-noise mode
// -------------------------------------------------------------------------------------------------
struct InjectedLaunch {
    int nw;      // consumer warps per CTA
    int stages;  // tile buffers per warp (each warp owns a private ring: no cross-warp barrier aliasing)
};

template <int A, bool TMA>
__global__ void __launch_bounds__(512, 1)
rollout_injected_kernel(const __grid_constant__ RolloutParams p, const InjectedLaunch L)
{
    constexpr int RS = Row<A>::RS;
    extern __shared__ float4 smem_f4[];
    float *smem = reinterpret_cast<float *>(smem_f4);
    const int TA = p.TA, TAp = (TA + 31) & ~31;
    const int NW = L.nw, ST = L.stages, NBUF = NW * ST;
    const int tile_words = 32 * TA;                // multiple of 32 words: every tile 128-B aligned
    float *sTiles = smem;                          // [NBUF][32][TA]
    float *sUV = sTiles + (size_t)NBUF * tile_words;   // [T][RS]
    float *sAccW = sUV + p.T * RS;                 // [NW][TAp] running per-warp sums
    float *sN = sAccW + NW * TAp;                  // [TAp]
    float *sWork = sN + TAp;                       // [TAp]
    float *sScale = sWork + TAp;                   // [kMaxParts]
    float *sRed = sScale + kMaxParts;              // [64]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sRed + 64);   // [NBUF]

    const int ctrl = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_tiles = (p.K_local + 31) >> 5;
    const int nseq = (n_tiles > (int)blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const float *eps = p.eps + (size_t)ctrl * p.K_local * TA;

    // Warp w consumes tile sequence numbers seq = w + j*NW (j = 0,1,..) of this CTA; its j-th tile
    // lands in its private buffer w*ST + j%ST and completes phase j/ST of that buffer's mbarrier.
    auto issue = [&](int w, int j) {   // one elected lane; TMA bulk copy of a whole tile (contiguous in HBM)
        const int b = w * ST + (j % ST);
        const int seq = w + j * NW;
        const int gt = seq * gridDim.x + blockIdx.x;
        const int rows = min(32, p.K_local - 32 * gt);
        const uint32_t bytes = (uint32_t)rows * TA * 4u;
        mbar_expect_tx(&bars[b], bytes);
        bulk_g2s(sTiles + (size_t)b * tile_words, eps + (size_t)gt * tile_words, bytes, &bars[b]);
    };

    if (TMA) {
        if (tid == 0) {
            for (int b = 0; b < NBUF; b++) mbar_init(&bars[b], 1);
            fence_mbar_init();
        }
    }
    stage_sequence<A, false>(p, ctrl, sUV);
    for (int i = tid; i < NW * TAp; i += blockDim.x) sAccW[i] = 0.f;
    __syncthreads();
    if (TMA && lane == 0 && warp < NW) {
        for (int j = 0; j < ST && warp + j * NW < nseq; j++) issue(warp, j);
    }

    float g[2 * A], q[2 * A], x0[2 * A];
    {
        const float *gp = p.goal + (p.goal_per_ctrl ? (size_t)ctrl * 2 * A : 0);
        const float *xp = p.x + (size_t)ctrl * 2 * A;
#pragma unroll
        for (int i = 0; i < 2 * A; i++) {
            q[i] = p.sqrt_q[i];
            g[i] = q[i] * gp[i];
            x0[i] = p.x_inline ? p.x0[i] : xp[i];
        }
    }

    float *costs = p.costs + (size_t)ctrl * p.K_local;
    float *accw = sAccW + warp * TAp;
    const int nblk = (p.T + 3) >> 2;
    float beta_w = kInf, eta_lane = 0.f;

    if (warp < NW) {
        for (int j = 0, seq = warp; seq < nseq; j++, seq += NW) {
            const int b = warp * ST + (j % ST);
            const int gt = seq * gridDim.x + blockIdx.x;
            const int rows = min(32, p.K_local - 32 * gt);
            float *tile = sTiles + (size_t)b * tile_words;
            if (TMA) {
                mbar_wait(&bars[b], (uint32_t)((j / ST) & 1));
                if (rows < 32) {   // zero the rows the copy did not write (e = 0 must not meet NaN garbage)
                    for (int i = rows * TA + lane; i < tile_words; i += 32) tile[i] = 0.f;
                }
            } else {           // generic fallback (T*a not a multiple of 4): coalesced loads by the warp
                const float *src = eps + (size_t)gt * tile_words;
                const int nvalid = rows * TA;
                for (int i = lane; i < tile_words; i += 32) tile[i] = (i < nvalid) ? __ldg(src + i) : 0.f;
            }
            __syncwarp();

            // ---- phase 1: lane = sample, rollout over the row held in shared memory --------------
            const float *row = tile + lane * TA;
            PointMass<A> x;
            x.init(x0);
            float S = 0.f;
            for (int tb = 0; tb < nblk; tb++) {
                float e[4 * A];
                if (TMA) {   // T*a % 4 == 0: rows are 16-B aligned, conflict-free LDS.128 when T*a/4 is odd
#pragma unroll
                    for (int c = 0; c < A; c++) {
                        const int i4 = tb * A + c;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (4 * i4 < TA) v = reinterpret_cast<const float4 *>(row)[i4];
                        e[4 * c] = v.x; e[4 * c + 1] = v.y; e[4 * c + 2] = v.z; e[4 * c + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4 * A; i++) {
                        const int idx = tb * 4 * A + i;
                        e[i] = (idx < TA) ? row[idx] : 0.f;
                    }
                }
#pragma unroll
                for (int tt = 0; tt < 4; tt++) {
                    const int t = 4 * tb + tt;
                    if (t < p.T) rollout_step<A, false, false>(x, S, sUV + t * RS, &e[tt * A], p, g, q);
                }
            }
            S += x.state_cost(g, q);
            if (lane < rows) costs[32 * gt + lane] = S; else S = kInf;

            // ---- phase 2: online max-shifted weights; lane = column of the resident tile ---------
            const float m = warp_min(S);
            if (m < beta_w) {
                if (beta_w != kInf) {
                    const float f = weight_exp(beta_w, m, p.neg_inv_lambda_log2e);
                    for (int c = lane; c < TA; c += 32) accw[c] *= f;
                    eta_lane *= f;
                }
                beta_w = m;
            }
            const float ek = (lane < rows) ? weight_exp(S, beta_w, p.neg_inv_lambda_log2e) : 0.f;
            eta_lane += ek;
            for (int cg = 0; cg < TA; cg += 128) {
                int col[4];
                float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int mm = 0; mm < 4; mm++) col[mm] = min(cg + 32 * mm + lane, TA - 1);
#pragma unroll 8
                for (int k = 0; k < 32; k++) {
                    const float w = __shfl_sync(0xffffffffu, ek, k);
                    const float *trow = tile + k * TA;
#pragma unroll
                    for (int mm = 0; mm < 4; mm++) a4[mm] = fmaf(w, trow[col[mm]], a4[mm]);
                }
#pragma unroll
                for (int mm = 0; mm < 4; mm++) {
                    const int c = cg + 32 * mm + lane;
                    if (c < TA) accw[c] += a4[mm];
                }
            }
            __syncwarp();
            if (TMA && lane == 0 && seq + ST * NW < nseq) {
                fence_proxy_async();
                issue(warp, j + ST);
            }
        }
        eta_lane = warp_sum(eta_lane);
        if (lane == 0) { sRed[warp] = beta_w; sRed[32 + warp] = eta_lane; }
    }
    __syncthreads();

    // ---- CTA merge of the per-warp running sums -------------------------------------------------
    float beta_c = kInf;
    for (int w = 0; w < NW; w++) beta_c = fminf(beta_c, sRed[w]);
    float eta_c = 0.f;
    for (int w = 0; w < NW; w++) {
        const float bw = sRed[w];
        if (bw != kInf) eta_c = fmaf(weight_exp(bw, beta_c, p.neg_inv_lambda_log2e), sRed[32 + w], eta_c);
    }
    for (int j = tid; j < TA; j += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < NW; w++) {
            const float bw = sRed[w];
            if (bw != kInf) s = fmaf(weight_exp(bw, beta_c, p.neg_inv_lambda_log2e), sAccW[w * TAp + j], s);
        }
        sN[j] = s;
    }
    __syncthreads();
    publish_and_finish<A, false>(p, ctrl, beta_c, eta_c, sN, sWork, sScale, sRed);
}

// -------------------------------------------------------------------------------------------------
// Multi-rank finish: merge the all-gathered payloads [world][n_ctrl][stride] and apply.
// -------------------------------------------------------------------------------------------------
template <int A, bool PHILOX>
__global__ void __launch_bounds__(256) finish_kernel(const __grid_constant__ RolloutParams p,
                                                     const float *gathered)
{
    extern __shared__ float4 smem_f4[];
    float *smem = reinterpret_cast<float *>(smem_f4);
    const int TA = p.TA, TAp = (TA + 31) & ~31;
    float *sN = smem, *sWork = sN + TAp, *sScale = sWork + TAp, *sRed = sScale + kMaxParts;
    const int ctrl = blockIdx.x, stride = partial_stride(TA);
    Merged m = merge_parts(gathered + (size_t)ctrl * stride, (size_t)p.n_ctrl * stride, p.world, TA,
                           p.neg_inv_lambda_log2e, sN, sScale, sRed);
    apply_update<A, PHILOX>(p, ctrl, m, sN, sWork);
}

// Regenerate eps = Sigma z of one update for this rank's samples (mppi_dump_noise).
template <int A>
__global__ void dump_noise_kernel(const __grid_constant__ RolloutParams p, float *out)
{
    const int ctrl = blockIdx.y;
    const int ncall = (p.TA + 3) >> 2;
    const long long total = (long long)p.K_local * ncall;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / ncall), call = (int)(i - (long long)k * ncall);
        float z[4];
        normals4((uint32_t)call, (uint32_t)(p.k_offset + k), (uint32_t)ctrl, p, z);
        float *row = out + ((size_t)ctrl * p.K_local + k) * p.TA;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (4 * call + j < p.TA) row[4 * call + j] = z[j];   // z first; scaled below
    }
}
template <int A>
__global__ void scale_noise_kernel(const __grid_constant__ RolloutParams p, float *io)
{
    const long long total = (long long)p.n_ctrl * p.K_local * p.T;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        float z[A], e[A];
#pragma unroll
        for (int j = 0; j < A; j++) z[j] = io[i * A + j];
#pragma unroll
        for (int j = 0; j < A; j++) {
            e[j] = 0.f;
#pragma unroll
            for (int l = 0; l < A; l++) e[j] = fmaf(p.sigma[j * A + l], z[l], e[j]);
        }
#pragma unroll
        for (int j = 0; j < A; j++) io[i * A + j] = e[j];
    }
}

// -------------------------------------------------------------------------------------------------
// Host launchers
// -------------------------------------------------------------------------------------------------
static size_t philox_smem_bytes(int A, int T, int TA)
{
    const int RS = (2 * A + 3) & ~3, TAp = (TA + 31) & ~31, NW = kPhiloxThreads / 32;
    return sizeof(float) * ((size_t)T * RS + (size_t)NW * TAp + 2 * TAp + kMaxParts + 32);
}

template <int A>
static cudaError_t launch_philox_A(const RolloutParams &p, dim3 grid, cudaStream_t st)
{
    const size_t smem = philox_smem_bytes(A, p.T, p.TA);
    cudaError_t err;
    if (p.sigma_diag) {
        err = cudaFuncSetAttribute(rollout_philox_kernel<A, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        rollout_philox_kernel<A, true><<<grid, kPhiloxThreads, smem, st>>>(p);
    } else {
        err = cudaFuncSetAttribute(rollout_philox_kernel<A, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        rollout_philox_kernel<A, false><<<grid, kPhiloxThreads, smem, st>>>(p);
    }
    return cudaGetLastError();
}

#define MPPI_DISPATCH_A(a, ...)                  \
    switch (a) {                                 \
        case 1: { constexpr int A_ = 1; __VA_ARGS__; } break; \
        case 2: { constexpr int A_ = 2; __VA_ARGS__; } break; \
        case 3: { constexpr int A_ = 3; __VA_ARGS__; } break; \
        case 4: { constexpr int A_ = 4; __VA_ARGS__; } break; \
        case 5: { constexpr int A_ = 5; __VA_ARGS__; } break; \
        case 6: { constexpr int A_ = 6; __VA_ARGS__; } break; \
        case 7: { constexpr int A_ = 7; __VA_ARGS__; } break; \
        case 8: { constexpr int A_ = 8; __VA_ARGS__; } break; \
        default: return cudaErrorInvalidValue;   \
    }

int philox_grid_x(int K_local, int n_ctrl, int num_sms, int *n_iter)
{
    const int ctas_total = num_sms * 4;                      // 4 resident CTAs of 256 threads per SM
    int per_ctrl = ctas_total / (n_ctrl > 0 ? n_ctrl : 1);
    if (per_ctrl < 1) per_ctrl = 1;
    const int need = (K_local + kPhiloxThreads - 1) / kPhiloxThreads;
    int gx = need < per_ctrl ? need : per_ctrl;
    if (gx > kMaxParts) gx = kMaxParts;
    int it = (K_local + gx * kPhiloxThreads - 1) / (gx * kPhiloxThreads);
    gx = (K_local + it * kPhiloxThreads - 1) / (it * kPhiloxThreads);   // rebalance
    *n_iter = it;
    return gx;
}

cudaError_t launch_rollout_philox(RolloutParams p, int a, int num_sms, cudaStream_t st, int *grid_x_out)
{
    int n_iter = 1;
    const int gx = philox_grid_x(p.K_local, p.n_ctrl, num_sms, &n_iter);
    p.n_iter = n_iter;
    if (grid_x_out) *grid_x_out = gx;
    dim3 grid(gx, p.n_ctrl);
    MPPI_DISPATCH_A(a, return launch_philox_A<A_>(p, grid, st));
    return cudaSuccess;
}

// Injected mode geometry: how many consumer warps / tile buffers fit in shared memory.
bool injected_geometry(int A, int T, int TA, int K_local, int n_ctrl, int num_sms, size_t smem_limit,
                       int *nw_out, int *stages_out, int *grid_x_out, size_t *smem_out)
{
    const int RS = (2 * A + 3) & ~3, TAp = (TA + 31) & ~31;
    const size_t tile_b = (size_t)128 * TA;
    const int n_tiles = (K_local + 31) / 32;
    // several CTAs per SM when there are many small controllers
    int ctas_per_sm = 1;
    if (n_ctrl >= 2 * num_sms) ctas_per_sm = 4;
    const size_t budget = smem_limit / ctas_per_sm - (ctas_per_sm > 1 ? 1024 : 0);
    const size_t fixed = sizeof(float) * ((size_t)T * RS + 2 * TAp + kMaxParts + 64) + 8 * 64 + 64;
    const size_t per_warp = sizeof(float) * (size_t)TAp;
    if (fixed + per_warp + tile_b > budget) return false;
    int nw = (int)((budget - fixed) / (per_warp + tile_b));
    if (nw > 16) nw = 16;
    if (nw > n_tiles) nw = n_tiles > 0 ? n_tiles : 1;
    int st = (int)((budget - fixed - nw * per_warp) / (nw * tile_b));
    if (st > 3) st = 3;
    if (st < 1) st = 1;
    int gx;
    const int need = (n_tiles + nw - 1) / nw;
    if (n_ctrl == 1) {
        gx = num_sms * ctas_per_sm;
    } else {
        gx = (num_sms * ctas_per_sm) / n_ctrl;
    }
    if (gx > need) gx = need;
    if (gx < 1) gx = 1;
    if (gx > kMaxParts) gx = kMaxParts;
    // do not keep more stages than this CTA has tiles per warp
    const int tiles_per_warp = (n_tiles + gx * nw - 1) / (gx * nw);
    if (st > tiles_per_warp) st = tiles_per_warp > 0 ? tiles_per_warp : 1;
    *nw_out = nw;
    *stages_out = st;
    *grid_x_out = gx;
    *smem_out = fixed + nw * per_warp + (size_t)nw * st * tile_b;
    return true;
}

template <int A>
static cudaError_t launch_injected_A(const RolloutParams &p, InjectedLaunch L, dim3 grid, size_t smem, bool tma,
                                     cudaStream_t st)
{
    cudaError_t err;
    const int threads = L.nw * 32;
    if (tma) {
        err = cudaFuncSetAttribute(rollout_injected_kernel<A, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        rollout_injected_kernel<A, true><<<grid, threads, smem, st>>>(p, L);
    } else {
        err = cudaFuncSetAttribute(rollout_injected_kernel<A, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        rollout_injected_kernel<A, false><<<grid, threads, smem, st>>>(p, L);
    }
    return cudaGetLastError();
}

cudaError_t launch_rollout_injected(RolloutParams p, int a, int num_sms, size_t smem_limit, cudaStream_t st,
                                    int *grid_x_out)
{
    InjectedLaunch L;
    int gx = 1;
    size_t smem = 0;
    if (!injected_geometry(a, p.T, p.TA, p.K_local, p.n_ctrl, num_sms, smem_limit, &L.nw, &L.stages, &gx, &smem))
        return cudaErrorInvalidConfiguration;
    if (grid_x_out) *grid_x_out = gx;
    // TMA path needs 16-byte aligned tiles and rows: T*a % 4 == 0 and an aligned base pointer.
    const bool tma = (p.TA % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.eps) & 15u) == 0);
    dim3 grid(gx, p.n_ctrl);
    MPPI_DISPATCH_A(a, return launch_injected_A<A_>(p, L, grid, smem, tma, st));
    return cudaSuccess;
}

int max_grid_x(int K_local, int n_ctrl, int num_sms)
{
    int n_iter;
    int g1 = philox_grid_x(K_local, n_ctrl, num_sms, &n_iter);
    int g2 = num_sms * 4;
    return g1 > g2 ? g1 : g2;
}

cudaError_t launch_finish(RolloutParams p, int a, bool philox, const float *gathered, cudaStream_t st)
{
    const int TAp = (p.TA + 31) & ~31;
    const size_t smem = sizeof(float) * (2 * (size_t)TAp + kMaxParts + 32);
    MPPI_DISPATCH_A(a, {
        if (philox) finish_kernel<A_, true><<<p.n_ctrl, 256, smem, st>>>(p, gathered);
        else finish_kernel<A_, false><<<p.n_ctrl, 256, smem, st>>>(p, gathered);
    });
    return cudaGetLastError();
}

cudaError_t launch_dump_noise(RolloutParams p, int a, float *out_dev, cudaStream_t st)
{
    dim3 grid(296, p.n_ctrl);
    MPPI_DISPATCH_A(a, {
        dump_noise_kernel<A_><<<grid, 256, 0, st>>>(p, out_dev);
        scale_noise_kernel<A_><<<296, 256, 0, st>>>(p, out_dev);
    });
    return cudaGetLastError();
}

}  // namespace mppi
