// Fused MPPI update with the AUV (Fossen) dynamics model (sm_100a) — SURVEY.md section 8f, row N4.
//
//   rollout_auv_kernel<PHILOX>   one launch per update, same structure as the point-mass kernels: phase 1 rolls
//                                every sample of the CTA through the horizon (explicit Euler / Heun / the
//                                reference's rk-4 variant of the 13-state rigid-body model, quaternion
//                                renormalised every step) and keeps only its cost; phase 2 forms the
//                                max-shifted weights and the weighted noise sums (Philox mode: the noise is
//                                regenerated from the same counters; injected mode: eps is re-read, lane =
//                                column); the last CTA merges the CTA partials, applies the update and shifts.
//   auv_predict_kernel           AUVModel.build_step_graph on a batch (the model's `predict`).
//   quat_cost_kernel             StaticQuatCost.state_cost on a batch.
//
// A sample-step is ~0.5 (rk1) to ~2 (rk4) kFLOP of dependent fp32 arithmetic against 24 bytes of noise: the
// kernel is bound by the FMA pipe, not by HBM; model parameters (160 words) ride in the kernel parameter block
// and are read as constant-bank operands.
//
// Reference maths: /root/reference/scripts/src/models/auv_model.py:285-559,
// scripts/src/controllers/controller_base.py:371-474, scripts/src/costs/static_cost.py:40-63,116-159,
// scripts/src/costs/cost_base.py:114-170 (restated in the CPU checker that tests compare against).
#include <cstdlib>

#include "mppi_auv.cuh"
#include "mppi_internal.h"
#include "mppi_update.cuh"

namespace mppi {

// CTA shape: THREADS x CTAS-per-SM is a template parameter pair so that the register budget (65536 / resident
// threads) can be chosen per integrator; see auv_geometry().

// n = z (Philox mode: eps = sigma z formed here) or eps (injected mode)
template <bool PHILOX, int RK>
__device__ __forceinline__ void auv_rollout_step(const RolloutParams &p, const AuvParams &P, const float *sNN, const float *uv_row,
                                                 const float (&n)[kAuvA], const float *g, float (&x)[kAuvS], float &S)
{
    constexpr int A = kAuvA, H = Row<A>::H;
    float u[A];
#pragma unroll
    for (int j = 0; j < A; j++) {
        float e;
        if (PHILOX) {
            e = 0.f;
#pragma unroll
            for (int l = 0; l < A; l++) e = fmaf(p.sigma[j * A + l], n[l], e);
        } else {
            e = n[j];
        }
        u[j] = uv_row[j] + e;
    }
    float ac = 0.f;
#pragma unroll
    for (int j = 0; j < A; j++) ac = fmaf(uv_row[H + j], n[j], ac);
    if (p.quad) ac += quad_cost<A>(p, n);                   // grid-uniform
    if (RK == 0) nn_auv_step(sNN, P.nn_hidden, x, u);       // the learned model (NNAUVModel)
    else auv_step<RK>(P, x, u);
    S += auv_state_cost(p.cost_kind, p.q, g, p.ell, x) + ac;
}

template <bool PHILOX, int RK, int kAuvThreads, int kAuvCtasPerSm>
__global__ void __launch_bounds__(kAuvThreads, kAuvCtasPerSm)
rollout_auv_kernel(const __grid_constant__ RolloutParams p, const __grid_constant__ AuvParams P)
{
    constexpr int A = kAuvA, RS = Row<A>::RS, NW = kAuvThreads / 32;
    extern __shared__ float4 smem_f4[];
    float *smem = reinterpret_cast<float *>(smem_f4);
    const int TA = p.TA, TAp = (TA + 31) & ~31;
    float *sUV = smem;                    // [T][RS]
    float *sAcc = sUV + p.T * RS;         // [NW][TAp]
    float *sN = sAcc + NW * TAp;          // [TAp]
    float *sWork = sN + TAp;              // [TAp]
    float *sScale = sWork + TAp;          // [kMaxParts]
    float *sRed = sScale + kMaxParts;     // [64]
    float *sGX = sRed + 64;               // [32] goal at 0, initial state at 16 (broadcast reads instead of 26 registers)
    float4 *sScratch = reinterpret_cast<float4 *>(sGX + 32);    // [kAuvThreads]
    float *sNN = reinterpret_cast<float *>(sScratch + kAuvThreads);   // RK == 0: the learned model's blob

    const int ctrl = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (RK == 0)
        for (int i = tid; i < nn_auv_blob_floats(P.nn_hidden); i += kAuvThreads) sNN[i] = P.nn_blob[i];
    stage_sequence<A, PHILOX>(p, ctrl, sUV);
    if (tid < kAuvS) {
        sGX[tid] = p.goal[(p.goal_per_ctrl ? (size_t)ctrl * kAuvS : 0) + tid];
        sGX[16 + tid] = p.x_inline ? p.x0[tid] : p.x[(size_t)ctrl * kAuvS + tid];
    }
    const float *g = sGX, *x0 = sGX + 16;
    __syncthreads();
    const float C0 = stage_c0<A>(p, ctrl, sWork, sRed);

    float *costs = p.costs + (size_t)ctrl * p.K_local;
    const float *eps = PHILOX ? nullptr : p.eps + (size_t)ctrl * p.K_local * TA;
    // contiguous sample range per CTA at 32-sample granularity; every warp walks whole 32-sample groups so the
    // loop trip counts are warp-uniform
    const int n_w = (p.K_local + 31) >> 5;
    const int w_lo = (int)((long long)n_w * blockIdx.x / gridDim.x);
    const int w_hi = (int)((long long)n_w * (blockIdx.x + 1) / gridDim.x);
    const uint32_t stream = (uint32_t)ctrl;

    // ---- phase 1: rollout + cost -----------------------------------------------------------------------
    float bmin = kInf, bmax = -kInf;
    for (int wg = w_lo + warp; wg < w_hi && p.norm_mode != 2; wg += NW) {
        const int k = 32 * wg + lane;
        if (k >= p.K_local) continue;
        const uint32_t kg = (uint32_t)(p.k_offset + k);
        float x[kAuvS];
#pragma unroll
        for (int i = 0; i < kAuvS; i++) x[i] = x0[i];
        KahanSum Sk_;                               // per-step costs added with compensation (mppi_device.cuh)
        Sk_.init(C0);
        if (PHILOX) {
            // flattened [T][6] row: call c yields normals 4c .. 4c+3; two steps = three calls
            uint32_t call = 0;
            int t = 0;
            for (; t + 2 <= p.T; t += 2) {
                float z[12];
#pragma unroll
                for (int c = 0; c < 3; c++) normals4(call + c, kg, stream, p, &z[4 * c]);
                call += 3;
#pragma unroll
                for (int tt = 0; tt < 2; tt++) {
                    float n[A];
#pragma unroll
                    for (int j = 0; j < A; j++) n[j] = z[tt * A + j];
                    float S1 = 0.f;
                    auv_rollout_step<true, RK>(p, P, sNN, sUV + (t + tt) * RS, n, g, x, S1);
                    Sk_.add(S1);
                }
            }
            if (t < p.T) {
                float z[8];
                normals4(call, kg, stream, p, &z[0]);
                normals4(call + 1, kg, stream, p, &z[4]);
                float n[A];
#pragma unroll
                for (int j = 0; j < A; j++) n[j] = z[j];
                float S1 = 0.f;
                auv_rollout_step<true, RK>(p, P, sNN, sUV + t * RS, n, g, x, S1);
                Sk_.add(S1);
            }
        } else {
            const float2 *row = reinterpret_cast<const float2 *>(eps + (size_t)k * TA);   // TA = 6 T: 8-byte aligned rows
            for (int t = 0; t < p.T; t++) {
                float n[A];
#pragma unroll
                for (int j = 0; j < A / 2; j++) {
                    const float2 v = __ldg(row + t * (A / 2) + j);
                    n[2 * j] = v.x;
                    n[2 * j + 1] = v.y;
                }
                float S1 = 0.f;
                auv_rollout_step<false, RK>(p, P, sNN, sUV + t * RS, n, g, x, S1);
                Sk_.add(S1);
            }
        }
        Sk_.add(auv_state_cost(p.cost_kind, p.q, g, p.ell, x));      // terminal cost on top of step T-1's
        const float S = Sk_.s;
        costs[k] = S;
        bmin = fminf(bmin, S);
        bmax = fmaxf(bmax, S);
    }
    bmin = warp_min(bmin);
    bmax = -warp_min(-bmax);
    if (lane == 0) { sRed[warp] = bmin; sRed[32 + warp] = bmax; }
    __syncthreads();
    float beta_c = sRed[0], max_c = sRed[32];
#pragma unroll
    for (int w = 1; w < NW; w++) { beta_c = fminf(beta_c, sRed[w]); max_c = fmaxf(max_c, sRed[32 + w]); }
    __syncthreads();
    if (p.norm_mode == 1) {
        publish_minmax(p, ctrl, beta_c, max_c, sRed);
        return;
    }
    float beta_fixed = 0.f;
    const float nil = weight_scale(p, ctrl, beta_fixed);
    if (p.norm_mode == 2) beta_c = beta_fixed;

    // ---- phase 2: eta and the weighted noise sums ----------------------------------------------------------
    float eta = 0.f;
    float *myacc = sAcc + warp * TAp;
    for (int j = lane; j < TAp; j += 32) myacc[j] = 0.f;
    __syncwarp();
    if (PHILOX) {
        const int ncall = (TA + 3) >> 2, nchunk = (ncall + 7) >> 3;
        for (int ch = 0; ch < nchunk; ch++) {
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; i++) acc[i] = 0.f;
            for (int wg = w_lo + warp; wg < w_hi; wg += NW) {
                const int k = 32 * wg + lane;
                const float e = (k < p.K_local) ? sample_weight(costs[k], beta_c, nil) : 0.f;
                if (ch == 0) eta += e;
                if (__ballot_sync(0xffffffffu, e != 0.f) == 0u) continue;        // the whole group underflowed
                const uint32_t kg = (uint32_t)(p.k_offset + k);
#pragma unroll
                for (int c8 = 0; c8 < 8; c8++) {
                    if (ch * 8 + c8 < ncall) {
                        float z[4];
                        normals4((uint32_t)(ch * 8 + c8), kg, stream, p, z);
#pragma unroll
                        for (int i = 0; i < 4; i++) acc[4 * c8 + i] = fmaf(e, z[i], acc[4 * c8 + i]);
                    }
                }
            }
            const float r = warp_transpose_sum32(acc, lane);
            myacc[ch * 32 + lane] = r;
        }
    } else {
        // lane = column: 256 columns per pass, the group's non-zero weights broadcast one by one
        for (int cb = 0; cb < TA; cb += 256) {
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = 0.f;
            for (int wg = w_lo + warp; wg < w_hi; wg += NW) {
                const int k = 32 * wg + lane;
                const float e = (k < p.K_local) ? sample_weight(costs[k], beta_c, nil) : 0.f;
                if (cb == 0) eta += e;
                unsigned live = __ballot_sync(0xffffffffu, e != 0.f);
                while (live) {
                    const int j = __ffs(live) - 1;
                    live &= live - 1;
                    const float ej = __shfl_sync(0xffffffffu, e, j);
                    const float *row = eps + (size_t)(32 * wg + j) * TA + cb;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int col = 32 * i + lane;
                        if (cb + col < TA) acc[i] = fmaf(ej, __ldg(row + col), acc[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (cb + 32 * i + lane < TAp) myacc[cb + 32 * i + lane] = acc[i];
        }
    }
    eta = warp_sum(eta);
    if (lane == 0) sRed[warp] = eta;
    __syncthreads();
    float eta_c = 0.f;
#pragma unroll
    for (int w = 0; w < NW; w++) eta_c += sRed[w];
    for (int j = tid; j < TA; j += kAuvThreads) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; w++) s += sAcc[w * TAp + j];
        sN[j] = s;
    }
    __syncthreads();
    publish_and_finish<A, PHILOX>(p, ctrl, beta_c, eta_c, sN, sWork, sScale, sRed, sScratch, kAuvThreads);
}

// AUVModel.build_step_graph on a batch: state [kst][13] (kst in {1, k}), action [k][6] -> next [k][13]
__global__ void __launch_bounds__(256) auv_predict_kernel(const __grid_constant__ AuvParams P, int kst, int k,
                                                          const float *state, const float *action, float *out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
        float x[kAuvS], u[kAuvA];
        const float *xs = state + (size_t)(kst == 1 ? 0 : i) * kAuvS;
#pragma unroll
        for (int j = 0; j < kAuvS; j++) x[j] = xs[j];
#pragma unroll
        for (int j = 0; j < kAuvA; j++) u[j] = action[(size_t)i * kAuvA + j];
        if (P.rk == 0) nn_auv_step(P.nn_blob, P.nn_hidden, x, u);      // weights straight from global memory (L1-cached)
        else if (P.rk == 2) auv_step<2>(P, x, u);
        else if (P.rk == 4) auv_step<4>(P, x, u);
        else auv_step<1>(P, x, u);
#pragma unroll
        for (int j = 0; j < kAuvS; j++) out[(size_t)i * kAuvS + j] = x[j];
    }
}

struct QuatCostArgs { float q[kAuvS]; float g[kAuvS]; float ell[12]; int kind; };

__global__ void __launch_bounds__(256) auv_cost_kernel(const __grid_constant__ QuatCostArgs a, int k, const float *state, float *out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
        float x[kAuvS], g[kAuvS];
#pragma unroll
        for (int j = 0; j < kAuvS; j++) { x[j] = state[(size_t)i * kAuvS + j]; g[j] = a.g[j]; }
        out[i] = auv_state_cost(a.kind, a.q, g, a.ell, x);
    }
}

// -------------------------------------------------------------------------------------------------
// Host launchers
// -------------------------------------------------------------------------------------------------
template <bool PHILOX, int RK, int THREADS, int CTAS>
static cudaError_t launch_auv_V(const RolloutParams &p, const AuvParams &P, int num_sms, cudaStream_t st, int *grid_x_out)
{
    constexpr int RS = Row<kAuvA>::RS, NW = THREADS / 32;
    const int TAp = (p.TA + 31) & ~31;
    const size_t smem = sizeof(float) * ((size_t)p.T * RS + (size_t)NW * TAp + 2 * TAp + kMaxParts + 64 + 32) + sizeof(float4) * THREADS +
                        (RK == 0 ? sizeof(float) * nn_auv_blob_floats(P.nn_hidden) : 0);
    int per_ctrl = num_sms * CTAS / (p.n_ctrl > 0 ? p.n_ctrl : 1);
    if (per_ctrl < 1) per_ctrl = 1;
    const int need = (p.K_local + 63) / 64;        // at least two warps of samples per CTA
    int gx = need < per_ctrl ? need : per_ctrl;
    if (gx > kMaxParts) gx = kMaxParts;
    if (gx < 1) gx = 1;
    if (grid_x_out) *grid_x_out = gx;
    cudaError_t err = ensure_dyn_smem<rollout_auv_kernel<PHILOX, RK, THREADS, CTAS>>(smem);
    if (err != cudaSuccess) return err;
    rollout_auv_kernel<PHILOX, RK, THREADS, CTAS><<<dim3(gx, p.n_ctrl), THREADS, smem, st>>>(p, P);
    return cudaGetLastError();
}

// CTA shapes: 0 = 256 x 2 (128 registers), 1 = 128 x 3 (168), 2 = 256 x 1 (255).  MPPI_AUV_GEOM overrides the
// measured default (developer knob).
static int auv_geometry(bool philox, int rk)
{
    static int forced = -2;
    if (forced == -2) {
        const char *e = getenv("MPPI_AUV_GEOM");
        forced = e ? atoi(e) : -1;
    }
    if (forced >= 0 && forced <= 2) return forced;
    // measured at K = 262144, T = 50 (scripts_dev/auv_bench.py): the Philox variants want ~240 registers (0.30 / 0.42 /
    // 0.79 ms for rk 1 / 2 / 4 against 0.37 / 0.68 / 2.15 ms at 128); the injected-noise variants fit 128 up to rk 2
    if (philox) return 2;
    return rk == 4 ? 1 : 0;
}

template <bool PHILOX, int RK>
static cudaError_t launch_auv_G(const RolloutParams &p, const AuvParams &P, int num_sms, cudaStream_t st, int *grid_x_out)
{
    switch (auv_geometry(PHILOX, RK)) {
        case 1: return launch_auv_V<PHILOX, RK, 128, 3>(p, P, num_sms, st, grid_x_out);
        case 2: return launch_auv_V<PHILOX, RK, 256, 1>(p, P, num_sms, st, grid_x_out);
        default: return launch_auv_V<PHILOX, RK, 256, 2>(p, P, num_sms, st, grid_x_out);
    }
}

template <bool PHILOX>
static cudaError_t launch_auv_R(const RolloutParams &p, const AuvParams &P, int num_sms, cudaStream_t st, int *grid_x_out)
{
    if (P.rk == 0) return launch_auv_G<PHILOX, 0>(p, P, num_sms, st, grid_x_out);
    if (P.rk == 2) return launch_auv_G<PHILOX, 2>(p, P, num_sms, st, grid_x_out);
    if (P.rk == 4) return launch_auv_G<PHILOX, 4>(p, P, num_sms, st, grid_x_out);
    return launch_auv_G<PHILOX, 1>(p, P, num_sms, st, grid_x_out);
}

cudaError_t launch_rollout_auv(RolloutParams p, const AuvParams &P, bool philox, int num_sms, cudaStream_t st, int *grid_x_out)
{
    return philox ? launch_auv_R<true>(p, P, num_sms, st, grid_x_out) : launch_auv_R<false>(p, P, num_sms, st, grid_x_out);
}

cudaError_t launch_auv_predict(const AuvParams &P, int kst, int k, const float *state, const float *action, float *out, cudaStream_t st)
{
    int blocks = (k + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    auv_predict_kernel<<<blocks, 256, 0, st>>>(P, kst, k, state, action, out);
    return cudaGetLastError();
}

cudaError_t launch_auv_cost(int kind, int k, const float *q, const float *goal, const float *ell, const float *state, float *out,
                            cudaStream_t st)
{
    QuatCostArgs a;
    for (int i = 0; i < kAuvS; i++) { a.q[i] = q ? q[i] : 0.f; a.g[i] = goal ? goal[i] : 0.f; }
    for (int i = 0; i < 12; i++) a.ell[i] = ell ? ell[i] : 0.f;
    a.kind = kind;
    int blocks = (k + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    auv_cost_kernel<<<blocks, 256, 0, st>>>(a, k, state, out);
    return cudaGetLastError();
}

}  // namespace mppi
