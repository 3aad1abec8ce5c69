// MPPI update with the learned-MLP dynamics on tcgen05 tensor cores (BASELINE config 4), plus the
// batched one-step `predict` stage used by ModelBase-style callers and by the unit tests.
// Device building blocks: mppi_mlp.cuh.  The update logic (softmin partials, merge, shift) is the
// same code the point-mass kernels use (mppi_update.cuh).
#define MPPI_TAIL_INLINE        // setmaxnreg kernel: no function calls (mppi_update.cuh)
#include "mppi_mlp.cuh"
#include "mppi_internal.h"
#include "mppi_update.cuh"

namespace mppi {

constexpr int kMlpListCap = 256;       // entries of a warp's non-zero-weight list

// -------------------------------------------------------------------------------------------------
// predict: next[k][s] = mlp(state[k|1][s], action[k][a])   (one CTA per 128 samples)
// -------------------------------------------------------------------------------------------------
template <int S, int A>
__global__ void __launch_bounds__(kMlpThreads) mlp_predict_kernel(MlpParams mp, int kst, int k, const float *state,
                                                                  const float *action, float *out)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *sW = smem_raw;
    float *sF = reinterpret_cast<float *>(sW + kWBlobBytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sF + kFvecFloats);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + kMlpNumBars);
    MlpTile t;
    mlp_tile_init(t, mp, sW, sF, bars, tslot);
    if (threadIdx.x < kMlpRows) {
        const int r = blockIdx.x * kMlpRows + threadIdx.x;
        float x[S], u[A];
#pragma unroll
        for (int i = 0; i < S; i++) x[i] = (r < k) ? state[(size_t)(kst == 1 ? 0 : r) * S + i] : 0.f;
#pragma unroll
        for (int i = 0; i < A; i++) u[i] = (r < k) ? action[(size_t)r * A + i] : 0.f;
        mlp_row_begin<S, A>(t, x, u);
        mlp_row_layer1(t, 0);
        mlp_row_layer2(t, 0);
        mlp_row_finish<S>(t, x);
        if (r < k) {
#pragma unroll
            for (int i = 0; i < S; i++) out[(size_t)r * S + i] = x[i];
        }
    } else if (threadIdx.x < 32 * kMlpEpiWarps) {
        mlp_helper_loop(t, 1);
    } else {
        mlp_mma_loop(t, 1);
    }
    mlp_tile_fini(t);
}

// -------------------------------------------------------------------------------------------------
// fused rollout: same structure as rollout_philox_kernel (phase 1 costs, block min, phase 2 weighted
// sums by regenerating / re-reading the noise, last-CTA merge) with the model step on tensor cores.
// Row thread r = sample row r of the 128-row tile; every row thread takes part in every hand-over
// with the MMA warp, so out-of-range rows roll a dummy sample.
// -------------------------------------------------------------------------------------------------
template <int A, bool PHILOX>
__device__ __forceinline__ void noise4(const RolloutParams &p, const float *eps_row, uint32_t call, uint32_t kg,
                                       uint32_t stream, bool valid, float (&z)[4])
{
    if (PHILOX) {
        normals4(call, kg, stream, p, z);
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int idx = 4 * (int)call + j;
            z[j] = (valid && idx < p.TA) ? __ldg(eps_row + idx) : 0.f;
        }
    }
}

// state cost of the selected functor (grid-uniform branch; the MLP step is not bound by the CUDA cores)
template <int S>
__device__ __forceinline__ float mlp_state_cost(const RolloutParams &p, const float (&x)[S], const float (&g)[S], const float (&q)[S])
{
    if (p.cost_kind == 1 && S == 4) return ellipse_cost(p.ell, x[0], x[1], x[S > 2 ? 2 : 0], x[S > 3 ? 3 : 0]);
    float c = 0.f;
#pragma unroll
    for (int i = 0; i < S; i++) {
        const float d = x[i] - g[i];
        c = fmaf(q[i] * d, d, c);
    }
    return c;
}

template <int A, bool PHILOX>
__global__ void __launch_bounds__(kMlpThreads, 2) rollout_mlp_kernel(const __grid_constant__ RolloutParams p, MlpParams mp)
{
    constexpr int S = 2 * A;
    constexpr int RS = Row<A>::RS, H = Row<A>::H;
    constexpr int NW = kMlpThreads / 32;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int TA = p.TA, TAp = (TA + 31) & ~31;
    uint8_t *sW = smem_raw;
    float *sF = reinterpret_cast<float *>(sW + kWBlobBytes);
    float *sUV = sF + kFvecFloats;                // [T][RS]
    float *sAcc = sUV + p.T * RS;                 // [NW][TAp]
    float *sN = sAcc + NW * TAp;                  // [TAp]
    float *sWork = sN + TAp;                      // [TAp]
    float *sScale = sWork + TAp;                  // [kMaxParts]
    float *sRed = sScale + kMaxParts;             // [64]
    float4 *sScratch = reinterpret_cast<float4 *>(sRed + 64);     // [kMlpThreads] merge scratch
    uint2 *sList = reinterpret_cast<uint2 *>(sScratch + kMlpThreads);   // [NW][kMlpListCap] non-zero-weight samples
    uint64_t *bars = reinterpret_cast<uint64_t *>(sList + NW * kMlpListCap);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + kMlpNumBars);

    const int ctrl = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    MlpTile t;
    stage_sequence<A, PHILOX>(p, ctrl, sUV);
    mlp_tile_init(t, mp, sW, sF, bars, tslot);    // ends with a CTA barrier: sUV is visible
    const float C0 = stage_c0<A>(p, ctrl, sWork, sRed);

    float *costs = p.costs + (size_t)ctrl * p.K_local;
    const float *eps = PHILOX ? nullptr : p.eps + (size_t)ctrl * p.K_local * TA;
    // contiguous sample range of this CTA, in tiles of 128 rows
    const int n_t = (p.K_local + kMlpRows - 1) / kMlpRows;
    const int t_lo = (int)((long long)n_t * blockIdx.x / gridDim.x);
    const int t_hi = (int)((long long)n_t * (blockIdx.x + 1) / gridDim.x);
    const int nblk = (p.T + 3) >> 2;
    const uint32_t stream = (uint32_t)ctrl;

    // ---- phase 1: rollout + cost -----------------------------------------------------------------
    float bmin = kInf, bmax = -kInf;
    if (p.norm_mode == 2) {
        // weight pass of a normalised update: costs are in HBM, no rollout
    } else if (warp < kMlpRowWarps) {
        reg_inc<kMlpRegsRow>();
        float g[S], q[S], x0[S];
        {
            const float *gp = p.goal + (p.goal_per_ctrl ? (size_t)ctrl * S : 0);
            const float *xp = p.x + (size_t)ctrl * S;
#pragma unroll
            for (int i = 0; i < S; i++) {
                g[i] = gp[i];
                q[i] = p.q[i];
                x0[i] = p.x_inline ? p.x0[i] : xp[i];
            }
        }
        for (int tile = t_lo; tile < t_hi; tile++) {
            const int k = tile * kMlpRows + tid;
            const bool valid = k < p.K_local;
            const uint32_t kg = (uint32_t)(p.k_offset + k);
            const float *eps_row = PHILOX ? nullptr : eps + (size_t)(valid ? k : 0) * TA;
            float x[S];
#pragma unroll
            for (int i = 0; i < S; i++) x[i] = x0[i];
            float Sk = C0;
            float z[4 * A], zn[4 * A];
#pragma unroll
            for (int c = 0; c < A; c++) {
                float z4[4];
                noise4<A, PHILOX>(p, eps_row, (uint32_t)c, kg, stream, valid, z4);
#pragma unroll
                for (int j = 0; j < 4; j++) z[4 * c + j] = z4[j];
            }
            for (int tb = 0; tb < nblk; tb++) {
                const bool more = tb + 1 < nblk;          // block-uniform
#pragma unroll
                for (int tt = 0; tt < 4; tt++) {
                    const int ts = 4 * tb + tt;
                    if (ts >= p.T) break;                 // block-uniform
                    const float *uv = sUV + ts * RS;
                    float u[A];
                    float ac = 0.f;
#pragma unroll
                    for (int j = 0; j < A; j++) {
                        const float n = z[tt * A + j];
                        float e = n;
                        if (PHILOX) {
                            e = 0.f;
#pragma unroll
                            for (int l = 0; l < A; l++) e = fmaf(p.sigma[j * A + l], z[tt * A + l], e);
                        }
                        u[j] = uv[j] + e;
                        ac = fmaf(uv[H + j], n, ac);
                    }
                    if (p.quad) {                         // Python-twin noise cost (grid-uniform branch)
                        float nn[A];
#pragma unroll
                        for (int j = 0; j < A; j++) nn[j] = z[tt * A + j];
                        ac += quad_cost<A>(p, nn);
                    }
                    mlp_row_begin<S, A>(t, x, u);
                    // in the shadow of layer 1: state cost of the state the step started from (it is
                    // step ts-1's q(x_ts)), and this step's action cost
                    if (ts > 0) {
                        Sk += mlp_state_cost<S>(p, x, g, q);
                    }
                    Sk += ac;
                    mlp_row_layer1(t, 0);
                    // in the shadow of layer 2 (the long MMA): this step's share of the next block's noise
                    if (more) {
#pragma unroll
                        for (int c = (tt * A) / 4; c < ((tt + 1) * A) / 4; c++) {
                            float z4[4];
                            noise4<A, PHILOX>(p, eps_row, (uint32_t)((tb + 1) * A + c), kg, stream, valid, z4);
#pragma unroll
                            for (int j = 0; j < 4; j++) zn[4 * c + j] = z4[j];
                        }
                    }
                    mlp_row_layer2(t, 0);
                    mlp_row_finish<S>(t, x);
                }
#pragma unroll
                for (int i = 0; i < 4 * A; i++) z[i] = zn[i];
            }
            {
                const float c = mlp_state_cost<S>(p, x, g, q);   // q(x_T) of step T-1 plus the terminal cost (src/controller_base.cpp:271-272)
                Sk += c;
                Sk += c;
            }
            if (valid) {
                costs[k] = Sk;
                bmin = fminf(bmin, Sk);
                bmax = fmaxf(bmax, Sk);
            }
        }
        reg_dec<kMlpRegsLaunch>();
    } else if (warp < kMlpEpiWarps) {
        reg_dec<kMlpRegsHelper>();
        mlp_helper_loop(t, (t_hi - t_lo) * p.T);
        reg_inc<kMlpRegsLaunch>();
    } else {
        reg_dec<kMlpRegsMma>();
        mlp_mma_loop(t, (t_hi - t_lo) * p.T);
        reg_inc<kMlpRegsLaunch>();
    }
    mlp_tile_fini(t);
    bmin = warp_min(bmin);
    bmax = -warp_min(-bmax);
    if (lane == 0) { sRed[warp] = bmin; sRed[32 + warp] = bmax; }
    __syncthreads();
    float beta_c = sRed[0], max_c = sRed[32];
#pragma unroll
    for (int w = 1; w < NW; w++) { beta_c = fminf(beta_c, sRed[w]); max_c = fmaxf(max_c, sRed[32 + w]); }
    __syncthreads();
    if (p.norm_mode == 1) {                 // cost pass of a normalised update: publish (min, max) and stop
        publish_minmax(p, ctrl, beta_c, max_c, sRed);
        return;
    }
    float beta_fixed = 0.f;
    const float nil = weight_scale(p, ctrl, beta_fixed);
    if (p.norm_mode == 2) beta_c = beta_fixed;

    // ---- phase 2: sum_k e_k n_k (n = z regenerated, or eps re-read); all five warps take samples ----
    const int ncall = (TA + 3) >> 2;
    const int nchunk = (ncall + 7) >> 3;
    const int k_lo = t_lo * kMlpRows, k_hi = min(t_hi * kMlpRows, p.K_local);
    float eta = 0.f;
    // zero-weight compaction as in rollout_philox_kernel: a weight that underflowed to 0.0f adds exactly
    // nothing, so only the other samples are revisited (per-warp ordered lists; dense loop when no weight
    // of the CTA can underflow)
    const bool sparse = (max_c - beta_c) * fabsf(nil) > kWeightCutLog2;         // CTA-uniform
    float *myacc = sAcc + warp * TAp;
    for (int j = lane; j < TAp; j += 32) myacc[j] = 0.f;
    uint2 *wlist = sList + warp * kMlpListCap;
    for (int kb = k_lo + tid - lane; kb < k_hi; kb += kMlpThreads * (kMlpListCap / 32)) {
        const int nit = min(kMlpListCap / 32, (k_hi - kb + kMlpThreads - 1) / kMlpThreads);   // warp-uniform
        int cnt = 0;
#pragma unroll 1
        for (int it = 0; it < nit; it++) {
            const int k = kb + it * kMlpThreads + lane;
            float e = 0.f;
            if (k < k_hi) e = sample_weight(costs[k], beta_c, nil);
            eta += e;
            const bool keep = sparse ? (e != 0.f) : (k < k_hi);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) wlist[cnt + __popc(m & ((1u << lane) - 1u))] = make_uint2((uint32_t)k, __float_as_uint(e));
            cnt += __popc(m);
        }
        __syncwarp();
        if (cnt == 0) continue;                                          // warp-uniform
        for (int ch = 0; ch < nchunk; ch++) {
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; i++) acc[i] = 0.f;
            for (int i0 = 0; i0 < cnt; i0 += 32) {
                const bool live = i0 + lane < cnt;
                const uint2 ent = live ? wlist[i0 + lane] : make_uint2(0u, 0u);     // padding lanes: weight 0
                const int k = (int)ent.x;
                const uint32_t kg = (uint32_t)(p.k_offset + k);
                const float *eps_row = PHILOX ? nullptr : eps + (size_t)k * TA;
                const float e = __uint_as_float(ent.y);
#pragma unroll
                for (int c8 = 0; c8 < 8; c8++) {
                    if (ch * 8 + c8 < ncall) {
                        float z[4];
                        noise4<A, PHILOX>(p, eps_row, (uint32_t)(ch * 8 + c8), kg, stream, live, z);
#pragma unroll
                        for (int j = 0; j < 4; j++) acc[4 * c8 + j] = fmaf(e, z[j], acc[4 * c8 + j]);
                    }
                }
            }
            const float r = warp_transpose_sum32(acc, lane);
            myacc[ch * 32 + lane] += r;
        }
        __syncwarp();
    }
    eta = warp_sum(eta);
    if (lane == 0) sRed[warp] = eta;
    __syncthreads();
    float eta_c = 0.f;
#pragma unroll
    for (int w = 0; w < NW; w++) eta_c += sRed[w];
    for (int j = tid; j < TA; j += kMlpThreads) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; w++) s += sAcc[w * TAp + j];
        sN[j] = s;
    }
    __syncthreads();
    // a CTA without samples must not win the min: beta_c = +inf is skipped by merge_parts
    publish_and_finish<A, PHILOX>(p, ctrl, beta_c, eta_c, sN, sWork, sScale, sRed, sScratch, kMlpThreads);
}

// -------------------------------------------------------------------------------------------------
// host launchers
// -------------------------------------------------------------------------------------------------
static size_t mlp_predict_smem() { return kWBlobBytes + sizeof(float) * kFvecFloats + kMlpNumBars * 8 + 16 + 128; }

static size_t mlp_rollout_smem(int A, int T, int TA)
{
    const int H = (A + 1) & ~1, RS = (2 * H + 3) & ~3, TAp = (TA + 31) & ~31, NW = kMlpThreads / 32;
    return kWBlobBytes + sizeof(float) * (kFvecFloats + (size_t)T * RS + (size_t)NW * TAp + 2 * TAp + kMaxParts + 32) +
           sizeof(float) * 32 + sizeof(float4) * kMlpThreads + sizeof(uint2) * NW * kMlpListCap + kMlpNumBars * 8 + 16 + 128;
}

#define MPPI_DISPATCH_MLP_A(a, ...)              \
    switch (a) {                                 \
        case 1: { constexpr int A_ = 1; __VA_ARGS__; } break; \
        case 2: { constexpr int A_ = 2; __VA_ARGS__; } break; \
        case 3: { constexpr int A_ = 3; __VA_ARGS__; } break; \
        case 4: { constexpr int A_ = 4; __VA_ARGS__; } break; \
        case 5: { constexpr int A_ = 5; __VA_ARGS__; } break; \
        default: return cudaErrorInvalidValue;   \
    }

cudaError_t launch_mlp_predict(const MlpParams &mp, int kst, int k, const float *state, const float *action, float *out,
                               cudaStream_t st)
{
    const size_t smem = mlp_predict_smem();
    const int grid = (k + kMlpRows - 1) / kMlpRows;
    MPPI_DISPATCH_MLP_A(mp.a, {
        cudaError_t err = ensure_dyn_smem<mlp_predict_kernel<2 * A_, A_>>(smem);
        if (err != cudaSuccess) return err;
        mlp_predict_kernel<2 * A_, A_><<<grid, kMlpThreads, smem, st>>>(mp, kst, k, state, action, out);
    });
    return cudaGetLastError();
}

int mlp_grid_x(int K_local, int n_ctrl, int num_sms)
{
    int per_ctrl = (num_sms * 2) / (n_ctrl > 0 ? n_ctrl : 1);   // 2 CTAs (2 x 256 TMEM columns) per SM
    if (per_ctrl < 1) per_ctrl = 1;
    const int need = (K_local + kMlpRows - 1) / kMlpRows;
    int gx = need < per_ctrl ? need : per_ctrl;
    if (gx > kMaxParts) gx = kMaxParts;
    return gx < 1 ? 1 : gx;
}

cudaError_t launch_rollout_mlp(RolloutParams p, const MlpParams &mp, int a, bool philox, int num_sms, cudaStream_t st,
                               int *grid_x_out)
{
    const int gx = mlp_grid_x(p.K_local, p.n_ctrl, num_sms);
    if (grid_x_out) *grid_x_out = gx;
    const size_t smem = mlp_rollout_smem(a, p.T, p.TA);
    dim3 grid(gx, p.n_ctrl);
    MPPI_DISPATCH_MLP_A(a, {
        cudaError_t err;
        if (philox) {
            err = ensure_dyn_smem<rollout_mlp_kernel<A_, true>>(smem);
            if (err != cudaSuccess) return err;
            rollout_mlp_kernel<A_, true><<<grid, kMlpThreads, smem, st>>>(p, mp);
        } else {
            err = ensure_dyn_smem<rollout_mlp_kernel<A_, false>>(smem);
            if (err != cudaSuccess) return err;
            rollout_mlp_kernel<A_, false><<<grid, kMlpThreads, smem, st>>>(p, mp);
        }
    });
    return cudaGetLastError();
}

#ifdef MPPI_MLP_TRACE
extern "C" int mppi_debug_mlp_trace(long long *out /*[2][1024]*/, int *n /*[2]*/)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_mlp_trace, sizeof(long long) * 2 * 1024);
    n[0] = n[1] = 16 * 64;
    return 0;
}
#endif

// Pack Keras-layout weights ([in][out], fp32) and biases into the bf16 canonical K-major blob the
// kernels stage: B[n][k] = W[k][n], with the bias of each layer in the K row that meets the constant 1.
void mlp_pack_weights(int s, int a, const float *W1, const float *b1, const float *W2, const float *b2,
                      const float *W3, const float *b3, void *blob_host)
{
    __nv_bfloat16 *b = static_cast<__nv_bfloat16 *>(blob_host);
    const int in = s + a;
    for (int i = 0; i < kWBlobBytes / 2; i++) b[i] = __float2bfloat16(0.f);
    uint8_t *base = static_cast<uint8_t *>(blob_host);
    auto put = [&](int off, int n, int k, int N, float v) {
        *reinterpret_cast<__nv_bfloat16 *>(base + off + canon_offset_bytes(n, k, N)) = __float2bfloat16(v);
    };
    for (int n = 0; n < kMlpH; n++) {
        for (int k = 0; k < in; k++) put(0, n, k, kMlpH, W1[k * kMlpH + n]);
        put(0, n, in, kMlpH, b1[n]);
        for (int k = 0; k < kMlpH; k++) put(kW1Bytes, n, k, kMlpH, W2[k * kMlpH + n]);
        put(kW1Bytes, n, kMlpH, kMlpH, b2[n]);
    }
    for (int n = 0; n < s; n++) {
        for (int k = 0; k < kMlpH; k++) put(kW1Bytes + kW2Bytes, n, k, kMlpNout, W3[k * s + n]);
        put(kW1Bytes + kW2Bytes, n, kMlpH, kMlpNout, b3[n]);
    }
}

}  // namespace mppi
