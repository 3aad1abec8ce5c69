// Philox-mode update kernels in the superposition form (mppi_linear.cuh): diagonal Sigma, q > 0, StaticCost, no
// noise-quadratic action-cost term, T <= kFastMaxT — BASELINE configs 1, 2, 3 and 5.  Six FMAs per axis-step instead of
// nine, compile-time Philox round count (7 or 10), generator without its last multiply.
//
//   rollout_philox_fast_kernel      long rows (config 3: T*a = 300): the noise is regenerated for the weighted sum
//                                   (mppi_philox_sum.cuh), as in rollout_philox_kernel
//   rollout_philox_resident_kernel  short rows (configs 1, 2, 5: T*a <= ~150): every warp owns a 32-sample tile whose
//                                   normals stay in shared memory between the rollout and the weighted sum — nothing is
//                                   generated twice, whatever the weights; warps run on their own (online max-shifted
//                                   softmin per warp, no CTA barrier between the prologue and the final merge)
//
// Reference maths: /root/reference/src/controller_base.cpp:166-329, src/model_base.cpp:53-82, src/cost_base.cpp:37-68.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "mppi_device.cuh"
#include "mppi_internal.h"
#include "mppi_linear.cuh"
#include "mppi_philox_sum.cuh"
#include "mppi_update.cuh"

namespace mppi {

template <int A>
__device__ __forceinline__ void vec_from4(const float *z, Vec<A> &n)
{
#pragma unroll
    for (int i = 0; i < A / 2; i++) n.pr[i] = make_float2(z[2 * i], z[2 * i + 1]);
    n.sc = z[A - 1];
}

// Rollout of one sample in the superposition form: returns sum_t w_t (P_t^2 + V_t^2) + sum_t L_t . n_t.
// STORE: the normals of every call are also written to `row[call]` (the resident kernel's tile row of this lane).
template <int A, int R, bool STORE>
__device__ __forceinline__ float fast_rollout(const RolloutParams &p, const uint4 *sTab, const float *sL, const PhiloxSample ps,
                                              const FastConsts<A> &fc, float4 *row)
{
    constexpr int RS = (A + 3) & ~3;
    const int nfull = p.T >> 2, trem = p.T & 3, ncall = (p.TA + 3) >> 2;
    Vec<A> P, V;
    P.fill(0.f);
    V.fill(0.f);
    // The terms of a 4-step block are summed from zero and the block sums are added with compensation (KahanSum): the
    // sum of a few hundred terms is then as good as its last rounding, instead of collecting one rounding at ulp(S) per
    // term.  On 4096 independent K = 1024 controllers (config 5) the worst update error drops from 1.1e-5 to 3.5e-6
    // (fp32 emulation of this kernel; the reference's own fp32 arithmetic: 4.9e-5).
    CostAcc Sq, Sl;
    KahanSum S;
    S.init(0.f);
    const float *l = sL;
    uint32_t call = 0;
    for (int tb = 0; tb < nfull; tb++) {            // full blocks of 4 steps = A Philox calls, no guards
        float z[4 * A];
#pragma unroll
        for (int c = 0; c < A; c++) {
            const float4 v = normals4_fast<R>(sTab, call + c, ps, p);
            MPPI_CHECK((int)(call + c) < ncall);
            if (STORE) row[call + c] = v;
            z[4 * c] = v.x; z[4 * c + 1] = v.y; z[4 * c + 2] = v.z; z[4 * c + 3] = v.w;
        }
        call += A;
        Sq.zero();
        Sl.zero();
#pragma unroll
        for (int tt = 0; tt < 4; tt++) {
            Vec<A> n;
            vec_from4<A>(&z[tt * A], n);
            fast_step<A>(P, V, Sq, Sl, l + tt * RS, n, fc);
        }
        {
            const float2 t2 = __fadd2_rn(Sq.a2, Sl.a2);
            S.add((t2.x + t2.y) + (Sq.a + Sl.a));
        }
        l += 4 * RS;
    }
    Sq.zero();
    Sl.zero();
    if (trem) {                                     // tail: T % 4 steps
        float z[4 * A];
#pragma unroll
        for (int c = 0; c < A; c++)
            if ((int)(call + c) < ncall) {
                const float4 v = normals4_fast<R>(sTab, call + c, ps, p);
                if (STORE) row[call + c] = v;
                z[4 * c] = v.x; z[4 * c + 1] = v.y; z[4 * c + 2] = v.z; z[4 * c + 3] = v.w;
            }
#pragma unroll
        for (int tt = 0; tt < 3; tt++)
            if (tt < trem) {
                Vec<A> n;
                vec_from4<A>(&z[tt * A], n);
                fast_step<A>(P, V, Sq, Sl, l + tt * RS, n, fc);
            }
    }
    fast_terminal<A>(P, V, Sq);                     // terminal cost on top of step T-1's (src/controller_base.cpp:271-272)
    S.add(Sq.total() + Sl.total());
    return S.s;
}

// -------------------------------------------------------------------------------------------------
// Long rows: noise regenerated for the weighted sum
// -------------------------------------------------------------------------------------------------
template <int A, int R>
__global__ void __launch_bounds__(kPhiloxThreads, kPhiloxCtasPerSm)
rollout_philox_fast_kernel(const __grid_constant__ RolloutParams p)
{
    constexpr int RS = (A + 3) & ~3;
    constexpr int NW = kPhiloxThreads / 32;
    extern __shared__ float4 smem_f4[];
    const int TA = p.TA, TAp = (TA + 31) & ~31;
    PhiloxSmem sm;
    sm.carve(reinterpret_cast<float *>(smem_f4), p.T, RS, TAp);
    float *sRed = sm.sRed;

    const int ctrl = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ unsigned s_slot;                  // hardware warp slot of the CTA's first warp (read after the barriers below)
    if (tid == 0) asm volatile("mov.u32 %0, %%warpid;" : "=r"(s_slot));
    trace_stamp(p, ctrl, 0);
    for (int c = tid; c < ((TA + 3) >> 2); c += kPhiloxThreads) sm.sTab[c] = philox_call_table((uint32_t)c, (uint32_t)ctrl, p);
    const uint32_t phA = philox_uniform_A(p);
    // per-controller tables: L_t into sUV; scratch (2 A T + T floats) in the per-warp sum rows, which phase 2 initialises itself
    if (p.norm_mode != 2) {
        float C = build_linear_tables<A, true>(p, ctrl, sm.sUV, sm.sAcc, sRed);
        C += stage_c0<A>(p, ctrl, sm.sWork, sRed);
        if (blockIdx.x == 0 && tid == 0) p.cost_base[ctrl] = C;      // costs[] holds S_k - C (see RolloutParams::cost_base)
    }
    FastConsts<A> fc;
    fc.init(p);
    __syncthreads();
    trace_stamp(p, ctrl, 1);

    float *costs = p.costs + (size_t)ctrl * p.K_local;
    const int n_w = (p.K_local + 31) >> 5;
    const int w_lo = (int)((long long)n_w * blockIdx.x / gridDim.x);
    const int w_hi = (int)((long long)n_w * (blockIdx.x + 1) / gridDim.x);
    // Which warps idle in the CTA's last (partly filled) iteration: the kernel is bound by the schedulers' dispatch rate, a
    // warp runs on scheduler %warpid mod 4, and the two CTAs of an SM would both park their idle warps on the same schedulers
    // (a K/8 shard of config 3 is one iteration of 13.8 tiles per CTA: 8 + 8 + 6 + 6 busy warps per scheduler instead of four
    // times 7).  The CTA in the upper half of the SM's warp slots therefore rotates its warp numbering by the idle count.  Only
    // the placement changes: samples, list order and summation order are functions of the virtual warp index.
    const int idle = (NW - (w_hi - w_lo) % NW) % NW;
    const int vwarp = (warp + ((s_slot >> 4) & 1) * idle) % NW;
    const int vtid = 32 * vwarp + lane;
    const int kfirst = 32 * w_lo + vtid;
    const int kend = min(p.K_local, 32 * w_hi);

    // ---- phase 1: rollout + cost ---------------------------------------------------------------
    float bmin = kInf, bmax = -kInf;
    for (int k = kfirst; k < kend && p.norm_mode != 2; k += kPhiloxThreads) {   // weight pass of a normalised update: costs are in HBM
        const PhiloxSample ps = philox_sample(phA, (uint32_t)(p.k_offset + k));
        const float Sk = fast_rollout<A, R, false>(p, sm.sTab, sm.sUV, ps, fc, nullptr);
        MPPI_CHECK(k >= 0 && k < p.K_local);
        costs[k] = Sk;
        bmin = fminf(bmin, Sk);
        bmax = fmaxf(bmax, Sk);
    }
    bmin = warp_min(bmin);
    bmax = -warp_min(-bmax);
    if (lane == 0) { sRed[warp] = bmin; sRed[32 + warp] = bmax; }
    __syncthreads();
    float beta_c = sRed[0], max_c = sRed[32];
#pragma unroll
    for (int w = 1; w < NW; w++) { beta_c = fminf(beta_c, sRed[w]); max_c = fmaxf(max_c, sRed[32 + w]); }
    __syncthreads();
    trace_stamp(p, ctrl, 2);
    if (p.norm_mode == 1) {                 // cost pass of a normalised update: publish (min, max) and stop
        publish_minmax(p, ctrl, beta_c, max_c, sRed);
        return;
    }
    float beta_fixed = 0.f;
    const float nil = weight_scale(p, ctrl, beta_fixed);
    if (p.norm_mode == 2) beta_c = beta_fixed;

    // ---- phase 2 (mppi_philox_sum.cuh) -----------------------------------------------------------
    philox_weighted_sum_and_finish<A, GenFast<R>>(p, ctrl, costs, w_lo, kfirst, kend, beta_c, max_c, nil, phA, sm, vtid);
}

// -------------------------------------------------------------------------------------------------
// Short rows: the tile's normals stay in shared memory
//
// Tile of warp w: 32 rows (lane = sample) of `rs` float4 (one per Philox call; rs odd, so the rows of a quarter warp
// fall into distinct 16-byte bank groups and the STS.128 of phase 1 are conflict-free).  Weighted sum: the lanes form
// 32 / CW row groups of CW lanes (CW = 8, 16 or 32 >= calls per row); lane (g, c) walks rows g, g + G, .. and keeps
// the float4 of call c (+32 for NC = 2) in registers across all the tiles of the warp, rescaled when the warp's running
// minimum drops (online max-shifted softmin); 8 consecutive lanes read 128 contiguous bytes: conflict-free LDS.128.
// -------------------------------------------------------------------------------------------------
struct ResidentLaunch {
    int rs;      // float4 per tile row
    int cw;      // lanes per row group in the weighted sum (8, 16, 32)
};

constexpr int kResScale = 64;     // merge weights: with the two-level merge no single merge takes more than 64 records

template <int A, int R, int NC>
__global__ void __launch_bounds__(1024, 1)
rollout_philox_resident_kernel(const __grid_constant__ RolloutParams p, const ResidentLaunch G)
{
    constexpr int RS = (A + 3) & ~3;
    extern __shared__ float4 smem_f4[];
    const int NW = blockDim.x >> 5;
    const int TA = p.TA, TAp = (TA + 31) & ~31, ncall = (TA + 3) >> 2, rs = G.rs;
    const int tile_f4 = 32 * rs;
    float4 *sTile = smem_f4;                                         // [NW][32][rs]
    float *sL = reinterpret_cast<float *>(sTile + (size_t)NW * tile_f4);   // [T][RS]
    float *sN = sL + p.T * RS;                                       // [TAp]
    float *sWork = sN + TAp;                                         // [TAp]
    float *sScale = sWork + TAp;                                     // [kResScale]
    float *sRed = sScale + kResScale;                                // [64]
    uint4 *sTab = reinterpret_cast<uint4 *>(sRed + 64);              // [ncall]

    const int ctrl = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    trace_stamp(p, ctrl, 0);
    for (int c = tid; c < ncall; c += blockDim.x) sTab[c] = philox_call_table((uint32_t)c, (uint32_t)ctrl, p);
    const uint32_t phA = philox_uniform_A(p);
    {
        float C = build_linear_tables<A, true>(p, ctrl, sL, reinterpret_cast<float *>(sTile), sRed);   // scratch: the tiles, not yet in use
        C += stage_c0<A>(p, ctrl, sWork, sRed);
        if (blockIdx.x == 0 && tid == 0) p.cost_base[ctrl] = C;      // costs[] holds S_k - C
    }
    FastConsts<A> fc;
    fc.init(p);
    __syncthreads();
    trace_stamp(p, ctrl, 1);

    float *costs = p.costs + (size_t)ctrl * p.K_local;
    const float nil = p.neg_inv_lambda_log2e;
    const int n_tiles = (p.K_local + 31) >> 5;
    float4 *tile = sTile + (size_t)warp * tile_f4;
    float4 *myrow = tile + lane * rs;
    const int cw = G.cw, ng = 32 / cw;                               // lanes per row group, row groups
    const int g = lane / cw, c = lane - g * cw;
    const unsigned gmask = (ng == 32) ? 0xffffffffu : ((1u << ng) - 1u);
    float beta_w = kInf, eta_l = 0.f;
    float4 acc[NC];
#pragma unroll
    for (int i = 0; i < NC; i++) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int t = blockIdx.x * NW + warp; t < n_tiles; t += gridDim.x * NW) {
        const int k = 32 * t + lane;
        const bool valid = k < p.K_local;
        // ---- rollout + cost; the normals of the row are left in the tile ------------------------------
        const PhiloxSample ps = philox_sample(phA, (uint32_t)(p.k_offset + k));
        MPPI_CHECK(t >= 0 && t < n_tiles && ncall <= rs && (size_t)(warp + 1) * tile_f4 <= (size_t)NW * tile_f4);
        float S = fast_rollout<A, R, true>(p, sTab, sL, ps, fc, myrow);
        if (valid) costs[k] = S; else S = kInf;
        // ---- online max-shifted weights -----------------------------------------------------------------
        const float m = warp_min(S);
        if (m < beta_w) {                                            // warp-uniform
            const float f = (beta_w == kInf) ? 0.f : weight_exp(beta_w, m, nil);
#pragma unroll
            for (int i = 0; i < NC; i++) { acc[i].x *= f; acc[i].y *= f; acc[i].z *= f; acc[i].w *= f; }
            eta_l *= f;
            beta_w = m;
        }
        const float e = valid ? sample_weight(S, beta_w, nil) : 0.f;
        eta_l += e;
        const unsigned nz = __ballot_sync(0xffffffffu, e != 0.f);
        __syncwarp();                                                // the tile is complete
        // ---- weighted sum over the resident tile; zero-weight rows add exactly nothing and are skipped ----
        const float4 *col = tile + g * rs + c;                      // row i*ng + g, call c: col[i * ng * rs]
        const int rstep = ng * rs;
        if (nz == 0xffffffffu) {                                     // every row carries weight: no per-row test
#pragma unroll 4
            for (int i = 0; i < cw; i++) {
                const float w = __shfl_sync(0xffffffffu, e, i * ng + g);
                const float2 w2 = make_float2(w, w);
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    if (c + 32 * j < ncall) {
                        MPPI_CHECK((i * ng + g) < 32 && g * rs + c + i * rstep + 32 * j < tile_f4);
                        const float4 v = col[i * rstep + 32 * j];
                        const float2 lo = __ffma2_rn(w2, make_float2(v.x, v.y), make_float2(acc[j].x, acc[j].y));
                        const float2 hi = __ffma2_rn(w2, make_float2(v.z, v.w), make_float2(acc[j].z, acc[j].w));
                        acc[j] = make_float4(lo.x, lo.y, hi.x, hi.y);
                    }
                }
            }
        } else {
            for (int i = 0; i < cw; i++) {                           // rows i*ng .. i*ng + ng - 1, one per row group
                if (((nz >> (i * ng)) & gmask) == 0u) continue;      // warp-uniform
                const float w = __shfl_sync(0xffffffffu, e, i * ng + g);
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    if (c + 32 * j < ncall) {
                        const float4 v = col[i * rstep + 32 * j];
                        acc[j].x = fmaf(w, v.x, acc[j].x); acc[j].y = fmaf(w, v.y, acc[j].y);
                        acc[j].z = fmaf(w, v.z, acc[j].z); acc[j].w = fmaf(w, v.w, acc[j].w);
                    }
                }
            }
        }
        __syncwarp();                                                // everyone is done with the tile before it is overwritten
    }
    // row groups -> one sum per call in lanes 0 .. cw-1 (fixed shuffle tree)
    for (int o = cw; o < 32; o <<= 1) {
#pragma unroll
        for (int j = 0; j < NC; j++) {
            acc[j].x += __shfl_xor_sync(0xffffffffu, acc[j].x, o); acc[j].y += __shfl_xor_sync(0xffffffffu, acc[j].y, o);
            acc[j].z += __shfl_xor_sync(0xffffffffu, acc[j].z, o); acc[j].w += __shfl_xor_sync(0xffffffffu, acc[j].w, o);
        }
    }
    const float eta_w = warp_sum(eta_l);
    if (lane == 0) { sRed[warp] = beta_w; sRed[32 + warp] = eta_w; }
    __syncthreads();
    trace_stamp(p, ctrl, 2);
    // ---- CTA merge of the warps' running sums (fixed order) -----------------------------------------
    float beta_c = kInf;
    for (int w = 0; w < NW; w++) beta_c = fminf(beta_c, sRed[w]);
    float eta_c = 0.f;
    for (int w = 0; w < NW; w++) {
        const float bw = sRed[w];
        if (bw != kInf) eta_c = fmaf(weight_exp(bw, beta_c, nil), sRed[32 + w], eta_c);
    }
    {
        const float sc = (beta_w == kInf) ? 0.f : weight_exp(beta_w, beta_c, nil);
        float4 *out = tile;                                          // the warp's own tile: first ncall float4 = its scaled sums
        if (g == 0) {
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const int cc = c + 32 * j;
                if (cc < ncall) out[cc] = make_float4(sc * acc[j].x, sc * acc[j].y, sc * acc[j].z, sc * acc[j].w);
            }
        }
    }
    __syncthreads();
    for (int j = tid; j < TA; j += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < NW; w++) s += reinterpret_cast<const float *>(sTile + (size_t)w * tile_f4)[j];
        sN[j] = s;
    }
    __syncthreads();
    // the tiles are free now: merge scratch
    publish_and_finish<A, true>(p, ctrl, beta_c, eta_c, sN, sWork, sScale, sRed, sTile, NW * tile_f4);
}

// -------------------------------------------------------------------------------------------------
// Host launchers
// -------------------------------------------------------------------------------------------------
template <int A, int R>
static cudaError_t launch_fast_big(const RolloutParams &p, dim3 grid, size_t smem, cudaStream_t st)
{
    cudaError_t err = ensure_dyn_smem<rollout_philox_fast_kernel<A, R>>(smem);
    if (err != cudaSuccess) return err;
    rollout_philox_fast_kernel<A, R><<<grid, kPhiloxThreads, smem, st>>>(p);
    return cudaGetLastError();
}

template <int A, int R, int NC>
static cudaError_t launch_resident_V(const RolloutParams &p, ResidentLaunch G, dim3 grid, int threads, size_t smem, int cluster_x,
                                     cudaStream_t st)
{
    cudaError_t err = ensure_dyn_smem<rollout_philox_resident_kernel<A, R, NC>>(smem);
    if (err != cudaSuccess) return err;
    if (cluster_x <= 1) {
        rollout_philox_resident_kernel<A, R, NC><<<grid, threads, smem, st>>>(p, G);
        return cudaGetLastError();
    }
    // thread-block clusters along x: the CTAs of a cluster are reduced over distributed shared memory (publish_and_finish)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster_x;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, rollout_philox_resident_kernel<A, R, NC>, p, G);
}

static size_t resident_smem_bytes(int A, int T, int TA, int nw, int rs)
{
    const int RS = (A + 3) & ~3, TAp = (TA + 31) & ~31;
    return (size_t)nw * 32 * rs * sizeof(float4) + sizeof(float) * ((size_t)T * RS + 2 * TAp + kResScale + 64) +
           sizeof(uint4) * (size_t)((TA + 3) >> 2);
}

// Geometry of the resident kernel, or false when the rows are too long for it (fewer than 12 resident warps per SM).
bool resident_geometry(int A, int T, int TA, int K_local, int n_ctrl, int num_sms, size_t smem_sm, size_t smem_cta_limit, int *nw_out,
                       int *rs_out, int *cw_out, int *grid_x_out, size_t *smem_out, bool *latency_out)
{
    const int ncall = (TA + 3) >> 2;
    if (ncall > 64) return false;
    const int rs = ncall | 1;
    const int n_tiles = (K_local + 31) / 32;
    int best_nw = 0, best_warps = 0;
    for (int nw = 16; nw >= 4; nw >>= 1) {
        const size_t b = resident_smem_bytes(A, T, TA, nw, rs) + 1024;     // + the per-CTA reservation
        if (b > smem_cta_limit) continue;
        int ctas = (int)(smem_sm / b);
        if (ctas > 32 / nw) ctas = 32 / nw;                                 // 64 registers per thread: 32 warps per SM
        const int warps = ctas * nw;
        if (warps > best_warps) { best_warps = warps; best_nw = nw; }
    }
    if (best_warps < 12) return false;
    int nw = best_nw, gx;
    const long long slots = (long long)num_sms * best_warps, tiles_total = (long long)n_ctrl * n_tiles;
    if (tiles_total >= slots) {
        // saturating: as many resident warps as fit; enough CTAs per controller to fill them
        gx = (int)((slots + (long long)n_ctrl * nw - 1) / ((long long)n_ctrl * nw));
        const int cap = (n_tiles + nw - 1) / nw;
        if (gx > cap) gx = cap;
    } else {
        // latency-bound: a warp per tile, the warps spread over as many SMs as a cluster reaches (eight CTAs reduced over
        // distributed shared memory: up to 64 tiles per controller need no partial record and no last-CTA election at all);
        // larger controllers take 16-warp CTAs in clusters of eight
        if (n_tiles <= 64) {
            int cs = 1;
            while (cs < 8 && cs * 2 <= n_tiles) cs <<= 1;
            const int per = (n_tiles + cs - 1) / cs;
            nw = 1;
            while (nw < per) nw <<= 1;
            gx = cs;
        } else {
            nw = 8;                     // two or more CTAs per SM: a cluster of eight finds room in any GPC
            while (nw > 4 && 2 * (resident_smem_bytes(A, T, TA, nw, rs) + 1024) > smem_sm) nw >>= 1;
            gx = (n_tiles + nw - 1) / nw;
            gx = (gx + 7) & ~7;
        }
    }
    if (gx < 1) gx = 1;
    if (gx > kMaxParts) gx = kMaxParts;
    *nw_out = nw;
    *rs_out = rs;
    *cw_out = ncall <= 8 ? 8 : (ncall <= 16 ? 16 : 32);
    *grid_x_out = gx;
    *smem_out = resident_smem_bytes(A, T, TA, nw, rs);
    *latency_out = tiles_total < slots;
    return true;
}

template <int A>
static cudaError_t launch_fast_A(const RolloutParams &p, int variant, int num_sms, size_t smem_sm, size_t smem_limit, cudaStream_t st,
                                 int *grid_x_out)
{
    int nw = 0, rs = 0, cw = 0, gx = 0;
    size_t smem = 0;
    bool latency = false;
    const bool can_res = p.norm_mode == 0 && resident_geometry(A, p.T, p.TA, p.K_local, p.n_ctrl, num_sms, smem_sm, smem_limit, &nw, &rs,
                                                               &cw, &gx, &smem, &latency);
    if (variant == 2 && !can_res) return cudaErrorInvalidConfiguration;
    if (can_res && variant != 1) {
        if (p.max_parts > 0 && gx > p.max_parts) gx = p.max_parts;
        if (grid_x_out) *grid_x_out = gx;
        // clusters only where the grid is far from filling the GPU (a full grid of 8-CTA clusters does not tile the GPCs)
        int cs = 1;
        if (latency && !getenv("MPPI_NO_CLUSTER"))
            while (cs < 8 && gx % (cs * 2) == 0) cs <<= 1;
        const ResidentLaunch G{rs, cw};
        const dim3 grid(gx, p.n_ctrl);
        const bool nc2 = ((p.TA + 3) >> 2) > 32;
        if (p.rounds == 7)
            return nc2 ? launch_resident_V<A, 7, 2>(p, G, grid, 32 * nw, smem, cs, st) : launch_resident_V<A, 7, 1>(p, G, grid, 32 * nw, smem, cs, st);
        return nc2 ? launch_resident_V<A, 10, 2>(p, G, grid, 32 * nw, smem, cs, st) : launch_resident_V<A, 10, 1>(p, G, grid, 32 * nw, smem, cs, st);
    }
    const size_t smem_big = philox_smem_bytes(A, p.T, p.TA);
    if (smem_big > smem_limit) return cudaErrorInvalidConfiguration;
    gx = philox_grid_x(p.K_local, p.n_ctrl, num_sms);
    if (p.max_parts > 0 && gx > p.max_parts) gx = p.max_parts;
    if (grid_x_out) *grid_x_out = gx;
    const dim3 grid(gx, p.n_ctrl);
    return p.rounds == 7 ? launch_fast_big<A, 7>(p, grid, smem_big, st) : launch_fast_big<A, 10>(p, grid, smem_big, st);
}

// variant: 0 = pick (resident when the rows are short enough), 1 = regenerating kernel, 2 = resident kernel
cudaError_t launch_rollout_philox_fast(RolloutParams p, int a, int variant, int num_sms, size_t smem_sm, size_t smem_limit,
                                       cudaStream_t st, int *grid_x_out)
{
    MPPI_DISPATCH_A(a, return launch_fast_A<A_>(p, variant, num_sms, smem_sm, smem_limit, st, grid_x_out));
    return cudaSuccess;
}

}  // namespace mppi
