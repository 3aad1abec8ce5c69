// Fused MPPI update kernels for the point-mass model (sm_100a).
//
//   rollout_philox_kernel    the hot path: noise regenerated in registers from a counter-based
//                            Philox stream, never stored.  Phase 1 rolls every sample of the CTA
//                            through the horizon and keeps only its cost; phase 2 re-walks the same
//                            Philox counters to accumulate sum_k e_k z_k.  One launch per update:
//                            the last CTA to finish merges all CTA partials, applies the update
//                            and shifts the sequence.
//   rollout_injected_kernel  parity/debug mode: eps is read from HBM exactly once.  A 32-sample tile
//                            is landed in shared memory by the TMA unit (cp.async.bulk) and stays
//                            resident between the rollout (lane = sample) and the weighted sum
//                            (lane = column) with an online max-shifted rescale.
//   finish_kernel            multi-rank only: merges the all-gathered rank payloads.
//
// Reference maths: /root/reference/src/controller_base.cpp:166-329, src/model_base.cpp:53-82,
// src/cost_base.cpp:37-68 (restated in the CPU checker that tests compare against).
#include <cstdio>
#include <cstdlib>

#include "mppi_device.cuh"
#include "mppi_internal.h"
#include "mppi_update.cuh"
#include "mppi_philox_sum.cuh"
#include "mppi_linear.cuh"

namespace mppi {

template <int A>
__device__ __forceinline__ void load_uv(const float *uv_row, Vec<A> &U, Vec<A> &w)
{
    constexpr int RS = Row<A>::RS, H = Row<A>::H, NP = A / 2;
    float uv[RS];
#pragma unroll
    for (int i = 0; i < RS / 4; i++) {
        const float4 v = reinterpret_cast<const float4 *>(uv_row)[i];
        uv[4 * i] = v.x; uv[4 * i + 1] = v.y; uv[4 * i + 2] = v.z; uv[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < NP; i++) {
        U.pr[i] = make_float2(uv[2 * i], uv[2 * i + 1]);
        w.pr[i] = make_float2(uv[H + 2 * i], uv[H + 2 * i + 1]);
    }
    U.sc = uv[A - 1];
    w.sc = uv[H + A - 1];
}

// Philox mode: the action-cost vector is w_t = w_scale * U_t with a uniform scale, so only U_t is staged (one LDS.128 per
// step at a <= 4) and the dot products U_t . z_t are summed on their own and scaled once per sample.
template <int A>
struct RowU { static constexpr int RS = (A + 3) & ~3; };

template <int A>
__device__ __forceinline__ void load_u(const float *u_row, Vec<A> &U)
{
    constexpr int RS = RowU<A>::RS, NP = A / 2;
    float uv[RS];
#pragma unroll
    for (int i = 0; i < RS / 4; i++) {
        const float4 v = reinterpret_cast<const float4 *>(u_row)[i];
        uv[4 * i] = v.x; uv[4 * i + 1] = v.y; uv[4 * i + 2] = v.z; uv[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < NP; i++) U.pr[i] = make_float2(uv[2 * i], uv[2 * i + 1]);
    U.sc = uv[A - 1];
}

template <int A>
__device__ __forceinline__ void stage_u(const RolloutParams &p, int ctrl, float *sU)
{
    constexpr int RS = RowU<A>::RS;
    const float *U = p.U + (size_t)ctrl * p.TA;
    for (int i = threadIdx.x; i < p.T * RS; i += blockDim.x) {
        const int t = i / RS, j = i - t * RS;
        sU[i] = (j < A) ? U[t * A + j] : 0.f;
    }
}

template <int A>
__device__ __forceinline__ void vec_from(const float *z, Vec<A> &n)
{
#pragma unroll
    for (int i = 0; i < A / 2; i++) n.pr[i] = make_float2(z[2 * i], z[2 * i + 1]);
    n.sc = z[A - 1];
}

// One rollout step for one sample (src/controller_base.cpp:251-269):
//   u = U_t + eps_t ; x <- A x + (B/m) u ; S += q(x) + lambda U_t^T Sigma^-1 eps_t
// `n` is z_t in Philox mode (eps_t = Sigma z_t formed here) and eps_t in injected mode.
template <int A, bool PHILOX, bool DIAG, bool QUAD, int COST>
__device__ __forceinline__ void rollout_step(PointMass<A> &x, CostAcc &S, CostAcc &Sw, const float *uv_row, const Vec<A> &n,
                                             const RolloutParams &p, const ModelConsts<A> &mc, const Vec<A> &sigd)
{
    constexpr int NP = A / 2;
    constexpr bool ODD = (A & 1) != 0;
    Vec<A> U, w, u;
    if (PHILOX) load_u<A>(uv_row, U);            // Sw collects U_t . z_t, scaled by w_scale once per sample
    else load_uv<A>(uv_row, U, w);
    if (PHILOX && !DIAG) {
#pragma unroll
        for (int j = 0; j < A; j++) {
            float e = 0.f;
#pragma unroll
            for (int l = 0; l < A; l++) e = fmaf(p.sigma[j * A + l], n.get(l), e);
            u.set(j, U.get(j) + e);
        }
    } else {
#pragma unroll
        for (int i = 0; i < NP; i++)
            u.pr[i] = PHILOX ? __ffma2_rn(sigd.pr[i], n.pr[i], U.pr[i]) : __fadd2_rn(U.pr[i], n.pr[i]);
        if (ODD) u.sc = PHILOX ? fmaf(sigd.sc, n.sc, U.sc) : U.sc + n.sc;
    }
    if (PHILOX) {
#pragma unroll
        for (int i = 0; i < NP; i++) Sw.a2 = __ffma2_rn(U.pr[i], n.pr[i], Sw.a2);
        if (ODD) Sw.a = fmaf(U.sc, n.sc, Sw.a);
    } else {
#pragma unroll
        for (int i = 0; i < NP; i++) S.a2 = __ffma2_rn(w.pr[i], n.pr[i], S.a2);
        if (ODD) S.a = fmaf(w.sc, n.sc, S.a);
    }
    if (QUAD) {                                  // Python-twin noise cost (cost_base.py:147-148,160-162)
        float nn[A];
#pragma unroll
        for (int j = 0; j < A; j++) nn[j] = n.get(j);
        S.a += quad_cost<A>(p, nn);
    }
    x.step(u, mc);
    add_state_cost<A, COST>(x, mc, p, S);
}

template <int A>
__device__ __forceinline__ void thread_consts(const RolloutParams &p, int ctrl, ModelConsts<A> &mc, Vec<A> &sigd,
                                              float (&x0)[2 * A])
{
    const float *gp = p.goal + (p.goal_per_ctrl ? (size_t)ctrl * 2 * A : 0);
    const float *xp = p.x + (size_t)ctrl * 2 * A;
    float g[2 * A];
#pragma unroll
    for (int i = 0; i < 2 * A; i++) {
        g[i] = gp[i];
        x0[i] = p.x_inline ? p.x0[i] : xp[i];
    }
    mc.init(p.dt, p.c_pu, p.c_vu, p.sqrt_q, g);
#pragma unroll
    for (int j = 0; j < A; j++) sigd.set(j, p.sigma[j * A + j]);
}

// -------------------------------------------------------------------------------------------------
// Philox mode
// -------------------------------------------------------------------------------------------------
template <int A, bool DIAG, bool QUAD, int COST>
__global__ void __launch_bounds__(kPhiloxThreads, kPhiloxCtasPerSm)
rollout_philox_kernel(const __grid_constant__ RolloutParams p)
{
    constexpr int RS = RowU<A>::RS;
    constexpr int NW = kPhiloxThreads / 32;
    extern __shared__ float4 smem_f4[];
    const int TA = p.TA, TAp = (TA + 31) & ~31;
    PhiloxSmem sm;
    sm.carve(reinterpret_cast<float *>(smem_f4), p.T, RS, TAp);
    float *sUV = sm.sUV, *sWork = sm.sWork, *sRed = sm.sRed;
    uint4 *sTab = sm.sTab;

    const int ctrl = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    stage_u<A>(p, ctrl, sUV);
    for (int c = tid; c < ((TA + 3) >> 2); c += kPhiloxThreads) sTab[c] = philox_call_table((uint32_t)c, (uint32_t)ctrl, p);
    const uint32_t phA = philox_uniform_A(p);

    ModelConsts<A> mc;
    Vec<A> sigd;
    float x0[2 * A];
    thread_consts<A>(p, ctrl, mc, sigd, x0);
    __syncthreads();
    const float C0 = stage_c0<A>(p, ctrl, sWork, sRed);

    float *costs = p.costs + (size_t)ctrl * p.K_local;
    // contiguous, evenly sized sample range per CTA (32-sample granularity): no whole-iteration
    // quantisation when K_local is not a multiple of gridDim.x * blockDim.x
    const int n_w = (p.K_local + 31) >> 5;
    const int w_lo = (int)((long long)n_w * blockIdx.x / gridDim.x);
    const int w_hi = (int)((long long)n_w * (blockIdx.x + 1) / gridDim.x);
    const int kfirst = 32 * w_lo + tid;
    const int kend = min(p.K_local, 32 * w_hi);
    constexpr int kstride = kPhiloxThreads;
    const int nfull = p.T >> 2, trem = p.T & 3;

    // ---- phase 1: rollout + cost ---------------------------------------------------------------
    float bmin = kInf, bmax = -kInf;
    for (int k = kfirst; k < kend && p.norm_mode != 2; k += kstride) {   // weight pass of a normalised update: costs are in HBM
        const PhiloxSample ps = philox_sample(phA, (uint32_t)(p.k_offset + k));
        PointMass<A> x;
        x.init(x0);
        CostAcc S, Sw;
        KahanSum Sk_;                               // block sums added with compensation (mppi_device.cuh)
        Sk_.init(C0);
        Sw.zero();
        const float *uv = sUV;
        uint32_t call = 0;
        for (int tb = 0; tb < nfull; tb++) {       // full blocks of 4 steps = A Philox calls, no guards
            float z[4 * A];
#pragma unroll
            for (int c = 0; c < A; c++) normals4_tab(sTab, call + c, ps, p, &z[4 * c]);
            call += A;
            S.zero();
#pragma unroll
            for (int tt = 0; tt < 4; tt++) {
                Vec<A> n;
                vec_from<A>(&z[tt * A], n);
                rollout_step<A, true, DIAG, QUAD, COST>(x, S, Sw, uv + tt * RS, n, p, mc, sigd);
            }
            Sk_.add(S.total());
            uv += 4 * RS;
        }
        S.zero();
        if (trem) {                                 // tail: T % 4 steps (calls past the row end read table entries that exist: ceil)
            float z[4 * A];
#pragma unroll
            for (int c = 0; c < A; c++)
                if ((int)(call + c) < ((TA + 3) >> 2)) normals4_tab(sTab, call + c, ps, p, &z[4 * c]);
#pragma unroll
            for (int tt = 0; tt < 3; tt++)
                if (tt < trem) {
                    Vec<A> n;
                    vec_from<A>(&z[tt * A], n);
                    rollout_step<A, true, DIAG, QUAD, COST>(x, S, Sw, uv + tt * RS, n, p, mc, sigd);
                }
        }
        add_state_cost<A, COST>(x, mc, p, S);    // terminal cost on top of step T-1's (src/controller_base.cpp:271-272)
        Sk_.add(fmaf(p.w_scale, Sw.total(), S.total()));
        const float Sk = Sk_.s;
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
    if (p.norm_mode == 1) {                 // cost pass of a normalised update: publish (min, max) and stop
        publish_minmax(p, ctrl, beta_c, max_c, sRed);
        return;
    }
    float beta_fixed = 0.f;
    const float nil = weight_scale(p, ctrl, beta_fixed);
    if (p.norm_mode == 2) beta_c = beta_fixed;

    // ---- phase 2: sum_k e_k z_k, regenerating z from the same counters (mppi_philox_sum.cuh) ----------
    philox_weighted_sum_and_finish<A, GenPlain>(p, ctrl, costs, w_lo, kfirst, kend, beta_c, max_c, nil, phA, sm);
}

// -------------------------------------------------------------------------------------------------
// Injected-noise mode
//
// A tile is 32 samples x (T*a) floats, contiguous in HBM, landed in shared memory by ONE TMA bulk copy
// and kept resident between the rollout and the weighted sum, so eps is read from HBM exactly once.
// Large rows (config 3: 38.4 KB per tile) leave room for only a few tiles per SM, so a tile is
// processed by a GROUP of C warps that split the horizon (lane = sample in every warp):
//   pass 1  the model is linear, so a chunk of n steps maps x_in -> M^n x_in + b with
//           M^n = [[1, n dt],[0, 1]] per axis and b the zero-state response.  With C = sum_t u_t and
//           D = sum_t (exclusive prefix of u), b = (c_pu C + dt c_vu D, c_vu C); u = U_t + eps_t splits
//           into a per-CTA part (from U, computed once) and a per-sample part (two adds per axis-step);
//   pass 2  every warp rebuilds its true incoming state from x0 and b_0..b_{c-1}, rolls its chunk
//           with the costs, and the chunk costs are summed in a fixed order;
//   sum     the weighted sum over the resident tile is split by column blocks across the C warps.
// With C = 1 this degenerates to one warp per tile and no barriers.
// -------------------------------------------------------------------------------------------------
struct InjectedLaunch {
    int ng;      // tile groups per CTA
    int c;       // warps per group (time chunks)
    int nbuf;    // tile buffers of the CTA: a shared pool, tile sequence number seq lives in buffer seq % nbuf
};

__device__ __forceinline__ void group_barrier(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Noise of the 4 steps of block tb (4A floats) from the lane's resident row.
template <int A, bool TMA, bool GUARD>
__device__ __forceinline__ void load_block(const float *row, int tb, int TA, float (&e)[4 * A])
{
    if (TMA) {   // T*a % 4 == 0: rows are 16-B aligned; LDS.128 is conflict-free when T*a/4 is odd
#pragma unroll
        for (int c = 0; c < A; c++) {
            const int i4 = tb * A + c;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!GUARD || 4 * i4 < TA) v = reinterpret_cast<const float4 *>(row)[i4];
            e[4 * c] = v.x; e[4 * c + 1] = v.y; e[4 * c + 2] = v.z; e[4 * c + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4 * A; i++) {
            const int idx = tb * 4 * A + i;
            e[i] = (!GUARD || idx < TA) ? row[idx] : 0.f;
        }
    }
}

// FAST: the superposition form of mppi_linear.cuh (q > 0, StaticCost, no noise-quadratic term; any Sigma, the noise arrives
// scaled): per-controller tables instead of the staged (U_t, w_t) rows, six FMAs per axis-step, costs[] relative to C.
template <int A, bool TMA, bool QUAD, int COST, bool FAST>
__global__ void __launch_bounds__(512, 1)
rollout_injected_kernel(const __grid_constant__ RolloutParams p, const InjectedLaunch L)
{
    constexpr int RS = FAST ? ((A + 3) & ~3) : Row<A>::RS;
    extern __shared__ float4 smem_f4[];
    float *smem = reinterpret_cast<float *>(smem_f4);
    const int TA = p.TA, TAp = (TA + 31) & ~31;
    const int NG = L.ng, C = L.c, NBUF = L.nbuf;
    const int tile_words = 32 * TA;                // multiple of 32 words: every tile 128-B aligned
    float *sTiles = smem;                          // [NBUF][32][TA]
    float *sUV = sTiles + (size_t)NBUF * tile_words;   // [T][RS]
    float *sAccG = sUV + p.T * RS;                 // [NG][TAp] running per-group sums
    float *sN = sAccG + NG * TAp;                  // [TAp]
    float *sWork = sN + TAp;                       // [TAp]
    float *sScale = sWork + TAp;                   // [kMaxParts]
    float *sRed = sScale + kMaxParts;              // [64]
    float *sB = sRed + 64;                         // [NG][C][2A][32] zero-state chunk responses
    float *sS = sB + NG * C * 2 * A * 32;          // [NG][C][32] chunk costs
    uint64_t *bars = reinterpret_cast<uint64_t *>(sS + NG * C * 32);   // [NBUF]
    volatile int *sIssued = reinterpret_cast<volatile int *>(bars + 64);   // [NBUF] tile sequence number last issued into a buffer

    const int ctrl = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp / C, cw = warp - grp * C;
    const int n_tiles = (p.K_local + 31) >> 5;
    const int nseq = (n_tiles > (int)blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const float *eps = p.eps + (size_t)ctrl * p.K_local * TA;

    // Tile sequence number seq (0, 1, .. of this CTA) is consumed by group seq % NG and lives in buffer
    // seq % NBUF, completing phase seq / NBUF of that buffer's mbarrier.  The group that finishes tile seq
    // refills its buffer with tile seq + NBUF at once - a tile another group will consume - so with
    // NBUF > NG every group finds its next tile already in flight (config 3: 4 groups, 5 buffers).
    auto issue = [&](int seq) {   // one elected lane; TMA bulk copy of a whole tile
        const int b = seq % NBUF;
        const int gt = seq * gridDim.x + blockIdx.x;
        const int rows = min(32, p.K_local - 32 * gt);
        const uint32_t bytes = (uint32_t)rows * TA * 4u;
        sIssued[b] = seq;
        mbar_expect_tx(&bars[b], bytes);
        bulk_g2s(sTiles + (size_t)b * tile_words, eps + (size_t)gt * tile_words, bytes, &bars[b]);
    };

    if (TMA) {
        if (tid == 0) {
            for (int b = 0; b < NBUF; b++) { mbar_init(&bars[b], 1); sIssued[b] = -1; }
            fence_mbar_init();
        }
    }
    if (!FAST) stage_sequence<A, false>(p, ctrl, sUV);
    for (int i = tid; i < NG * TAp; i += blockDim.x) sAccG[i] = 0.f;
    __syncthreads();
    float C0 = 0.f;
    if (FAST) {
        // tables first (their scratch is the tile pool), then the first loads
        if (p.norm_mode != 2) {
            float Cb = build_linear_tables<A, false>(p, ctrl, sUV, sTiles, sRed);
            Cb += stage_c0<A>(p, ctrl, sWork, sRed);
            if (blockIdx.x == 0 && tid == 0) p.cost_base[ctrl] = Cb;
        }
        __syncthreads();
    }
    if (TMA && tid == 0) {
        if (FAST) fence_proxy_async();           // generic-proxy writes to the pool (table scratch) before the bulk copies land
        for (int seq = 0; seq < NBUF && seq < nseq; seq++) issue(seq);
    }

    ModelConsts<A> mc;
    Vec<A> sigd;
    float x0[2 * A];
    thread_consts<A>(p, ctrl, mc, sigd, x0);
    FastConsts<A> fc;
    if (FAST) fc.init(p);
    if (!FAST) C0 = stage_c0<A>(p, ctrl, sWork, sRed);
    float beta_fixed = 0.f;
    const float nil = weight_scale(p, ctrl, beta_fixed);

    float *costs = p.costs + (size_t)ctrl * p.K_local;
    float *accg = sAccG + grp * TAp;
    // time chunk of this warp, in blocks of 4 steps
    const int nblk = (p.T + 3) >> 2;
    const int nb = (nblk + C - 1) / C;
    const int tb0 = min(nblk, cw * nb), tb1 = min(nblk, tb0 + nb);
    const int bar_id = 1 + grp, bar_n = C * 32;
    // column blocks (32 columns each) of this warp in the weighted sum
    const int NB = (TA + 31) >> 5;
    const int nbw = (NB + C - 1) / C;
    const int cb0 = min(NB, cw * nbw), cb1 = min(NB, cb0 + nbw);
    float beta_g = (p.norm_mode == 2) ? beta_fixed : kInf, eta_lane = 0.f;
    float gmin = kInf, gmax = -kInf;               // cost pass of a normalised update

    // U-part of this chunk's zero-state sums (same for every sample): CU = sum U_t, DU = sum prefix
    Vec<A> CU, DU;
    CU.fill(0.f);
    DU.fill(0.f);
    if (C > 1 && !FAST) {
        const int s0 = 4 * tb0, s1 = min(p.T, 4 * tb1);
        for (int t = s0; t < s1; t++) {
#pragma unroll
            for (int j = 0; j < A; j++) {
                DU.set(j, DU.get(j) + CU.get(j));
                CU.set(j, CU.get(j) + sUV[t * RS + j]);
            }
        }
    }

    if (grp < NG) {
        for (int seq = grp; seq < nseq; seq += NG) {
            const int b = TMA ? seq % NBUF : grp;          // cooperative-load fallback: a private buffer per group
            const int gt = seq * gridDim.x + blockIdx.x;
            const int rows = min(32, p.K_local - 32 * gt);
            MPPI_CHECK(b >= 0 && b < NBUF && gt >= 0 && gt < n_tiles && rows >= 1 && rows <= 32);
            float *tile = sTiles + (size_t)b * tile_words;
            if (TMA) {
                // A parity wait cannot tell phase n from phase n - 2: with a shared pool a group may get here before
                // the load of ITS tile has even been issued (the buffer's previous user, another group, is still
                // busy), so first wait until the buffer has been handed to this tile, then for the bytes.
                while (sIssued[b] != seq) __nanosleep(32);
                mbar_wait(&bars[b], (uint32_t)((seq / NBUF) & 1));
                if (rows < 32 && cw == 0) {   // zero the rows the copy did not write (e = 0 must not meet NaN garbage)
                    for (int i = rows * TA + lane; i < tile_words; i += 32) tile[i] = 0.f;
                }
            } else {           // generic fallback (T*a not a multiple of 4): coalesced loads by the group
                const float *src = eps + (size_t)gt * tile_words;
                const int nvalid = rows * TA;
                for (int i = cw * 32 + lane; i < tile_words; i += C * 32) tile[i] = (i < nvalid) ? __ldg(src + i) : 0.f;
                if (C > 1) group_barrier(bar_id, bar_n);
            }
            __syncwarp();
            const float *row = tile + lane * TA;

            float S;
            if (p.norm_mode == 2) {                // weight pass of a normalised update: costs are in HBM
                S = (lane < rows) ? costs[32 * gt + lane] : 0.f;
                if (TMA && C > 1) group_barrier(bar_id, bar_n);   // the zeroed tail rows are visible to the whole group
            } else {
                PointMass<A> x;
                if (C > 1) {
                    // ---- pass 1: per-sample part of the chunk's zero-state sums ------------------------
                    Vec<A> Cs, Ds;
                    Cs.fill(0.f);
                    Ds.fill(0.f);
                    for (int tb = tb0; tb < tb1; tb++) {
                        float e[4 * A];
                        if (4 * tb + 4 <= p.T) load_block<A, TMA, false>(row, tb, TA, e);
                        else load_block<A, TMA, true>(row, tb, TA, e);    // zeros past the row end
    #pragma unroll
                        for (int tt = 0; tt < 4; tt++) {
                            Vec<A> n;
                            vec_from<A>(&e[tt * A], n);
    #pragma unroll
                            for (int i = 0; i < A / 2; i++) {
                                Ds.pr[i] = __fadd2_rn(Ds.pr[i], Cs.pr[i]);
                                Cs.pr[i] = __fadd2_rn(Cs.pr[i], n.pr[i]);
                            }
                            if (A & 1) { Ds.sc += Cs.sc; Cs.sc += n.sc; }
                        }
                    }
                    // the guarded tail appends zeros: (4*tb1 - T) extra prefix terms entered D; remove them
                    const int extra = 4 * tb1 - min(p.T, 4 * tb1);
                    float *myb = sB + ((size_t)(grp * C + cw) * 2 * A) * 32 + lane;
                    const float dtcvu = p.dt * p.c_vu;
    #pragma unroll
                    for (int i = 0; i < A; i++) {
                        if (FAST) {           // zero-state response of the noise-driven part (P, V): no U part
                            const float Ci = Cs.get(i), Di = Ds.get(i) - (float)extra * Cs.get(i);
                            myb[(2 * i) * 32] = fmaf(p.fb1[i], Ci, (p.fa1[i] * p.fb2[i]) * Di);
                            myb[(2 * i + 1) * 32] = p.fb2[i] * Ci;
                        } else {
                            const float Ci = Cs.get(i) + CU.get(i);
                            const float Di = (Ds.get(i) - (float)extra * Cs.get(i)) + DU.get(i);
                            myb[(2 * i) * 32] = fmaf(p.c_pu, Ci, dtcvu * Di);
                            myb[(2 * i + 1) * 32] = p.c_vu * Ci;
                        }
                    }
                    group_barrier(bar_id, bar_n);
                    // true incoming state: x0 pushed through chunks 0..cw-1 (free response + zero-state response);
                    // FAST: x holds (P, V), which start at zero and whose free response is P += n a1 V
                    if (FAST) x.zero(); else x.init(x0);
                    for (int cc = 0; cc < cw; cc++) {
                        const int s0 = 4 * min(nblk, cc * nb), s1 = min(p.T, 4 * min(nblk, cc * nb + nb));
                        const float ndt = (float)(s1 - s0) * p.dt;
                        const float *ob = sB + ((size_t)(grp * C + cc) * 2 * A) * 32 + lane;
    #pragma unroll
                        for (int i = 0; i < A; i++) {
                            const float vi = x.v.get(i);
                            const float fr = FAST ? (float)(s1 - s0) * p.fa1[i] : ndt;
                            x.p.set(i, fmaf(fr, vi, x.p.get(i)) + ob[(2 * i) * 32]);
                            x.v.set(i, vi + ob[(2 * i + 1) * 32]);
                        }
                    }
                } else {
                    if (FAST) x.zero(); else x.init(x0);
                }

                // ---- pass 2: rollout with costs over this warp's chunk ----------------------------------
                CostAcc Sa, Sl;
                Sa.zero();
                Sl.zero();
                KahanSum Sk_;                       // block sums (4 steps, from zero) added with compensation
                Sk_.init(cw == 0 ? C0 : 0.f);
                for (int tb = tb0; tb < tb1; tb++) {
                    float e[4 * A];
                    Sk_.add(FAST ? Sa.total() + Sl.total() : Sa.total());
                    Sa.zero();
                    Sl.zero();
                    if (4 * tb + 4 <= p.T) {
                        load_block<A, TMA, false>(row, tb, TA, e);
    #pragma unroll
                        for (int tt = 0; tt < 4; tt++) {
                            Vec<A> n;
                            vec_from<A>(&e[tt * A], n);
                            if (FAST) fast_step<A>(x.p, x.v, Sa, Sl, sUV + (4 * tb + tt) * RS, n, fc);
                            else rollout_step<A, false, false, QUAD, COST>(x, Sa, Sa, sUV + (4 * tb + tt) * RS, n, p, mc, sigd);
                        }
                    } else {
                        load_block<A, TMA, true>(row, tb, TA, e);
    #pragma unroll
                        for (int tt = 0; tt < 4; tt++)
                            if (4 * tb + tt < p.T) {
                                Vec<A> n;
                                vec_from<A>(&e[tt * A], n);
                                if (FAST) fast_step<A>(x.p, x.v, Sa, Sl, sUV + (4 * tb + tt) * RS, n, fc);
                                else rollout_step<A, false, false, QUAD, COST>(x, Sa, Sa, sUV + (4 * tb + tt) * RS, n, p, mc, sigd);
                            }
                    }
                }
                if (tb1 == nblk && tb0 < tb1) {                  // terminal cost (src/controller_base.cpp:271-272)
                    if (FAST) fast_terminal<A>(x.p, x.v, Sa);
                    else add_state_cost<A, COST>(x, mc, p, Sa);
                }
                Sk_.add(FAST ? Sa.total() + Sl.total() : Sa.total());
                S = Sk_.s;
                if (C > 1) {
                    sS[(grp * C + cw) * 32 + lane] = S;
                    group_barrier(bar_id, bar_n);
                    S = 0.f;
                    for (int cc = 0; cc < C; cc++) S += sS[(grp * C + cc) * 32 + lane];   // fixed order in every warp
                }
            }
            MPPI_CHECK(lane >= rows || 32 * gt + lane < p.K_local);
            if (p.norm_mode != 2 && cw == 0 && lane < rows) costs[32 * gt + lane] = S;
            if (p.norm_mode == 1) {                // cost pass: track (min, max), no weights yet
                if (lane < rows) { gmin = fminf(gmin, S); gmax = fmaxf(gmax, S); }
                if (C > 1) group_barrier(bar_id, bar_n); else __syncwarp();
                if (TMA && lane == 0 && cw == 0 && seq + NBUF < nseq) {
                    fence_proxy_async();
                    issue(seq + NBUF);
                }
                continue;
            }
            if (lane >= rows) S = kInf;

            // ---- weighted sum: online max-shifted weights; lane = column of the resident tile ---------
            const float m = warp_min(S);
            if (m < beta_g) {
                if (beta_g != kInf) {
                    const float f = weight_exp(beta_g, m, nil);
                    for (int c = cb0 * 32 + lane; c < min(TA, cb1 * 32); c += 32) accg[c] *= f;
                    eta_lane *= f;
                }
                beta_g = m;
            }
            const float ek = (lane < rows) ? sample_weight(S, beta_g, nil) : 0.f;
            eta_lane += ek;
            // rows whose weight is exactly zero (fp32 underflow) add nothing: when they are the majority,
            // visit the others only (bit scan); otherwise the unrolled dense loop is cheaper per row
            const unsigned nzrows = __ballot_sync(0xffffffffu, ek != 0.f);
            const bool sparse = __popc(nzrows) <= 12;
            for (int cb = cb0; cb < cb1 && nzrows != 0u; cb += 4) {
                // 4 column blocks per pass: offsets 32*mm are immediates.  Blocks past cb1 / columns past TA
                // read neighbouring shared memory and are discarded.
                const float *base = tile + cb * 32 + lane;
                float a4[4] = {0.f, 0.f, 0.f, 0.f};
                if (sparse) {
                    for (unsigned m = nzrows; m != 0u; m &= m - 1u) {
                        const int k = __ffs((int)m) - 1;
                        const float w = __shfl_sync(0xffffffffu, ek, k);
                        const float *ptr = base + k * TA;
#pragma unroll
                        for (int mm = 0; mm < 4; mm++) a4[mm] = fmaf(w, ptr[32 * mm], a4[mm]);
                    }
                } else {
                    const float *ptr = base;
#pragma unroll 8
                    for (int k = 0; k < 32; k++) {
                        const float w = __shfl_sync(0xffffffffu, ek, k);
#pragma unroll
                        for (int mm = 0; mm < 4; mm++) a4[mm] = fmaf(w, ptr[32 * mm], a4[mm]);
                        ptr += TA;
                    }
                }
#pragma unroll
                for (int mm = 0; mm < 4; mm++) {
                    const int c = (cb + mm) * 32 + lane;
                    MPPI_CHECK(!(cb + mm < cb1 && c < TA) || c < TAp);
                    if (cb + mm < cb1 && c < TA) accg[c] += a4[mm];
                }
            }
            if (C > 1) group_barrier(bar_id, bar_n); else __syncwarp();   // everyone is done with the tile
            if (TMA && lane == 0 && cw == 0 && seq + NBUF < nseq) {
                fence_proxy_async();
                issue(seq + NBUF);
            }
        }
        eta_lane = warp_sum(eta_lane);
        if (p.norm_mode == 1) {
            gmin = warp_min(gmin);
            gmax = -warp_min(-gmax);
            if (lane == 0 && cw == 0) { sRed[grp] = gmin; sRed[32 + grp] = gmax; }
        } else if (lane == 0 && cw == 0) {
            sRed[grp] = beta_g;
            sRed[32 + grp] = eta_lane;
        }
    }
    __syncthreads();
    if (p.norm_mode == 1) {
        float lo = kInf, hi = -kInf;
        for (int w = 0; w < NG; w++) { lo = fminf(lo, sRed[w]); hi = fmaxf(hi, sRed[32 + w]); }
        __syncthreads();
        publish_minmax(p, ctrl, lo, hi, sRed);
        return;
    }

    // ---- CTA merge of the per-group running sums ---------------------------------------------------
    float beta_c = kInf;
    for (int w = 0; w < NG; w++) beta_c = fminf(beta_c, sRed[w]);
    float eta_c = 0.f;
    for (int w = 0; w < NG; w++) {
        const float bw = sRed[w];
        if (bw != kInf) eta_c = fmaf(weight_exp(bw, beta_c, nil), sRed[32 + w], eta_c);
    }
    for (int j = tid; j < TA; j += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < NG; w++) {
            const float bw = sRed[w];
            if (bw != kInf) s = fmaf(weight_exp(bw, beta_c, nil), sAccG[w * TAp + j], s);
        }
        sN[j] = s;
    }
    __syncthreads();
    // the tile buffers are free now (every group is past its last tile): merge scratch
    publish_and_finish<A, false>(p, ctrl, beta_c, eta_c, sN, sWork, sScale, sRed, reinterpret_cast<float4 *>(sTiles),
                                 (NBUF * tile_words) >> 2);
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
    float4 *sScratch = reinterpret_cast<float4 *>(sRed + 32);
    const int ctrl = blockIdx.x, stride = partial_stride(TA);
    Merged m = merge_parts(gathered + (size_t)ctrl * stride, (size_t)p.n_ctrl * stride, p.world, TA,
                           p.neg_inv_lambda_log2e, sN, sScale, sRed, sScratch, 256);
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
        normals4_as_used((uint32_t)call, (uint32_t)(p.k_offset + k), (uint32_t)ctrl, p, z);
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
            if (p.sigma_diag) {
                e[j] = p.sigma[j * A + j] * z[j];
            } else {
                e[j] = 0.f;
#pragma unroll
                for (int l = 0; l < A; l++) e[j] = fmaf(p.sigma[j * A + l], z[l], e[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < A; j++) io[i * A + j] = e[j];
    }
}

// -------------------------------------------------------------------------------------------------
// Host launchers
// -------------------------------------------------------------------------------------------------
template <int A, bool DIAG, bool QUAD, int COST>
static cudaError_t launch_philox_V(const RolloutParams &p, dim3 grid, size_t smem, cudaStream_t st)
{
    cudaError_t err = ensure_dyn_smem<rollout_philox_kernel<A, DIAG, QUAD, COST>>(smem);
    if (err != cudaSuccess) return err;
    rollout_philox_kernel<A, DIAG, QUAD, COST><<<grid, kPhiloxThreads, smem, st>>>(p);
    return cudaGetLastError();
}

template <int A>
static cudaError_t launch_philox_A(const RolloutParams &p, dim3 grid, cudaStream_t st)
{
    const size_t smem = philox_smem_bytes(A, p.T, p.TA);
    if (p.cost_kind == 1) {                      // ElipseCost: point_mass2d only, general-Sigma variants
        if constexpr (A == 2) {
            return p.quad ? launch_philox_V<2, false, true, 1>(p, grid, smem, st) : launch_philox_V<2, false, false, 1>(p, grid, smem, st);
        } else {
            return cudaErrorInvalidValue;
        }
    }
    if (p.quad) return launch_philox_V<A, false, true, 0>(p, grid, smem, st);      // noise-quadratic cost: general-Sigma variant
    if (p.sigma_diag) return launch_philox_V<A, true, false, 0>(p, grid, smem, st);
    return launch_philox_V<A, false, false, 0>(p, grid, smem, st);
}

int philox_grid_x(int K_local, int n_ctrl, int num_sms)
{
    const int ctas_total = num_sms * kPhiloxCtasPerSm;
    int per_ctrl = ctas_total / (n_ctrl > 0 ? n_ctrl : 1);
    if (per_ctrl < 1) per_ctrl = 1;
    // at least 128 samples (4 warps) per CTA so small problems do not pay for empty CTAs
    const int need = (K_local + 127) / 128;
    int gx = need < per_ctrl ? need : per_ctrl;
    if (gx > kMaxParts) gx = kMaxParts;
    if (gx < 1) gx = 1;
    return gx;
}

cudaError_t launch_rollout_philox(RolloutParams p, int a, int num_sms, size_t smem_limit, cudaStream_t st, int *grid_x_out)
{
    if (p.fast) return cudaErrorInvalidValue;        // the superposition kernels are launched by launch_rollout_philox_fast
    // 16 per-warp rows of T*a partial sums + the sequence and call tables must fit one CTA's shared memory
    // (T*a up to about 2400 on a 227 KB part)
    if (philox_smem_bytes(a, p.T, p.TA) > smem_limit) return cudaErrorInvalidConfiguration;
    int gx = philox_grid_x(p.K_local, p.n_ctrl, num_sms);
    if (p.max_parts > 0 && gx > p.max_parts) gx = p.max_parts;
    p.n_iter = (p.K_local + gx * kPhiloxThreads - 1) / (gx * kPhiloxThreads);
    if (grid_x_out) *grid_x_out = gx;
    dim3 grid(gx, p.n_ctrl);
    MPPI_DISPATCH_A(a, return launch_philox_A<A_>(p, grid, st));
    return cudaSuccess;
}

// Injected mode geometry: tile groups / warps per group / tile buffers that fit in shared memory.
bool injected_geometry(int A, int T, int TA, int K_local, int n_ctrl, int num_sms, size_t smem_limit,
                       int *ng_out, int *c_out, int *nbuf_out, int *grid_x_out, size_t *smem_out)
{
    const int H = (A + 1) & ~1, RS = (2 * H + 3) & ~3, TAp = (TA + 31) & ~31;
    const size_t tile_b = (size_t)128 * TA;
    const int n_tiles = (K_local + 31) / 32;
    const int nblk = (T + 3) / 4;
    // several CTAs per SM when there are many small controllers
    int ctas_per_sm = 1;
    if (n_ctrl >= 2 * num_sms) ctas_per_sm = 4;
    const int max_warps = 16 / ctas_per_sm;         // 512 threads per SM: <= 128 registers per thread
    const size_t budget = smem_limit / ctas_per_sm - (ctas_per_sm > 1 ? 1024 : 0);
    const size_t fixed = sizeof(float) * ((size_t)T * RS + 2 * TAp + kMaxParts + 64) + 8 * 64 + 4 * 64 + 64;
    // per group: running sums + chunk responses/costs for up to 4 warps
    const size_t per_group = sizeof(float) * ((size_t)TAp + 4 * (2 * A + 1) * 32);
    if (fixed + per_group + tile_b > budget) return false;
    int fit = (int)((budget - fixed) / (per_group + tile_b));      // tiles that fit if every tile had its own group
    int ng = fit;
    if (ng > max_warps) ng = max_warps;
    if (ng > n_tiles) ng = n_tiles > 0 ? n_tiles : 1;
    // few large tiles (config 3: five of 38 KB): give one buffer up as a prefetch slot - every group then finds
    // its next tile in flight - and spend the warps on splitting the horizon instead
    if (ng == fit && ng >= 3 && ng < max_warps && n_tiles > ng) ng -= 1;
    int c = max_warps / ng;
    if (c > 4) c = 4;
    if (c > nblk / 6) c = nblk / 6;                 // at least 24 steps per time chunk
    if (c < 1) c = 1;
    if (c > 1 && ng > 15) c = 1;                    // named barriers 1..15
    int nbuf = (int)((budget - fixed - ng * per_group) / tile_b);
    if (nbuf > 3 * ng) nbuf = 3 * ng;
    if (nbuf < ng) nbuf = ng;
    int gx;
    const int need = (n_tiles + ng - 1) / ng;
    if (n_ctrl == 1) gx = num_sms * ctas_per_sm;
    else gx = (num_sms * ctas_per_sm) / n_ctrl;
    if (gx > need) gx = need;
    if (gx < 1) gx = 1;
    if (gx > kMaxParts) gx = kMaxParts;
    const int tiles_per_cta = (n_tiles + gx - 1) / gx;
    if (nbuf > tiles_per_cta) nbuf = tiles_per_cta > ng ? tiles_per_cta : ng;
    *ng_out = ng;
    *c_out = c;
    *nbuf_out = nbuf;
    *grid_x_out = gx;
    *smem_out = fixed + ng * per_group + (size_t)nbuf * tile_b;
    return true;
}

template <int A, bool TMA, bool QUAD, int COST, bool FAST = false>
static cudaError_t launch_injected_V(const RolloutParams &p, InjectedLaunch L, dim3 grid, size_t smem, cudaStream_t st)
{
    cudaError_t err = ensure_dyn_smem<rollout_injected_kernel<A, TMA, QUAD, COST, FAST>>(smem);
    if (err != cudaSuccess) return err;
    rollout_injected_kernel<A, TMA, QUAD, COST, FAST><<<grid, L.ng * L.c * 32, smem, st>>>(p, L);
    return cudaGetLastError();
}

template <int A>
static cudaError_t launch_injected_A(const RolloutParams &p, InjectedLaunch L, dim3 grid, size_t smem, bool tma,
                                     cudaStream_t st)
{
    if (p.cost_kind == 1) {                      // ElipseCost: point_mass2d only
        if constexpr (A == 2) {
            if (p.quad) return tma ? launch_injected_V<2, true, true, 1>(p, L, grid, smem, st) : launch_injected_V<2, false, true, 1>(p, L, grid, smem, st);
            return tma ? launch_injected_V<2, true, false, 1>(p, L, grid, smem, st) : launch_injected_V<2, false, false, 1>(p, L, grid, smem, st);
        } else {
            return cudaErrorInvalidValue;
        }
    }
    if (p.quad) return tma ? launch_injected_V<A, true, true, 0>(p, L, grid, smem, st) : launch_injected_V<A, false, true, 0>(p, L, grid, smem, st);
    if (p.fast) return tma ? launch_injected_V<A, true, false, 0, true>(p, L, grid, smem, st) : launch_injected_V<A, false, false, 0, true>(p, L, grid, smem, st);
    return tma ? launch_injected_V<A, true, false, 0>(p, L, grid, smem, st) : launch_injected_V<A, false, false, 0>(p, L, grid, smem, st);
}

cudaError_t launch_rollout_injected(RolloutParams p, int a, int num_sms, size_t smem_limit, cudaStream_t st,
                                    int *grid_x_out)
{
    InjectedLaunch L;
    int gx = 1;
    size_t smem = 0;
    if (!injected_geometry(a, p.T, p.TA, p.K_local, p.n_ctrl, num_sms, smem_limit, &L.ng, &L.c, &L.nbuf, &gx, &smem))
        return cudaErrorInvalidConfiguration;
    if (const char *ov = getenv("MPPI_INJ_GEOM")) {      // developer knob: "ng,c,nbuf" (no validation beyond the smem size)
        int ng = 0, c = 0, nb = 0;
        if (sscanf(ov, "%d,%d,%d", &ng, &c, &nb) == 3 && ng >= 1 && c >= 1 && nb >= ng && ng * c <= 16) {
            const size_t tile_b = (size_t)128 * p.TA;
            const size_t grown = smem - (size_t)L.nbuf * tile_b + (size_t)nb * tile_b + 4096;
            if (grown <= smem_limit) { L.ng = ng; L.c = c; L.nbuf = nb; smem = grown; }
        }
    }
    if (p.max_parts > 0 && gx > p.max_parts) gx = p.max_parts;
    if (grid_x_out) *grid_x_out = gx;
    // TMA path needs 16-byte aligned tiles and rows: T*a % 4 == 0 and an aligned base pointer.
    const bool tma = (p.TA % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.eps) & 15u) == 0);
    dim3 grid(gx, p.n_ctrl);
    MPPI_DISPATCH_A(a, return launch_injected_A<A_>(p, L, grid, smem, tma, st));
    return cudaSuccess;
}

// Upper bound of grid.x over every update kernel (the handle sizes its partial-record buffer with it and passes it back
// as RolloutParams::max_parts): the regenerating kernels use at most 2 CTAs per SM, the injected and resident kernels at
// most 4 resident CTAs per SM, shared among the controllers.
int max_grid_x(int K_local, int n_ctrl, int num_sms)
{
    int g1 = philox_grid_x(K_local, n_ctrl, num_sms);
    int g2 = (num_sms * 8) / (n_ctrl > 0 ? n_ctrl : 1);
    if (g2 < 1) g2 = 1;
    if (g2 > kMaxParts) g2 = kMaxParts;
    return g1 > g2 ? g1 : g2;
}

cudaError_t launch_finish(RolloutParams p, int a, bool philox, const float *gathered, cudaStream_t st)
{
    const int TAp = (p.TA + 31) & ~31;
    const size_t smem = sizeof(float) * (2 * (size_t)TAp + kMaxParts + 32) + sizeof(float4) * 256;
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
