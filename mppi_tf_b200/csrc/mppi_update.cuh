// Update-side device code shared by every rollout kernel: the per-step sequence table, the
// deterministic merge of {beta, eta, N} partial records, the sequence update / shift, and the
// last-CTA election.  (src/controller_base.cpp:166-192,215-224,310-329 of the reference.)
#pragma once
#include "mppi_device.cuh"

// The merge of partial records runs once per launch, in one CTA, after everyone else has finished, up to three times in a
// row (group, groups, ranks): a real function (one copy per kernel, warm on its second call) with rolled loops - the
// tail's cost is instruction fetch and dependent L2 trips, not arithmetic.  A kernel that re-partitions its registers
// with setmaxnreg (mppi_mlp.cu) cannot call functions: it defines MPPI_TAIL_INLINE before including this header.
#ifdef MPPI_TAIL_INLINE
#define MPPI_TAIL_FN inline
#else
#define MPPI_TAIL_FN __noinline__
#endif

namespace mppi {

constexpr int kMaxParts = 1024;   // CTA partials per controller
constexpr int kMergeGroup = 16;   // two-level merge: the last CTA of every group of 16 merges the group's records, the last
                                  // group merges the group records - the serial tail after the last CTA finishes is two
                                  // short merges (16 and <= 64 records) instead of one over every CTA of the grid

// Per-step uniforms staged in shared memory, one row per step:
//   [0, A)      U_t   mean action (src/controller_base.cpp:205-208)
//   [H, H + A)  w_t   action-cost vector: Philox mode lambda*U_t (since eps = Sigma z,
//                     lambda U^T Sigma^-1 eps = lambda U^T z); injected mode lambda*Sigma^-T U_t
//                     (src/cost_base.cpp:63-68).  With the Python-twin action cost gamma takes the
//                     place of lambda and eps = upsilon Sigma z (w_scale / lam_inv_sigma_T, host side)
// H = A rounded up to even so both halves start on an aligned pair.
template <int A>
struct Row {
    static constexpr int H = (A + 1) & ~1;
    static constexpr int RS = (2 * H + 3) & ~3;
};

template <int A, bool PHILOX>
__device__ __forceinline__ void stage_sequence(const RolloutParams &p, int ctrl, float *sUV)
{
    constexpr int RS = Row<A>::RS, H = Row<A>::H;
    const float *U = p.U + (size_t)ctrl * p.TA;
    for (int i = threadIdx.x; i < p.T * RS; i += blockDim.x) {
        const int t = i / RS, j = i - t * RS;
        float v = 0.f;
        if (j < A) {
            v = U[t * A + j];
        } else if (j >= H && j < H + A) {
            const int r = j - H;
            if (PHILOX) {
                v = p.w_scale * U[t * A + r];
            } else {
#pragma unroll
                for (int l = 0; l < A; l++) v = fmaf(p.lam_inv_sigma_T[r * A + l], U[t * A + l], v);
            }
        }
        sUV[i] = v;
    }
}

// Sample-independent part of the Python-twin action cost, summed over the horizon:
//   C0 = sum_t c0_scale * U_t^T Sigma^-1 U_t   (0.5*gamma*u^T Sigma^-1 u, cost_base.py:151-155,165-167)
// Block-uniform result; fixed summation order.  sTmp: >= T floats, sRed: >= 1 float.  Contains CTA barriers.
template <int A>
__device__ __forceinline__ float stage_c0(const RolloutParams &p, int ctrl, float *sTmp, float *sRed)
{
    if (p.c0_scale == 0.f) return 0.f;             // grid-uniform
    const float *U = p.U + (size_t)ctrl * p.TA;
    for (int t = threadIdx.x; t < p.T; t += blockDim.x) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < A; i++) {
            float r = 0.f;
#pragma unroll
            for (int j = 0; j < A; j++) r = fmaf(p.inv_sigma[i * A + j], U[t * A + j], r);
            acc = fmaf(U[t * A + i], r, acc);
        }
        sTmp[t] = p.c0_scale * acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float c = 0.f;
        for (int t = 0; t < p.T; t++) c += sTmp[t];
        sRed[0] = c;
    }
    __syncthreads();
    const float c0 = sRed[0];
    __syncthreads();
    return c0;
}

// n^T quadm n for one step (Python-twin noise cost 0.5*lambda*(1-1/upsilon) eps^T Sigma^-1 eps)
template <int A>
__device__ __forceinline__ float quad_cost(const RolloutParams &p, const float (&n)[A])
{
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < A; i++) {
        float r = 0.f;
#pragma unroll
        for (int j = 0; j < A; j++) r = fmaf(p.quadm[i * A + j], n[j], r);
        acc = fmaf(n[i], r, acc);
    }
    return acc;
}

// Cost normalisation, pass 1 (controller_base.py:468-474 needs max_k(S_k - beta) before any weight):
// every CTA publishes its (min, max); the last one reduces them into norm[ctrl] = {beta, max - beta}.
__device__ inline void publish_minmax(const RolloutParams &p, int ctrl, float bmin, float bmax, float *sRed)
{
    __shared__ int s_last_mm;
    const int stride = partial_stride(p.TA), nparts = gridDim.x;
    float *mine = p.partials + ((size_t)ctrl * nparts + blockIdx.x) * stride;
    if (threadIdx.x == 0) { mine[0] = bmin; mine[1] = bmax; }
    __threadfence();
    __syncthreads();
    unsigned int *ctr = p.counters + (size_t)ctrl * (1 + p.max_groups);
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(ctr, 1u);
        s_last_mm = (prev == (unsigned)nparts - 1u);
    }
    __syncthreads();
    if (!s_last_mm) return;
    __threadfence();
    if (threadIdx.x == 0) *ctr = 0u;
    const float *parts = p.partials + (size_t)ctrl * nparts * stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    float lo = kInf, hi = -kInf;
    for (int c = tid; c < nparts; c += blockDim.x) {
        lo = fminf(lo, __ldcg(parts + (size_t)c * stride));
        hi = fmaxf(hi, __ldcg(parts + (size_t)c * stride + 1));
    }
    lo = warp_min(lo);
    hi = -warp_min(-hi);
    if (lane == 0) { sRed[warp] = lo; sRed[32 + warp] = hi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < nw; w++) { lo = fminf(lo, sRed[w]); hi = fmaxf(hi, sRed[32 + w]); }
        sRed[0] = lo;
        sRed[32] = hi;
    }
    __syncthreads();
    lo = sRed[0];
    hi = sRed[32];
    if (p.world > 1) {
        // sharded samples: the range must be global before any rank forms a weight.  Same mailbox hand-over as
        // the payload exchange (publish_and_finish), with the (min, max) pair in the first two payload words.
        const int world = p.world;
        const uint32_t par = p.epoch & 1u;
        const size_t slot = ((size_t)par * world + p.rank) * p.n_ctrl + ctrl;
        if (tid < world) {
            float *dst = p.peer_mail[tid] + slot * stride;
            dst[0] = lo;
            dst[1] = hi;
        }
        __threadfence_system();
        __syncthreads();
        if (tid < world) st_release_sys(p.peer_flag[tid] + slot, p.epoch);
        if (tid < world) {
            const size_t src = ((size_t)par * world + tid) * p.n_ctrl + ctrl;
            const uint32_t *f = p.peer_flag[p.rank] + src;
            const long long t0 = clock64();
            while (ld_acquire_sys(f) != p.epoch) {
                if (clock64() - t0 > (1LL << 31)) {
                    atomicExch(p.peer_status, 1u);
                    break;
                }
            }
            const float *m = p.peer_mail[p.rank] + src * stride;
            sRed[tid] = __ldcg(m);
            sRed[32 + tid] = __ldcg(m + 1);
        }
        __syncthreads();
        if (tid == 0) {
            lo = sRed[0];
            hi = sRed[32];
            for (int r = 1; r < world; r++) { lo = fminf(lo, sRed[r]); hi = fmaxf(hi, sRed[32 + r]); }
        }
    }
    if (tid == 0) {
        p.norm[2 * ctrl] = lo;
        p.norm[2 * ctrl + 1] = hi - lo;
    }
}

// Weight exponent scale of the update: -log2(e)/lambda, divided by max(S - beta) in the weight pass of a
// normalised update (all costs equal: every weight is 1, where the reference divides 0 by 0).
__device__ __forceinline__ float weight_scale(const RolloutParams &p, int ctrl, float &beta_fixed)
{
    if (p.norm_mode != 2) return p.neg_inv_lambda_log2e;
    beta_fixed = p.norm[2 * ctrl];
    const float r = p.norm[2 * ctrl + 1];
    return r > 0.f ? p.neg_inv_lambda_log2e / r : 0.f;
}

// -------------------------------------------------------------------------------------------------
// Merge of partial records {beta, eta, -, -, N[TA]} (CTA partials or rank payloads), block-wide and
// in a fixed order (deterministic).  On return sN[0..TA) holds sum_c scale_c N_c and the returned
// beta/eta are the merged values: beta = min_c beta_c, scale_c = exp(-(beta_c - beta)/lambda).
// -------------------------------------------------------------------------------------------------
struct Merged { float beta, eta; };

// sScratch: `scratch_f4` float4 of shared memory, free to clobber.  The column sums are the serial tail
// of every update (one CTA reads nparts x TA floats through L2), so they are spread over the whole CTA:
// work item (slice, column quad) sums every nsl-th record with eight 16-byte loads in flight, then the
// slices are added in a fixed order.
static __device__ MPPI_TAIL_FN Merged merge_parts(const float *parts, size_t part_stride, int nparts, int TA,
                              float neg_inv_lambda_log2e, float *sN, float *sScale, float *sRed,
                              float4 *sScratch, int scratch_f4)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int ncol4 = (TA + 3) >> 2;               // records are padded to whole quads (partial_stride)
    int nsl = (int)blockDim.x / ncol4;
    if (nsl > scratch_f4 / ncol4) nsl = scratch_f4 / ncol4;
    if (nsl > nparts) nsl = nparts;
    if (nsl < 1) nsl = 1;
    // The merge is the serial tail of every update and its cost is L2 round trips, so everything that does not depend
    // on beta is requested first: the first eight records of this thread's first (slice, column quad) item ...
    const bool has_item = tid < ncol4 * nsl;
    const int sl0 = has_item ? tid / ncol4 : 0, c40 = has_item ? tid - sl0 * ncol4 : 0;
    constexpr int PRE = 4;
    float4 pre[PRE];
#pragma unroll
    for (int i = 0; i < PRE; i++) {
        const int c = sl0 + i * nsl;
        pre[i] = (has_item && c < nparts) ? __ldcg(reinterpret_cast<const float4 *>(parts + 4 + 4 * c40 + (size_t)c * part_stride))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ... and (beta, eta) of the records, one 8-byte load each
    float2 be = make_float2(kInf, 0.f);
    if (tid < nparts) be = __ldcg(reinterpret_cast<const float2 *>(parts + (size_t)tid * part_stride));
    float b = be.x;
#pragma unroll 1
    for (int c = tid + blockDim.x; c < nparts; c += blockDim.x) b = fminf(b, __ldcg(parts + c * part_stride));
    b = warp_min(b);
    if (lane == 0) sRed[warp] = b;
    __syncthreads();
    float beta = sRed[0];
#pragma unroll 1
    for (int w = 1; w < nw; w++) beta = fminf(beta, sRed[w]);
    __syncthreads();
    float e = 0.f;
    if (tid < nparts) {
        const float sc = (be.x == kInf) ? 0.f : weight_exp(be.x, beta, neg_inv_lambda_log2e);
        sScale[tid] = sc;
        e = sc * be.y;
    }
#pragma unroll 1
    for (int c = tid + blockDim.x; c < nparts; c += blockDim.x) {
        const float bc = __ldcg(parts + c * part_stride);
        const float sc = (bc == kInf) ? 0.f : weight_exp(bc, beta, neg_inv_lambda_log2e);
        sScale[c] = sc;
        e = fmaf(sc, __ldcg(parts + c * part_stride + 1), e);
    }
    e = warp_sum(e);
    if (lane == 0) sRed[warp] = e;
    __syncthreads();
    float eta = 0.f;
#pragma unroll 1
    for (int w = 0; w < nw; w++) eta += sRed[w];

    // column sums: work item (slice, column quad) sums every nsl-th record, eight 16-byte loads in flight, in order
#pragma unroll 1
    for (int it = tid; it < ncol4 * nsl; it += blockDim.x) {
        const int sl = it / ncol4, c4 = it - sl * ncol4;
        const float *col = parts + 4 + 4 * c4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int c = sl;
        if (it == tid) {                                   // the prefetched batch
#pragma unroll
            for (int i = 0; i < PRE; i++) {
                if (c + i * nsl < nparts) {
                    const float w = sScale[c + i * nsl];
                    acc.x = fmaf(w, pre[i].x, acc.x); acc.y = fmaf(w, pre[i].y, acc.y);
                    acc.z = fmaf(w, pre[i].z, acc.z); acc.w = fmaf(w, pre[i].w, acc.w);
                }
            }
            c += PRE * nsl;
        }
#pragma unroll 1
        for (; c + (PRE - 1) * nsl < nparts; c += PRE * nsl) {
            float4 v[PRE];
#pragma unroll
            for (int i = 0; i < PRE; i++) v[i] = __ldcg(reinterpret_cast<const float4 *>(col + (size_t)(c + i * nsl) * part_stride));
#pragma unroll
            for (int i = 0; i < PRE; i++) {
                const float w = sScale[c + i * nsl];
                acc.x = fmaf(w, v[i].x, acc.x); acc.y = fmaf(w, v[i].y, acc.y);
                acc.z = fmaf(w, v[i].z, acc.z); acc.w = fmaf(w, v[i].w, acc.w);
            }
        }
#pragma unroll 1
        for (; c < nparts; c += nsl) {
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(col + (size_t)c * part_stride));
            const float w = sScale[c];
            acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
            acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
        }
        if (nsl > 1) {
            sScratch[it] = acc;
        } else {                                           // pad columns of the last quad hold no data
            const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (4 * c4 + i < TA) sN[4 * c4 + i] = a4[i];
        }
    }
    __syncthreads();
    if (nsl > 1) {
        const float *sc = reinterpret_cast<const float *>(sScratch);
#pragma unroll 1
        for (int j = tid; j < TA; j += blockDim.x) {
            float acc = 0.f;
#pragma unroll 1
            for (int sl = 0; sl < nsl; sl++) acc += sc[(size_t)sl * 4 * ncol4 + j];
            sN[j] = acc;
        }
        __syncthreads();
    }
    return Merged{beta, eta};
}

// -------------------------------------------------------------------------------------------------
// Tagged words for the rank exchange: every 4-byte value of a rank payload travels over NVLink together with the
// update's epoch in ONE 8-byte store (single-copy atomic), and the receiver polls the words in its own mailbox until
// the epoch shows: no system fence (a round trip to the farthest peer) and no flag hand-over stand between the
// sender's last store and the receiver's first load.  The region is zeroed at creation and epochs start at 1.
// (The same scheme for the CTA records inside one GPU was measured slower than fence + atomic, section 3.4 of DESIGN.md.)
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(uint2 *dst, float v, uint32_t tag)
{
    asm volatile("st.relaxed.sys.global.v2.b32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
// two consecutive words {v0, tag0, v1, tag1}; src 16-byte aligned
__device__ __forceinline__ uint4 ll_load2(const uint2 *src)
{
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src) : "memory");
    return v;
}
// Polls until both words carry `tag`; after about a second the status word is raised and the poll gives up - and so does
// every later poll of the CTA at once (a missing rank costs one timeout, not one per word).
__device__ __forceinline__ uint4 ll_wait2(const uint2 *src, uint4 v, uint32_t tag, unsigned int *status)
{
    if (v.y == tag && v.w == tag) return v;
    const long long t0 = clock64();
#pragma unroll 1
    for (;;) {
        v = ll_load2(src);
        if (v.y == tag && v.w == tag) return v;
        if (clock64() - t0 > (1LL << 31) || *reinterpret_cast<volatile unsigned int *>(status) != 0u) {
            atomicExch(status, 1u);
            return v;
        }
    }
}
// Merge of the `world` rank payloads in this rank's mailbox (tagged words), polling for them.  Same result contract as
// merge_parts; records are added in rank order per column (independent of the CTA size, identical on every rank).
static __device__ MPPI_TAIL_FN Merged merge_world_ll(const uint2 *base, size_t rec_stride, uint32_t tag, unsigned int *status, int world,
                                                     int TA, float neg_inv_lambda_log2e, float *sN)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const int ncol4 = (TA + 3) >> 2;
    // requested together: (beta, eta) of rank `lane` (every warp) and the first four ranks' words of this thread's column quad
    uint4 hraw = make_uint4(0u, 0u, 0u, 0u);
    if (lane < world) hraw = ll_load2(base + (size_t)lane * rec_stride);
    uint4 ra[4], rb[4];
    if (tid < ncol4) {
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (q < world) {
                const uint2 *src = base + (size_t)q * rec_stride + 4 + 4 * tid;
                ra[q] = ll_load2(src);
                rb[q] = ll_load2(src + 2);
            }
    }
    float b = kInf, e = 0.f;
    if (lane < world) {
        const uint4 h = ll_wait2(base + (size_t)lane * rec_stride, hraw, tag, status);
        b = __uint_as_float(h.x);
        e = __uint_as_float(h.z);
    }
    const float beta = warp_min(b);
    const float sc_lane = (b == kInf) ? 0.f : weight_exp(b, beta, neg_inv_lambda_log2e);
    const float eta = warp_sum(sc_lane * e);
#pragma unroll 1
    for (int c0 = 0; c0 < ncol4; c0 += blockDim.x) {          // warp-uniform trip count (the weights travel by shuffle)
        const int c4 = c0 + tid;
        const bool live = c4 < ncol4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int r0 = 0; r0 < world; r0 += 4) {
            if (live && (c0 | r0) != 0) {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (r0 + q < world) {
                        const uint2 *src = base + (size_t)(r0 + q) * rec_stride + 4 + 4 * c4;
                        ra[q] = ll_load2(src);
                        rb[q] = ll_load2(src + 2);
                    }
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const bool on = r0 + q < world;                // warp-uniform
                if (!on) continue;
                const float w = __shfl_sync(0xffffffffu, sc_lane, r0 + q);
                if (live) {
                    const uint2 *src = base + (size_t)(r0 + q) * rec_stride + 4 + 4 * c4;
                    const uint4 a = ll_wait2(src, ra[q], tag, status), bb = ll_wait2(src + 2, rb[q], tag, status);
                    acc.x = fmaf(w, __uint_as_float(a.x), acc.x); acc.y = fmaf(w, __uint_as_float(a.z), acc.y);
                    acc.z = fmaf(w, __uint_as_float(bb.x), acc.z); acc.w = fmaf(w, __uint_as_float(bb.z), acc.w);
                }
            }
        }
        if (live) {
            const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (4 * c4 + i < TA) sN[4 * c4 + i] = a4[i];
        }
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
    const float inv_eta = (PHILOX ? p.z_scale : 1.0f) / m.eta;     // sN holds sum e n, n = z / z_scale (1 off the superposition path)
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
        float un = U[i] + d * inv_eta;
        if (p.clip) un = fminf(fmaxf(un, p.act_min[j]), p.act_max[j]);      // tf.clip_by_value (controller_base.py:500-504)
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
    if (p.next_host != nullptr) {              // zero-copy result for the host (grid-uniform; n_ctrl == 1)
        // Every word travels with the update's epoch in ONE 8-byte store to mapped pinned memory (slot i = {action_i, epoch},
        // slot 8 = {exchange status, epoch}): a slot is valid as soon as its epoch shows, whatever order the stores arrive
        // in, so no system-scope fence (a PCIe round trip each) is needed before the host may read.
        if (threadIdx.x < A || threadIdx.x == 8) {
            const unsigned int val = threadIdx.x < A ? __float_as_uint(sOut[threadIdx.x])
                                                     : (p.peer_on ? *reinterpret_cast<volatile unsigned int *>(p.peer_status) : 0u);
            unsigned int *slot = reinterpret_cast<unsigned int *>(p.next_host) + 2 * threadIdx.x;
            asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(slot), "r"(val), "r"(p.done_epoch) : "memory");
        }
    }
}

// Publish this CTA's partial; the last CTA of the controller merges and finishes the update.
// sN: CTA sums [TA]; sWork: >= TA floats scratch; sScale: kMaxParts floats; sRed: 32 floats;
// sScratch: scratch_f4 float4 (16-byte aligned) the merge may clobber.
template <int A, bool PHILOX>
__device__ void publish_and_finish(const RolloutParams &p, int ctrl, float beta_c, float eta_c, float *sN,
                                   float *sWork, float *sScale, float *sRed, float4 *sScratch, int scratch_f4)
{
    __shared__ int s_is_last;
    const int TA = p.TA, stride = partial_stride(TA);
    int nparts = gridDim.x, part = blockIdx.x;
    trace_stamp(p, ctrl, 3);
    // Launched as thread-block clusters (small grids: latency matters, capacity does not): the CTAs of a cluster - always
    // the same controller - are reduced over distributed shared memory first, by the rank-0 CTA, and only that CTA
    // goes on: cluster-size times fewer partial records, and none at all when one cluster owns the controller.
    const uint32_t cs = cluster_nctarank();
    if (cs > 1) {
        const uint32_t cr = cluster_ctarank();
        if (threadIdx.x == 0) { sRed[0] = beta_c; sRed[1] = eta_c; }
        cluster_sync();
        if (cr == 0) {
            float br[8], er[8];
            float b = beta_c;
#pragma unroll
            for (uint32_t r = 1; r < 8; r++) {
                br[r] = (r < cs) ? ld_dsmem(sRed, r) : kInf;
                er[r] = (r < cs) ? ld_dsmem(sRed + 1, r) : 0.f;
            }
#pragma unroll
            for (uint32_t r = 1; r < 8; r++) b = fminf(b, br[r]);
            const float nil = p.neg_inv_lambda_log2e;
            const float sc0 = (beta_c == kInf) ? 0.f : weight_exp(beta_c, b, nil);
            float eta = sc0 * eta_c;
#pragma unroll
            for (uint32_t r = 1; r < 8; r++) {
                br[r] = (br[r] == kInf) ? 0.f : weight_exp(br[r], b, nil);      // now the scale of rank r (0 beyond the cluster)
                eta = fmaf(br[r], er[r], eta);
            }
            for (int j = threadIdx.x; j < TA; j += blockDim.x) {
                float v[8];
#pragma unroll
                for (uint32_t r = 1; r < 8; r++) v[r] = (r < cs) ? ld_dsmem(sN + j, r) : 0.f;   // all in flight together
                float acc = sc0 * sN[j];
#pragma unroll
                for (uint32_t r = 1; r < 8; r++)
                    if (r < cs) acc = fmaf(br[r], v[r], acc);                                   // fixed order
                sN[j] = acc;
            }
            beta_c = b;
            eta_c = eta;
        }
        cluster_sync();                         // the peers' shared memory has been read: they may leave
        if (cr != 0) return;
        nparts = gridDim.x / cs;
        part = blockIdx.x / cs;
    }
    Merged m{beta_c, eta_c};
    const int tid = threadIdx.x;
    if (!(nparts == 1 && p.world == 1)) {       // else a single CTA (or cluster) owns the controller: nothing to merge
        MPPI_CHECK(part >= 0 && part < nparts && nparts <= p.max_parts && ctrl < p.n_ctrl);
        // record = {beta, eta, 0, 0, N[TA], 0 ..}
        auto store_record = [&](float *dst, float b, float e) {
#pragma unroll 1
            for (int j = tid; j < stride; j += blockDim.x) dst[j] = j == 0 ? b : (j == 1 ? e : ((j >= 4 && j - 4 < TA) ? sN[j - 4] : 0.f));
        };
        // Last-arriver election after the CTA's stores: the barrier orders every thread's stores before thread 0's fence
        // (cumulative), one fence and one atomic per CTA instead of a fence in every thread; the winner's fence and the
        // second barrier order the atomic before the CTA's L2 loads.  CTA-wide (two barriers); true in the last CTA.
        auto last_of = [&](unsigned int *counter, int n) {
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                const bool last = atomicAdd(counter, 1u) == (unsigned)n - 1u;
                if (last) {
                    __threadfence();
                    *counter = 0u;              // ready for the next launch
                }
                s_is_last = last;
            }
            __syncthreads();
            return s_is_last != 0;
        };
        store_record(p.partials + ((size_t)ctrl * nparts + part) * stride, beta_c, eta_c);
        trace_stamp(p, ctrl, 4);
        unsigned int *ctr = p.counters + (size_t)ctrl * (1 + p.max_groups);
        const int ngroups = (nparts + kMergeGroup - 1) / kMergeGroup;
        const bool two_level = nparts > 2 * kMergeGroup && ngroups <= p.max_groups;
        if (!two_level) {
            if (!last_of(ctr, nparts)) return;
            m = merge_parts(p.partials + (size_t)ctrl * nparts * stride, stride, nparts, TA,
                            p.neg_inv_lambda_log2e, sN, sScale, sRed, sScratch, scratch_f4);
        } else {
            const int gi = part / kMergeGroup, g0 = gi * kMergeGroup;
            const int gn = min(kMergeGroup, nparts - g0);
            MPPI_CHECK(gi >= 0 && gi < ngroups && ngroups <= p.max_groups && gn >= 1);
            if (!last_of(ctr + 1 + gi, gn)) return;
            const Merged mg = merge_parts(p.partials + ((size_t)ctrl * nparts + g0) * stride, stride, gn, TA,
                                          p.neg_inv_lambda_log2e, sN, sScale, sRed, sScratch, scratch_f4);
            store_record(p.partials2 + ((size_t)ctrl * p.max_groups + gi) * stride, mg.beta, mg.eta);
            if (!last_of(ctr, ngroups)) return;
            m = merge_parts(p.partials2 + (size_t)ctrl * p.max_groups * stride, stride, ngroups, TA,
                            p.neg_inv_lambda_log2e, sN, sScale, sRed, sScratch, scratch_f4);
        }
        trace_stamp(p, ctrl, 5);
        if (p.world > 1 && !p.peer_on) {        // NCCL / caller-side exchange: the payload, then finish_kernel
            store_record(p.payload + (size_t)ctrl * stride, m.beta, m.eta);
            return;
        }
        if (p.world > 1) {
            // Fused exchange: this CTA writes the rank payload as tagged words straight into every rank's mailbox over
            // NVLink, polls its own mailbox for the other ranks' words and finishes the update in the same launch - no
            // collective call, no second kernel, no system fence, no flag.  Buffers alternate with the epoch's parity, so
            // a rank that races ahead (by at most one update: its next exchange needs this rank's next payload) cannot
            // overwrite words that are still being read.
            const int world = p.world;
            const uint32_t par = p.epoch & 1u;
            const size_t slot = ((size_t)par * world + p.rank) * p.n_ctrl + ctrl;
#pragma unroll 1
            for (int r = 0; r < world; r++) {
                uint2 *dst = p.peer_ll[r] + slot * stride;
#pragma unroll 1
                for (int j = tid; j < stride; j += blockDim.x)
                    ll_store(dst + j, j == 0 ? m.beta : (j == 1 ? m.eta : ((j >= 4 && j - 4 < TA) ? sN[j - 4] : 0.f)), p.epoch);
            }
            __syncthreads();                    // the payload has been read out of sN
            m = merge_world_ll(p.peer_ll[p.rank] + ((size_t)par * world * p.n_ctrl + ctrl) * stride, (size_t)p.n_ctrl * stride, p.epoch,
                               p.peer_status, world, TA, p.neg_inv_lambda_log2e, sN);
            trace_stamp(p, ctrl, 6);
            if (*reinterpret_cast<volatile unsigned int *>(p.peer_status) != 0u) {
                // a payload is missing: leave U untouched (merging a stale mailbox would let the ranks' sequences diverge) and
                // hand the status to a host that waits on the zero-copy slots; mppi_fetch_action reports MPPI_ERR_COMM and
                // clears the word
                if (p.next_host != nullptr && (tid < A || tid == 8)) {
                    unsigned int *slot = reinterpret_cast<unsigned int *>(p.next_host) + 2 * tid;
                    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(slot), "r"(tid == 8 ? 1u : 0u), "r"(p.done_epoch) : "memory");
                }
                return;
            }
        }
    }
    apply_update<A, PHILOX>(p, ctrl, m, sN, sWork);      // the one copy of it
    trace_stamp(p, ctrl, 7);
}

}  // namespace mppi
