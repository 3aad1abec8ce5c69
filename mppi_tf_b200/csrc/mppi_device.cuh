// Device-side building blocks shared by the rollout kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mppi_b200.h"

// Bounds checks of every computed shared / global index in the rollout kernels, compiled in by `make debug`
// (-DMPPI_DEBUG_BOUNDS -> mppi_tf_b200/_build_dbg/libmppi_b200.so, loaded through MPPI_B200_LIB): the stand-in for
// compute-sanitizer, which is closed on this pool.  A violated check traps the kernel (the launch then reports an error).
#ifdef MPPI_DEBUG_BOUNDS
#include <cassert>
#define MPPI_CHECK(cond) assert(cond)
#else
#define MPPI_CHECK(cond) ((void)0)
#endif

namespace mppi {

constexpr int kMaxA = MPPI_MAX_A;
constexpr int kMaxWorld = MPPI_MAX_PEERS;
constexpr int kMaxS = MPPI_MAX_S;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kInf = __builtin_huge_valf();

// Exchange / partial record layout: {beta, eta, 0, 0, N[TA padded to 4]}.
__host__ __device__ inline int partial_stride(int TA) { return 4 + ((TA + 3) & ~3); }

// Kernel parameters (one struct, passed by value as __grid_constant__).
struct RolloutParams {
    // sizes
    int K_local;      // samples of this rank, per controller
    int k_offset;     // global index of the first local sample (Philox counter word 1)
    int T, TA;        // horizon, T * a
    int n_ctrl;
    int n_iter;       // samples per thread in the Philox kernel
    int world;        // > 1: the last CTA writes the rank payload instead of applying the update
    int max_parts;    // partial records per controller the handle has room for: every launcher keeps grid.x within it
    // model (ModelBase: A = blkdiag([[1,dt],[0,1]]), B = blkdiag([[dt^2/2],[dt]]) / mass)
    float dt, c_pu, c_vu;
    // cost (CostBase): lambda, diag(Q), Sigma (scale) and lambda * Sigma^-T (action cost)
    float lambda, neg_inv_lambda_log2e;
    float q[kMaxS];
    float sqrt_q[kMaxS];                    // state cost as sum (sqrt(q_i) x_i - sqrt(q_i) g_i)^2
    float sigma[kMaxA * kMaxA];
    float lam_inv_sigma_T[kMaxA * kMaxA];   // v_t = lam_inv_sigma_T * U_t  (injected mode)
    int sigma_diag;
    int goal_per_ctrl;
    // Python-twin extras (scripts/src/costs/cost_base.py:114-170, controllers/controller_base.py:368,468-474):
    //   action cost = c0_t + w_t . n + n^T quadm n   per step, n = z (Philox mode) or eps (injected mode)
    float w_scale;                          // Philox mode: w_t = w_scale * U_t  (lambda*upsilon, or gamma*upsilon)
    float c0_scale;                         // c0_t = c0_scale * U_t^T Sigma^-1 U_t  (0.5*gamma in the Python form, else 0)
    float inv_sigma[kMaxA * kMaxA];         // Sigma^-1 (for c0_t)
    int quad;                               // != 0: the quadratic noise term is present
    float quadm[kMaxA * kMaxA];             // 0.5*lambda*(1-1/upsilon) * M, M = Sigma^-1 (n = eps) or (uS)^T S^-1 (uS) (n = z)
    // state-cost functor: 0 = StaticCost (x-g)^T Q (x-g); 1 = ElipseCost (scripts/src/costs/elipse_cost.py:46-79),
    // point_mass2d only: ell = {1/a, 1/b, cx, cy, speed, m_state, m_vel}; AUV state: 2 = StaticQuatCost (q[10]),
    // 3 = ElipseCost3D (elipse_cost.py:99-246): ell = {plane quaternion (x,y,z,w), 1/a, 1/b, -a/b, b/a, speed^2, m_state, m_vel}
    int cost_kind;
    float ell[12];
    int clip;                               // clip_act (controller_base.py:500-504): U' = clip(U + Delta, act_min, act_max)
    float act_min[kMaxA], act_max[kMaxA];
    int norm_mode;                          // cost normalisation: 0 off, 1 = cost pass (min/max only), 2 = weight pass
    float *norm;                            // [n_ctrl][2] beta, max(S - beta) written by pass 1, read by pass 2
    // Philox key / counter words
    uint32_t key0, key1, update;
    uint32_t rk0[10], rk1[10];              // Philox round keys key + r * W (hoisted to the host)
    int rounds;                             // Philox4x32-R: 10 (Random123 / TensorFlow's RandomNormal) or 7 (the shortest
                                            // Crush-resistant variant of Salmon et al.; mppi_config.philox_rounds)
    // Superposition form of the rollout (diagonal Sigma, q > 0, StaticCost, no noise-quadratic term; mppi_linear.cuh):
    // the state splits into the noise-free trajectory (per-CTA tables) and a noise-driven part (P, V) per sample,
    //   P' = P + fa1 V + fb1 n,  V' = V + fb2 n,  S = C + sum_t w_t (P_t^2 + V_t^2) + sum_t L_t . n_t
    // n = z / z_scale is what the generator returns without its last multiply (z_scale = sqrt(2 ln 2) folded into
    // fb1, fb2, L and applied once more in apply_update); z_scale = 1 on every other path.
    int fast;
    float *cost_base;                       // [n_ctrl]: on the superposition path costs[] holds S_k - C and the weights are
                                            // formed from those differences (the part all samples share never enters an
                                            // exponent); C is added back on the host (mppi_get_costs / mppi_get_weight_stats)
    float z_scale;
    float fa1[kMaxA], fb1[kMaxA], fb2[kMaxA];
    // single-controller fast path: state passed by value (no H2D copy)
    int x_inline;
    float x0[kMaxS];
    // device buffers
    const float *x;          // [n_ctrl][s]
    const float *goal;       // [1 or n_ctrl][s]
    float *U;                // [n_ctrl][T][a]   in: mean sequence, out: shifted sequence
    float *U_new;            // [n_ctrl][T][a]   pre-shift update
    float *next;             // [n_ctrl][a]
    float *costs;            // [n_ctrl][K_local]
    float *partials;         // [n_ctrl][gridDim.x][stride]
    float *partials2;        // [n_ctrl][max_groups][stride]  group records of the two-level merge (mppi_update.cuh)
    int max_groups;
    float *payload;          // [n_ctrl][stride]   (world > 1)
    float *stats;            // [n_ctrl][2]        beta, eta of the last update
    unsigned int *counters;  // [n_ctrl][1 + max_groups] last-CTA election: top level, then one counter per merge group
    const float *eps;        // injected noise [n_ctrl][K_local][T][a] or nullptr
    // fused exchange over peer memory (world > 1, mppi_peer_attach): every rank's mailbox
    //   mail [2][world][n_ctrl][stride] floats + flag [2][world][n_ctrl] epochs (the (min, max) exchange of a normalised update),
    //   then the tagged words of the payload exchange; mapped into this process
    int peer_on, rank;
    uint32_t epoch;          // exchange epoch of this update (> 0, same on every rank); parity selects the buffer
    float *peer_mail[kMaxWorld];
    uint32_t *peer_flag[kMaxWorld];
    uint2 *peer_ll[kMaxWorld];   // [2][world][n_ctrl][stride] tagged words {value, epoch}: the payload exchange (mppi_update.cuh)
    unsigned int *peer_status;   // set != 0 if a peer's payload did not arrive in time
    // zero-copy result (single controller): the finishing CTA also stores the action into mapped pinned host
    // memory and publishes `done_epoch`, so the synchronous next() needs neither a D2H copy nor a stream sync
    // developer knob (MPPI_TRACE=1 at mppi_create, mppi_debug_trace): %globaltimer stamps [n_ctrl][gridDim.x][12] of the phases
    //   0 start, 1 tables built, 2 rollout done, 3 weighted sum done, 4 partial published, 5 partials merged (last CTA),
    //   6 peer payloads in (last CTA), 7 update applied (last CTA); 8 weights / list built, 9 list walked (regenerating kernels)
    unsigned long long *trace;
    float *next_host;            // device alias of the mapped host buffer [a], or nullptr
    unsigned int *done_host;     // device alias of the mapped completion word
    unsigned int done_epoch;
};

constexpr int kTraceSlots = 12;
__device__ __forceinline__ void trace_stamp(const RolloutParams &p, int ctrl, int slot)
{
    if (p.trace != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[((size_t)ctrl * gridDim.x + blockIdx.x) * kTraceSlots + slot] = t;
        if (slot == 0) {                                   // where the CTA runs: slot 10 = 1 + %smid, slot 11 = 1 + %warpid of thread 0
            unsigned sm, wid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
            p.trace[((size_t)ctrl * gridDim.x + blockIdx.x) * kTraceSlots + 10] = 1ull + sm;
            p.trace[((size_t)ctrl * gridDim.x + blockIdx.x) * kTraceSlots + 11] = 1ull + wid;
        }
    }
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  Integer contract shared with oracle/mppi_oracle.c.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1, int rounds = 10)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        if (r >= rounds) break;
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Same generator with the ten round keys read from the kernel parameter block (constant bank
// operands of LOP3: no per-call key arithmetic, no registers).  `rounds` (7 or 10) is grid-uniform: rounds 8-10 sit
// behind one uniform branch.
__device__ __forceinline__ uint4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  const uint32_t (&rk0)[10], const uint32_t (&rk1)[10], int rounds = 10)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        if (r == 7 && rounds <= 7) break;
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk0[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk1[r];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float bits_to_1_2(uint32_t x)
{
    return __uint_as_float((x >> 9) | 0x3f800000u);   // [1, 2)
}

// Box-Muller on two 32-bit words -> two standard normals (MUFU lg2 / sqrt / sin / cos).
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float &z0, float &z1)
{
    const float u1 = 2.0f - bits_to_1_2(xa);                       // (0, 1]
    const float th = fmaf(bits_to_1_2(xb), 6.2831853071795865f, -6.2831853071795865f);  // [0, 2pi)
    float r;                                                        // sqrt(-2 ln u1), one MUFU.SQRT
    float l2;                                                       // raw MUFU.LG2: u1 is never denormal
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * l2));
    float sn, cs;
    __sincosf(th, &sn, &cs);
    z0 = r * cs;
    z1 = r * sn;
}

// Box-Muller on the four words of one Philox call: (x0,x1) and (x2,x3), the two pairs processed together with packed
// fp32 instructions (FFMA2 / FMUL2) wherever both need the same op.
__device__ __forceinline__ void normals_from_words(const uint4 x, float z[4])
{
    const float2 fa = make_float2(bits_to_1_2(x.x), bits_to_1_2(x.z));           // radius uniforms of both pairs
    const float2 fb = make_float2(bits_to_1_2(x.y), bits_to_1_2(x.w));           // angle uniforms of both pairs
    const float2 u1 = __ffma2_rn(fa, make_float2(-1.f, -1.f), make_float2(2.f, 2.f));                       // (0, 1]
    const float2 th = __ffma2_rn(fb, make_float2(6.2831853071795865f, 6.2831853071795865f),
                                 make_float2(-6.2831853071795865f, -6.2831853071795865f));                  // [0, 2pi)
    float2 l2;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2.x) : "f"(u1.x));                   // u1 is never denormal
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2.y) : "f"(u1.y));
    const float2 a2 = __fmul2_rn(l2, make_float2(-1.3862943611198906f, -1.3862943611198906f));             // -2 ln u1
    float2 r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(a2.x));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(a2.y));
    float2 csA, csB;                                                              // (cos, sin) of each pair
    __sincosf(th.x, &csA.y, &csA.x);
    __sincosf(th.y, &csB.y, &csB.x);
    const float2 zA = __fmul2_rn(make_float2(r.x, r.x), csA);
    const float2 zB = __fmul2_rn(make_float2(r.y, r.y), csB);
    z[0] = zA.x; z[1] = zA.y; z[2] = zB.x; z[3] = zB.y;
}

// Four standard normals z[4c .. 4c+3] of sample `k` (global index).
__device__ __forceinline__ void normals4(uint32_t call, uint32_t k, uint32_t stream, const RolloutParams &p,
                                         float z[4])
{
    normals_from_words(philox4x32_10_rk(call, k, p.update, stream, p.rk0, p.rk1, p.rounds), z);
}

// ------------------------------------------------------------------------------------------
// The same Philox4x32-10 words with the warp-uniform part of the first rounds taken out of the per-sample stream.
// The counter is (call, sample, update, stream): only the second word differs between samples, so
//   round 1  both products are uniform; the only per-sample value is c0' = A ^ sample, A = hi(M1 update) ^ key0_0
//   round 2  M0 c0' depends on the sample alone: computed ONCE PER SAMPLE (PhiloxSample), not per call;
//            M1 c2' is uniform per call
//   round 3  M0 c0'' is uniform per call;  round 4  one XOR operand is uniform per call
// A per-call table {E, B, C, D} (philox_call_table, staged in shared memory once per CTA) folds the uniform words and
// round keys; per call and sample 15 IMAD.WIDE + 17 LOP3 remain of 20 + 20.  Bit-identical to philox4x32_10_rk
// (checked against the generic generator by every store-then-replay test: mppi_dump_noise uses the generic one).
// ------------------------------------------------------------------------------------------
struct PhiloxSample { uint32_t h0, l0; };

__device__ __forceinline__ uint32_t philox_uniform_A(const RolloutParams &p)
{
    return (uint32_t)(((uint64_t)kPhiloxM1 * p.update) >> 32) ^ p.rk0[0];
}
__device__ __forceinline__ PhiloxSample philox_sample(uint32_t A, uint32_t k)
{
    const uint64_t p0 = (uint64_t)kPhiloxM0 * (A ^ k);
    return PhiloxSample{(uint32_t)(p0 >> 32), (uint32_t)p0};
}
__device__ __forceinline__ uint4 philox_call_table(uint32_t call, uint32_t stream, const RolloutParams &p)
{
    const uint64_t p0 = (uint64_t)kPhiloxM0 * call;                                   // round 1
    const uint32_t c2p = (uint32_t)(p0 >> 32) ^ stream ^ p.rk1[0], c3p = (uint32_t)p0;
    const uint32_t c1p = (uint32_t)((uint64_t)kPhiloxM1 * p.update);
    const uint64_t p1 = (uint64_t)kPhiloxM1 * c2p;                                    // round 2, uniform half
    const uint32_t T0 = (uint32_t)(p1 >> 32) ^ c1p ^ p.rk0[1];
    const uint64_t q0 = (uint64_t)kPhiloxM0 * T0;                                     // round 3, uniform half
    return make_uint4(c3p ^ p.rk1[1], (uint32_t)p1 ^ p.rk0[2], (uint32_t)(q0 >> 32) ^ p.rk1[2], (uint32_t)q0 ^ p.rk1[3]);
}
// R = 7 / 10: compile-time round count (the superposition kernels); R = 0: p.rounds decides behind a uniform branch
template <int R = 0>
__device__ __forceinline__ uint4 philox4x32_10_tab(const uint4 t, const PhiloxSample s, const RolloutParams &p)
{
    uint32_t c0, c1, c2, c3;
    c2 = s.h0 ^ t.x;                                                                  // end of round 2
    {                                                                                 // round 3
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        c0 = (uint32_t)(p1 >> 32) ^ t.y;
        c2 = s.l0 ^ t.z;
        c1 = (uint32_t)p1;
    }
    {                                                                                 // round 4
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        c0 = (uint32_t)(p1 >> 32) ^ c1 ^ p.rk0[3];
        c2 = (uint32_t)(p0 >> 32) ^ t.w;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
    }
#pragma unroll
    for (int r = 4; r < 10; r++) {
        if (R == 0 ? (r == 7 && p.rounds <= 7) : (r >= R)) break;
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ p.rk0[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ p.rk1[r];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ void normals4_tab(const uint4 *tab, uint32_t call, const PhiloxSample s, const RolloutParams &p, float z[4])
{
    normals_from_words(philox4x32_10_tab<0>(tab[call], s, p), z);
}
// The superposition kernels' generator: compile-time rounds, and the normals WITHOUT Box-Muller's sqrt(2 ln 2):
// n = z / kZScale (RolloutParams::z_scale).  One Philox call -> float4 (n0, n1, n2, n3).
constexpr float kZScale = 1.1774100225154747f;      // sqrt(2 ln 2)
__device__ __forceinline__ float4 normals_unscaled_from_words(const uint4 x)
{
    const float2 fa = make_float2(bits_to_1_2(x.x), bits_to_1_2(x.z));
    const float2 fb = make_float2(bits_to_1_2(x.y), bits_to_1_2(x.w));
    const float2 u1 = __ffma2_rn(fa, make_float2(-1.f, -1.f), make_float2(2.f, 2.f));                       // (0, 1]
    const float2 th = __ffma2_rn(fb, make_float2(6.2831853071795865f, 6.2831853071795865f),
                                 make_float2(-6.2831853071795865f, -6.2831853071795865f));                  // [0, 2pi)
    float2 l2, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2.x) : "f"(u1.x));                   // <= 0; u1 is never denormal
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2.y) : "f"(u1.y));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(-l2.x));                  // the negation is an operand modifier
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(-l2.y));
    float2 csA, csB;
    __sincosf(th.x, &csA.y, &csA.x);
    __sincosf(th.y, &csB.y, &csB.x);
    const float2 zA = __fmul2_rn(make_float2(r.x, r.x), csA);
    const float2 zB = __fmul2_rn(make_float2(r.y, r.y), csB);
    return make_float4(zA.x, zA.y, zB.x, zB.y);
}
template <int R>
__device__ __forceinline__ float4 normals4_fast(const uint4 *tab, uint32_t call, const PhiloxSample s, const RolloutParams &p)
{
    return normals_unscaled_from_words(philox4x32_10_tab<R>(tab[call], s, p));
}
// The standard normals a handle's update kernel saw, from the plain generator (mppi_dump_noise, store-then-replay):
// on the superposition path z = z_scale * n, elsewhere the Box-Muller of normals_from_words.
__device__ __forceinline__ void normals4_as_used(uint32_t call, uint32_t k, uint32_t stream, const RolloutParams &p, float z[4])
{
    const uint4 x = philox4x32_10_rk(call, k, p.update, stream, p.rk0, p.rk1, p.rounds);
    if (p.fast) {
        const float4 n = normals_unscaled_from_words(x);
        z[0] = p.z_scale * n.x; z[1] = p.z_scale * n.y; z[2] = p.z_scale * n.z; z[3] = p.z_scale * n.w;
    } else {
        normals_from_words(x, z);
    }
}

// ------------------------------------------------------------------------------------------
// Warp / block reductions
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reduce-scatter of 32 per-lane values across the warp: on return v[0] of lane l holds
// sum over lanes of the input v[l].  31 shuffles for 32 sums.
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane)
{
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; i++) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

// exp(-(S - beta)/lambda) through one MUFU.EX2 with the max-shift already applied (flush-to-zero).  Used for the
// rescaling of partial sums (CTA / warp / rank records), where nothing is cut.
__device__ __forceinline__ float weight_exp(float S, float beta, float neg_inv_lambda_log2e)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((S - beta) * neg_inv_lambda_log2e));
    return r;
}
// The weight of ONE SAMPLE relative to the best sample seen so far by its CTA (or warp).  Weights below 2^-50 of that
// best one are set to exactly 0 and the kernels skip zero-weight samples in the weighted noise sum: over K <= 2^20
// samples the dropped mass is below 2^-30 = 9.3e-10 of eta - under one fp32 ulp (6e-8) of the sums it would enter, where
// such terms are rounded away one by one anyway.  (Round 1 cut at fp32 underflow, 2^-126; the samples between the two
// cuts are the bulk of what the weighted sum used to revisit: 22 % -> 7 % of K at config 3's inputs.)
constexpr float kWeightCutLog2 = 50.0f;
__device__ __forceinline__ float sample_weight(float S, float beta, float neg_inv_lambda_log2e)
{
    const float a = (S - beta) * neg_inv_lambda_log2e;      // <= 0 (up to rounding), log2 of the weight
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return a < -kWeightCutLog2 ? 0.f : r;
}

// ------------------------------------------------------------------------------------------
// Point-mass model + quadratic cost, one sample in registers.
//   x' = A x + (B/m) u  (src/model_base.cpp:53-82), block structure exploited per axis:
//   p' = (p + dt v) + (dt^2/2m) u ;  v' = v + (dt/m) u
//   q(x) = sum_i q_i (x_i - g_i)^2   (src/cost_base.cpp:56-61, Q = Diag(q))
// Axes are processed two at a time with Blackwell's packed fp32 instructions (FFMA2 / FADD2:
// __ffma2_rn, __fadd2_rn); an odd last axis stays scalar.
// ------------------------------------------------------------------------------------------
template <int A>
struct Vec {                       // A floats as A/2 aligned pairs + optional scalar
    static constexpr int NP = A / 2;
    static constexpr bool ODD = (A & 1) != 0;
    float2 pr[NP > 0 ? NP : 1];
    float sc;
    __device__ __forceinline__ float get(int j) const { return (j < 2 * NP) ? ((j & 1) ? pr[j >> 1].y : pr[j >> 1].x) : sc; }
    __device__ __forceinline__ void set(int j, float v)
    {
        if (j < 2 * NP) { if (j & 1) pr[j >> 1].y = v; else pr[j >> 1].x = v; } else sc = v;
    }
    __device__ __forceinline__ void fill(float v)
    {
#pragma unroll
        for (int i = 0; i < NP; i++) pr[i] = make_float2(v, v);
        sc = v;
    }
};

template <int A>
struct ModelConsts {               // everything the step needs, as pairs, built once per thread
    float2 dt2, cpu2, cvu2;
    float dt, cpu, cvu;
    Vec<A> sqp, sqv, sgp, sgv;     // sqrt(q) and sqrt(q)*g for positions / velocities
    __device__ __forceinline__ void init(float dt_, float cpu_, float cvu_, const float *sqrt_q, const float *goal)
    {
        dt = dt_; cpu = cpu_; cvu = cvu_;
        dt2 = make_float2(dt_, dt_); cpu2 = make_float2(cpu_, cpu_); cvu2 = make_float2(cvu_, cvu_);
#pragma unroll
        for (int j = 0; j < A; j++) {
            sqp.set(j, sqrt_q[2 * j]);
            sqv.set(j, sqrt_q[2 * j + 1]);
            sgp.set(j, -sqrt_q[2 * j] * goal[2 * j]);          // stored negated: d = sq*x + (-sq*g)
            sgv.set(j, -sqrt_q[2 * j + 1] * goal[2 * j + 1]);
        }
    }
};

template <int A>
struct PointMass {
    static constexpr int NP = A / 2;
    static constexpr bool ODD = (A & 1) != 0;
    Vec<A> p, v;
    __device__ __forceinline__ void init(const float *x0)
    {
#pragma unroll
        for (int j = 0; j < A; j++) { p.set(j, x0[2 * j]); v.set(j, x0[2 * j + 1]); }
    }
    __device__ __forceinline__ void zero() { p.fill(0.f); v.fill(0.f); }
    __device__ __forceinline__ void step(const Vec<A> &u, const ModelConsts<A> &c)
    {
#pragma unroll
        for (int i = 0; i < NP; i++) {
            p.pr[i] = __ffma2_rn(c.cpu2, u.pr[i], __ffma2_rn(c.dt2, v.pr[i], p.pr[i]));
            v.pr[i] = __ffma2_rn(c.cvu2, u.pr[i], v.pr[i]);
        }
        if (ODD) {
            p.sc = fmaf(c.cpu, u.sc, fmaf(c.dt, v.sc, p.sc));
            v.sc = fmaf(c.cvu, u.sc, v.sc);
        }
    }
    // adds q(x) into the running (pair, scalar) accumulators
    __device__ __forceinline__ void state_cost(const ModelConsts<A> &c, float2 &acc2, float &acc) const
    {
#pragma unroll
        for (int i = 0; i < NP; i++) {
            const float2 dp = __ffma2_rn(c.sqp.pr[i], p.pr[i], c.sgp.pr[i]);
            const float2 dv = __ffma2_rn(c.sqv.pr[i], v.pr[i], c.sgv.pr[i]);
            acc2 = __ffma2_rn(dp, dp, acc2);
            acc2 = __ffma2_rn(dv, dv, acc2);
        }
        if (ODD) {
            const float dp = fmaf(c.sqp.sc, p.sc, c.sgp.sc), dv = fmaf(c.sqv.sc, v.sc, c.sgv.sc);
            acc = fmaf(dp, dp, acc);
            acc = fmaf(dv, dv, acc);
        }
    }
};

// ElipseCost.state_cost (scripts/src/costs/elipse_cost.py:46-79), state (x, vx, y, vy):
//   m_state |((x-cx)/a)^2 + ((y-cy)/b)^2 - 1| + m_vel (sqrt(vx^2 + vy^2) - speed)^2
__device__ __forceinline__ float ellipse_cost(const float *ell, float x, float vx, float y, float vy)
{
    const float dx = (x - ell[2]) * ell[0], dy = (y - ell[3]) * ell[1];
    const float d = fabsf(fmaf(dx, dx, dy * dy) - 1.0f);
    const float dv = __fsqrt_rn(fmaf(vx, vx, vy * vy)) - ell[4];
    return fmaf(ell[5], d, ell[6] * dv * dv);
}

// Per-sample cost accumulator: pair lanes + scalar, folded once at the end.
struct CostAcc {
    float2 a2;
    float a;
    __device__ __forceinline__ void zero() { a2 = make_float2(0.f, 0.f); a = 0.f; }
    __device__ __forceinline__ float total() const { return (a2.x + a2.y) + a; }
};

// Compensated (Kahan) running sum for the per-sample cost of the direct-form kernels: the steps of a block are summed from
// zero (small values, small roundings) and the block sums are added with the lost low-order part carried along, so the
// total is as good as its final fp32 rounding instead of collecting one rounding at ulp(S) per term.
struct KahanSum {
    float s, c;
    __device__ __forceinline__ void init(float v) { s = v; c = 0.f; }
    __device__ __forceinline__ void add(float v)
    {
        const float y = v - c;
        const float t = s + y;
        c = (t - s) - y;
        s = t;
    }
};

// q(x) of the selected state-cost functor, added into S
template <int A, int COST>
__device__ __forceinline__ void add_state_cost(const PointMass<A> &x, const ModelConsts<A> &mc, const RolloutParams &p, CostAcc &S)
{
    if (COST == 1) {
        S.a += ellipse_cost(p.ell, x.p.get(0), x.v.get(0), x.p.get(A > 1 ? 1 : 0), x.v.get(A > 1 ? 1 : 0));
    } else {
        x.state_cost(mc, S.a2, S.a);
    }
}

// Thread-block clusters: rank / size of this CTA's cluster (1 when the kernel was launched without a cluster dimension),
// the cluster barrier, and a load from the same shared-memory address of another CTA of the cluster (DSMEM).
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem(const float *local_smem_ptr, uint32_t rank)
{
    uint32_t remote;
    float v;
    // not volatile: the loads of a reduction are independent and must be free to overlap (they are fenced by cluster_sync)
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((uint32_t)__cvta_generic_to_shared(local_smem_ptr)), "r"(rank));
    asm("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote));
    return v;
}

// mbarrier / bulk-copy (TMA) helpers -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// Bounded: a bulk copy that never completes (a bad address, a byte count that does not match expect_tx) would otherwise
// spin for ever and take the GPU with it; after about 2^26 polls (seconds) the kernel traps and the launch reports an error.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 26); spin++) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    asm volatile("trap;");
}
// The tight loop, for hand-overs between warps of one CTA (tcgen05.commit arrivals, operand-ready barriers of the MLP
// kernel): nothing outside the CTA can keep such a barrier from completing, and the counter and compare of the bounded loop
// sit on that kernel's latency chain (its MMA and helper warps run on 40 / 56 registers: config 4 0.478 -> 0.589 ms, 4.1 M
// local loads instead of 0.23 M, when every wait was bounded).
__device__ __forceinline__ void mbar_wait_spin(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (UBLKCP in SASS).
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

}  // namespace mppi
