// Superposition form of the point-mass rollout (diagonal Sigma, q > 0, StaticCost).
//
// The model is linear (src/model_base.cpp:53-82: x' = A x + (B/m) u) and the state cost quadratic
// (src/cost_base.cpp:56-61), so with u_t = U_t + eps_t the state of a sample splits exactly into
//   x_t = xd_t + xn_t      xd: the noise-free trajectory under U (the same for every sample of the controller)
//                          xn: driven by the noise alone, xn_0 = 0
// and, per axis, with D = sqrt(q) (xd - g), P = sqrt(q_p) pn, V = sqrt(q_v) vn, w_t = 1 (t < T), w_T = 2 (the terminal
// cost comes on top of step T-1's, src/controller_base.cpp:271-272):
//   S = sum_t w_t |D_t + (P_t, V_t)|^2 + action cost
//     = C + sum_t w_t (P_t^2 + V_t^2) + sum_tau L_tau . n_tau
//   C     = sum_t w_t |D_t|^2 + C0                                   per controller, computed once per CTA
//   L_tau = G_tau + w_scale U_tau                                    (the action cost lambda U^T Sigma^-1 eps is linear in n too)
//   G_tau = 2 sum_{t > tau} w_t [D_p,t (b1 + a1 b2 (t-1-tau)) + D_v,t b2]     the adjoint of the noise response
//   P' = P + a1 V + b1 n ,  V' = V + b2 n        a1 = dt sqrt(q_p/q_v), b1 = sqrt(q_p) (dt^2/2m) sigma_j s, b2 = sqrt(q_v) (dt/m) sigma_j s
// n is the generator's output (Philox mode: n = z / s with s = z_scale; injected mode: n = eps, sigma_j s := 1).
// Six FMAs per axis-step remain of nine (u = U + sigma z, U.z, p, v, two deviations, two squares), and U_t itself is
// never needed per sample.  Exact algebra: checked against the op-for-op oracle in fp64 to 2e-15 (scripts_dev/superposition_check.py).
#pragma once
#include "mppi_device.cuh"

namespace mppi {

constexpr int kFastMaxT = 1024;    // the tables are built by 2T dependent steps per axis; longer horizons take the direct form

template <int A>
struct FastConsts {
    Vec<A> a1, b1, b2;
    __device__ __forceinline__ void init(const RolloutParams &p)
    {
#pragma unroll
        for (int j = 0; j < A; j++) { a1.set(j, p.fa1[j]); b1.set(j, p.fb1[j]); b2.set(j, p.fb2[j]); }
    }
};

// One step of the noise-driven part for one sample.  `l_row` = L_t (shared memory, 16-byte aligned row of RowU<A>::RS floats).
template <int A>
__device__ __forceinline__ void fast_step(Vec<A> &P, Vec<A> &V, CostAcc &Sq, CostAcc &Sl, const float *l_row, const Vec<A> &n,
                                          const FastConsts<A> &fc)
{
    constexpr int NP = A / 2, RS = (A + 3) & ~3;
    constexpr bool ODD = (A & 1) != 0;
    float lv[RS];
#pragma unroll
    for (int i = 0; i < RS / 4; i++) {
        const float4 v = reinterpret_cast<const float4 *>(l_row)[i];
        lv[4 * i] = v.x; lv[4 * i + 1] = v.y; lv[4 * i + 2] = v.z; lv[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < NP; i++) {
        P.pr[i] = __ffma2_rn(fc.b1.pr[i], n.pr[i], __ffma2_rn(fc.a1.pr[i], V.pr[i], P.pr[i]));
        V.pr[i] = __ffma2_rn(fc.b2.pr[i], n.pr[i], V.pr[i]);
        Sq.a2 = __ffma2_rn(P.pr[i], P.pr[i], Sq.a2);
        Sl.a2 = __ffma2_rn(make_float2(lv[2 * i], lv[2 * i + 1]), n.pr[i], Sl.a2);
        Sq.a2 = __ffma2_rn(V.pr[i], V.pr[i], Sq.a2);
    }
    if (ODD) {
        P.sc = fmaf(fc.b1.sc, n.sc, fmaf(fc.a1.sc, V.sc, P.sc));
        V.sc = fmaf(fc.b2.sc, n.sc, V.sc);
        Sq.a = fmaf(P.sc, P.sc, Sq.a);
        Sl.a = fmaf(lv[A - 1], n.sc, Sl.a);
        Sq.a = fmaf(V.sc, V.sc, Sq.a);
    }
}
// the terminal cost: |(P_T, V_T)|^2 once more
template <int A>
__device__ __forceinline__ void fast_terminal(const Vec<A> &P, const Vec<A> &V, CostAcc &Sq)
{
#pragma unroll
    for (int i = 0; i < A / 2; i++) {
        Sq.a2 = __ffma2_rn(P.pr[i], P.pr[i], Sq.a2);
        Sq.a2 = __ffma2_rn(V.pr[i], V.pr[i], Sq.a2);
    }
    if (A & 1) {
        Sq.a = fmaf(P.sc, P.sc, Sq.a);
        Sq.a = fmaf(V.sc, V.sc, Sq.a);
    }
}

// Builds the per-controller tables in shared memory; every thread of the CTA must call it (contains CTA barriers).
//   sL   [T][RS]   L_tau (zero padded), RS = RowU<A>::RS
//   sD   scratch, >= (3 A + 1) T floats: U staged [T][A], then D_p / D_v of steps 1..T [T][2A], then T partial sums
//   sRed scratch, >= 1 float
// Philox mode: L = G + w_scale z_scale U (n = z / z_scale); injected mode: L = G + lambda Sigma^-T U (n = eps).
// The forward (noise-free trajectory) and backward (adjoint) passes are the model's own recurrences, one thread per
// axis — 2T dependent steps of two or three FMAs, about a microsecond at T = 100 — and the sum C = sum_t w_t |D_t|^2
// is taken over per-step partials in a fixed order.  Returns C (without C0), CTA-uniform.
template <int A, bool PHILOX>
__device__ __forceinline__ float build_linear_tables(const RolloutParams &p, int ctrl, float *sL, float *sD, float *sRed)
{
    constexpr int RS = (A + 3) & ~3;
    const int T = p.T, tid = threadIdx.x, nthr = blockDim.x;
    const float *U = p.U + (size_t)ctrl * p.TA;
    float *sU = sD;                                   // [T][A]
    float *sDev = sD + A * T;                         // [T][2A]: D_p[A], D_v[A] of step t + 1
    float *sC = sDev + 2 * A * T;                     // [T]
    // goal and state of this thread's axis: loaded before the barrier, together with the sequence (one L2 round trip)
    float gpj = 0.f, gvj = 0.f, pj = 0.f, vj = 0.f;
    if (tid < A) {
        const float *gp = p.goal + (p.goal_per_ctrl ? (size_t)ctrl * 2 * A : 0);
        const float *xp = p.x + (size_t)ctrl * 2 * A;
        gpj = gp[2 * tid];
        gvj = gp[2 * tid + 1];
        pj = p.x_inline ? p.x0[2 * tid] : xp[2 * tid];
        vj = p.x_inline ? p.x0[2 * tid + 1] : xp[2 * tid + 1];
    }
    for (int i = tid; i < T * A; i += nthr) sU[i] = U[i];
    for (int i = tid; i < T * RS; i += nthr) sL[i] = 0.f;
    __syncthreads();
    if (tid < A) {
        const int j = tid;
        const float sqp = p.sqrt_q[2 * j], sqv = p.sqrt_q[2 * j + 1];
        // forward: x' = A x + (B/m) U_t, exactly the model step (src/model_base.cpp:53-82); eight steps at a time so that the
        // shared-memory loads of a chunk are in flight together (the recurrence itself is two dependent FMAs per step)
        constexpr int CH = 8;
        for (int t0 = 0; t0 < T; t0 += CH) {
            float u[CH];
#pragma unroll
            for (int i = 0; i < CH; i++) u[i] = (t0 + i < T) ? sU[(t0 + i) * A + j] : 0.f;
            float dp[CH], dv[CH];
#pragma unroll
            for (int i = 0; i < CH; i++) {
                pj = fmaf(p.c_pu, u[i], fmaf(p.dt, vj, pj));
                vj = fmaf(p.c_vu, u[i], vj);
                dp[i] = sqp * (pj - gpj);
                dv[i] = sqv * (vj - gvj);
            }
#pragma unroll
            for (int i = 0; i < CH; i++)
                if (t0 + i < T) {
                    sDev[2 * (t0 + i) * A + j] = dp[i];
                    sDev[(2 * (t0 + i) + 1) * A + j] = dv[i];
                }
        }
        // backward: aP_t = g_P,t + aP_{t+1}; aV_t = g_V,t + aV_{t+1} + a1 aP_{t+1}; g = 2 w_t D_t; G_tau = b1 aP_{tau+1} + b2 aV_{tau+1}
        const float a1 = p.fa1[j], b1 = p.fb1[j], b2 = p.fb2[j];
        float aP = 0.f, aV = 0.f;
        for (int t0 = T; t0 >= 1; t0 -= CH) {                      // steps t0, t0-1, .. t0-CH+1
            float dp[CH], dv[CH], lin[CH];
#pragma unroll
            for (int i = 0; i < CH; i++) {
                const int t = t0 - i;
                dp[i] = (t >= 1) ? sDev[2 * (t - 1) * A + j] : 0.f;
                dv[i] = (t >= 1) ? sDev[(2 * (t - 1) + 1) * A + j] : 0.f;
                lin[i] = 0.f;
                if (t >= 1) {
                    if (PHILOX) {
                        lin[i] = p.w_scale * p.z_scale * sU[(t - 1) * A + j];
                    } else {
#pragma unroll
                        for (int l = 0; l < A; l++) lin[i] = fmaf(p.lam_inv_sigma_T[j * A + l], sU[(t - 1) * A + l], lin[i]);
                    }
                }
            }
            float Lv[CH];
#pragma unroll
            for (int i = 0; i < CH; i++) {
                const int t = t0 - i;
                const float w2 = (t == T) ? 4.f : 2.f;
                const float aPn = fmaf(w2, dp[i], aP);
                aV = fmaf(a1, aP, fmaf(w2, dv[i], aV));
                aP = aPn;
                Lv[i] = fmaf(b1, aP, fmaf(b2, aV, lin[i]));
            }
#pragma unroll
            for (int i = 0; i < CH; i++)
                if (t0 - i >= 1) sL[(t0 - i - 1) * RS + j] = Lv[i];
        }
    }
    __syncthreads();
    for (int t = tid; t < T; t += nthr) {
        float c = 0.f;
#pragma unroll
        for (int j = 0; j < 2 * A; j++) { const float d = sDev[2 * t * A + j]; c = fmaf(d, d, c); }
        sC[t] = (t == T - 1) ? 2.f * c : c;
    }
    __syncthreads();
    if (tid < 32) {                                   // warp 0: lane-strided partials, then a fixed shuffle tree
        float c = 0.f;
        for (int t = tid; t < T; t += 32) c += sC[t];
        c = warp_sum(c);
        if (tid == 0) sRed[0] = c;
    }
    __syncthreads();
    const float C = sRed[0];
    __syncthreads();
    return C;
}

}  // namespace mppi
