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
// The forward (noise-free trajectory) and backward (adjoint) passes are the model's own recurrences, one warp per axis
// as a blocked scan, and the sum C = sum_t w_t |D_t|^2 is taken over per-step partials in a fixed order.
// Returns C (without C0), CTA-uniform.
template <int A, bool PHILOX>
__device__ __forceinline__ float build_linear_tables(const RolloutParams &p, int ctrl, float *sL, float *sD, float *sRed)
{
    constexpr int RS = (A + 3) & ~3;
    const int T = p.T, tid = threadIdx.x, nthr = blockDim.x;
    const float *U = p.U + (size_t)ctrl * p.TA;
    float *sU = sD;                                   // [T][A]
    float *sDev = sD + A * T;                         // [T][2A]: D_p[A], D_v[A] of step t + 1
    float *sC = sDev + 2 * A * T;                     // [T]
    // Both passes are the model's own recurrences (src/model_base.cpp:53-82 forward, its adjoint backward), one warp per
    // axis, as a blocked scan: lane l owns the contiguous chunk of L = ceil(T/32) steps, runs it from a zero state, the 32
    // chunk responses are composed by a warp scan of the affine maps (x -> M^n x + b with M^n = [[1, n dt], [0, 1]] forward,
    // [[1, 0], [n a1, 1]] for the adjoint) and the lane replays its chunk from its true incoming state: 2 L + 5 dependent
    // steps per pass instead of T (T = 100: 13).
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5, L = (T + 31) >> 5;
    // goal and state of the warp's first axis: loaded before the barrier, together with the sequence (one L2 round trip)
    const float *gp = p.goal + (p.goal_per_ctrl ? (size_t)ctrl * 2 * A : 0);
    const float *xp = p.x + (size_t)ctrl * 2 * A;
    float gp0 = 0.f, gv0 = 0.f, p0 = 0.f, v0 = 0.f;
    if (warp < A) {
        gp0 = gp[2 * warp];
        gv0 = gp[2 * warp + 1];
        p0 = p.x_inline ? p.x0[2 * warp] : xp[2 * warp];
        v0 = p.x_inline ? p.x0[2 * warp + 1] : xp[2 * warp + 1];
    }
    for (int i = tid; i < T * A; i += nthr) sU[i] = U[i];
    for (int i = tid; i < T * RS; i += nthr) sL[i] = 0.f;
    __syncthreads();
    for (int j = warp; j < A; j += nwarp) {                            // warp-uniform
        if (j != warp) {
            gp0 = gp[2 * j];
            gv0 = gp[2 * j + 1];
            p0 = p.x_inline ? p.x0[2 * j] : xp[2 * j];
            v0 = p.x_inline ? p.x0[2 * j + 1] : xp[2 * j + 1];
        }
        const float sqp = p.sqrt_q[2 * j], sqv = p.sqrt_q[2 * j + 1];
        const int t_lo = min(T, lane * L), t_hi = min(T, t_lo + L);      // this lane's steps [t_lo, t_hi)
        // ---- forward: x' = A x + (B/m) U_t ------------------------------------------------------------------
        float bp = 0.f, bv = 0.f;
        for (int t = t_lo; t < t_hi; t++) {
            const float u = sU[t * A + j];
            bp = fmaf(p.c_pu, u, fmaf(p.dt, bv, bp));
            bv = fmaf(p.c_vu, u, bv);
        }
        float sp = bp, sv = bv;                                         // inclusive scan: response of chunks 0 .. lane
        int sn = t_hi - t_lo;                                           // ... and their step count
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float lp = __shfl_up_sync(0xffffffffu, sp, o), lv = __shfl_up_sync(0xffffffffu, sv, o);
            const int ln = __shfl_up_sync(0xffffffffu, sn, o);
            if (lane >= o) {                                            // left segment first, then (sp, sv, sn)
                sp = sp + fmaf((float)sn * p.dt, lv, lp);
                sv = sv + lv;
                sn += ln;
            }
        }
        {
            float ip = __shfl_up_sync(0xffffffffu, sp, 1), iv = __shfl_up_sync(0xffffffffu, sv, 1);
            if (lane == 0) { ip = 0.f; iv = 0.f; }
            float pj = p0 + fmaf((float)t_lo * p.dt, v0, ip), vj = v0 + iv;     // state before step t_lo
            for (int t = t_lo; t < t_hi; t++) {
                const float u = sU[t * A + j];
                pj = fmaf(p.c_pu, u, fmaf(p.dt, vj, pj));
                vj = fmaf(p.c_vu, u, vj);
                sDev[2 * t * A + j] = sqp * (pj - gp0);
                sDev[(2 * t + 1) * A + j] = sqv * (vj - gv0);
            }
        }
        __syncwarp();
        // ---- backward (adjoint): aP_t = g_P,t + aP_{t+1}; aV_t = g_V,t + aV_{t+1} + a1 aP_{t+1}; g = 2 w_t D_t;
        //      L_tau = b1 aP_{tau+1} + b2 aV_{tau+1} + (action-cost term).  Lane l walks steps t = T - l L down. ----------
        const float a1 = p.fa1[j], b1 = p.fb1[j], b2 = p.fb2[j];
        const int u_hi = max(0, T - lane * L), u_lo = max(0, u_hi - L);  // this lane's steps t in (u_lo, u_hi], downwards
        float cP = 0.f, cV = 0.f;
        for (int t = u_hi; t > u_lo; t--) {
            const float w2 = (t == T) ? 4.f : 2.f;
            const float dp = sDev[2 * (t - 1) * A + j], dv = sDev[(2 * (t - 1) + 1) * A + j];
            cV = fmaf(a1, cP, fmaf(w2, dv, cV));
            cP = fmaf(w2, dp, cP);
        }
        float aPs = cP, aVs = cV;
        int an = u_hi - u_lo;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float lP = __shfl_up_sync(0xffffffffu, aPs, o), lV = __shfl_up_sync(0xffffffffu, aVs, o);
            const int ln = __shfl_up_sync(0xffffffffu, an, o);
            if (lane >= o) {                                            // earlier (higher-t) segment first
                aVs = aVs + fmaf((float)an * a1, lP, lV);
                aPs = aPs + lP;
                an += ln;
            }
        }
        {
            float aP = __shfl_up_sync(0xffffffffu, aPs, 1), aV = __shfl_up_sync(0xffffffffu, aVs, 1);
            if (lane == 0) { aP = 0.f; aV = 0.f; }
            for (int t = u_hi; t > u_lo; t--) {
                const float w2 = (t == T) ? 4.f : 2.f;
                const float dp = sDev[2 * (t - 1) * A + j], dv = sDev[(2 * (t - 1) + 1) * A + j];
                float lin = 0.f;
                if (PHILOX) {
                    lin = p.w_scale * p.z_scale * sU[(t - 1) * A + j];
                } else {
#pragma unroll
                    for (int l = 0; l < A; l++) lin = fmaf(p.lam_inv_sigma_T[j * A + l], sU[(t - 1) * A + l], lin);
                }
                aV = fmaf(a1, aP, fmaf(w2, dv, aV));
                aP = fmaf(w2, dp, aP);
                sL[(t - 1) * RS + j] = fmaf(b1, aP, fmaf(b2, aV, lin));
            }
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
