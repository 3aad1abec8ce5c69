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

constexpr int kFastMaxT = 256;     // the tables are built with O(T) work per entry; longer horizons take the direct form

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
//   sD   scratch, >= 2 * A * T floats (D_p, D_v of steps 1..T), free afterwards
//   sRed scratch, >= 1 float
// `lin_scale`: w_scale * z_scale (Philox mode, n = z / z_scale) or 1 with `lin_from_w` set (injected mode: the action-cost
// vector is lambda Sigma^-T U_t, computed here from p.lam_inv_sigma_T).  Returns C (without C0), CTA-uniform, fixed order.
template <int A, bool PHILOX>
__device__ __forceinline__ float build_linear_tables(const RolloutParams &p, int ctrl, float *sL, float *sD, float *sRed)
{
    constexpr int RS = (A + 3) & ~3;
    const int T = p.T, tid = threadIdx.x, nthr = blockDim.x;
    const float *U = p.U + (size_t)ctrl * p.TA;
    const float *gp = p.goal + (p.goal_per_ctrl ? (size_t)ctrl * 2 * A : 0);
    const float *xp = p.x + (size_t)ctrl * 2 * A;
    const float dtcvu = p.dt * p.c_vu;
    // forward: noise-free deviations at t = 1..T, one (t, axis) item per thread, closed form
    //   v_t = v_0 + c_vu sum_{tau<t} U_tau ;  p_t = p_0 + t dt v_0 + sum_{tau<t} [c_pu + dt c_vu (t-1-tau)] U_tau
    for (int it = tid; it < T * A; it += nthr) {
        const int t = it / A + 1, j = it - (t - 1) * A;
        float su = 0.f, sp = 0.f;
        for (int tau = 0; tau < t; tau++) {
            const float u = U[tau * A + j];
            su += u;
            sp = fmaf(fmaf(dtcvu, (float)(t - 1 - tau), p.c_pu), u, sp);
        }
        const float p0 = p.x_inline ? p.x0[2 * j] : xp[2 * j], v0 = p.x_inline ? p.x0[2 * j + 1] : xp[2 * j + 1];
        const float pt = fmaf((float)t * p.dt, v0, p0) + sp;
        const float vt = fmaf(p.c_vu, su, v0);
        sD[(2 * (t - 1)) * A + j] = p.sqrt_q[2 * j] * (pt - gp[2 * j]);
        sD[(2 * (t - 1) + 1) * A + j] = p.sqrt_q[2 * j + 1] * (vt - gp[2 * j + 1]);
    }
    __syncthreads();
    // backward: G_tau = 2 sum_{t>tau} w_t [D_p,t (b1 + a1 b2 (t-1-tau)) + D_v,t b2]
    for (int it = tid; it < T * RS; it += nthr) {
        const int tau = it / RS, j = it - tau * RS;
        float L = 0.f;
        if (j < A) {
            const float a1b2 = p.fa1[j] * p.fb2[j];
            float g = 0.f;
            for (int t = tau + 1; t <= T; t++) {
                const float w = (t == T) ? 4.f : 2.f;
                const float dp = sD[(2 * (t - 1)) * A + j], dv = sD[(2 * (t - 1) + 1) * A + j];
                g = fmaf(w, fmaf(dp, fmaf(a1b2, (float)(t - 1 - tau), p.fb1[j]), dv * p.fb2[j]), g);
            }
            float lin;
            if (PHILOX) {
                lin = p.w_scale * p.z_scale * U[tau * A + j];
            } else {
                lin = 0.f;
#pragma unroll
                for (int l = 0; l < A; l++) lin = fmaf(p.lam_inv_sigma_T[j * A + l], U[tau * A + l], lin);
            }
            L = g + lin;
        }
        sL[it] = L;
    }
    // C = sum_t w_t |D_t|^2: per-step partials, summed in a fixed order by warp 0
    __syncthreads();                                 // everyone is done reading... (sD is read again below: no writes yet)
    float *sC = sD + 2 * A * T;                      // [T] partials (the caller sizes sD as 2*A*T + T)
    for (int t = tid; t < T; t += nthr) {
        float c = 0.f;
#pragma unroll
        for (int j = 0; j < 2 * A; j++) { const float d = sD[2 * t * A + j]; c = fmaf(d, d, c); }
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
