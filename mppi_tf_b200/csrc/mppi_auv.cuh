// AUV (Fossen) dynamics and the quaternion goal cost as device functors.
// Behaviour: /root/reference/scripts/src/models/auv_model.py (state_dot :308-333, body2inertial_transform
// :353-398, damping_matrix :478-506, coriolis_matrix :508-542, restoring_forces :450-476, acc :544-559,
// step :285-306, normalize_quat :426-448) and scripts/src/costs/static_cost.py (StaticCost :40-63,
// StaticQuatCost :116-159).  State x = (p[3], q = (qx, qy, qz, qw), nu[6]), action u[6].
#pragma once
#include "mppi_device.cuh"

namespace mppi {

constexpr int kAuvS = 13, kAuvA = 6;

// Derived parameters, built on the host (mppi_set_auv_model): 160 floats, staged into shared memory per CTA.
struct AuvParams {
    float fng, fnb;            // -m g and V rho g: z components of gravity / buoyancy in the inertial frame
    float cog[3], cob[3];
    float Mtot[36], invM[36];  // rigid body + added mass (with the reference's transposed skew of cog), inverse
    float Dl[36], dq[6], Dlf[36];
    float dt;
    int rk;                    // 1, 2 or 4; 0 = the learned model below (NNAUVModel) instead of the Fossen equations
    const float *nn_blob;      // device, nn_auv_blob_floats(nn_hidden) floats (mppi_set_nn_auv_model)
    int nn_hidden;             // hidden layers (1..kNnMaxHidden), each kNnW wide on the device (narrower ones zero padded)
};
constexpr int kAuvParamWords = sizeof(AuvParams) / 4;

__device__ __forceinline__ void auv_state_dot(const AuvParams &P, const float (&x)[kAuvS], const float (&u)[kAuvA], float (&xd)[kAuvS])
{
    const float qx = x[3], qy = x[4], qz = x[5], qw = x[6];
    const float *nu = &x[7];
    const float R[9] = {1.f - 2.f * (qy * qy + qz * qz), 2.f * (qx * qy - qz * qw), 2.f * (qx * qz + qy * qw),
                        2.f * (qx * qy + qz * qw), 1.f - 2.f * (qx * qx + qz * qz), 2.f * (qy * qz - qx * qw),
                        2.f * (qx * qz - qy * qw), 2.f * (qy * qz + qx * qw), 1.f - 2.f * (qx * qx + qy * qy)};
#pragma unroll
    for (int r = 0; r < 3; r++) xd[r] = fmaf(R[3 * r + 2], nu[2], fmaf(R[3 * r + 1], nu[1], R[3 * r] * nu[0]));
    xd[3] = 0.5f * (qw * nu[3] - qz * nu[4] + qy * nu[5]);
    xd[4] = 0.5f * (qz * nu[3] + qw * nu[4] - qx * nu[5]);
    xd[5] = 0.5f * (-qy * nu[3] + qx * nu[4] + qw * nu[5]);
    xd[6] = 0.5f * (-qx * nu[3] - qy * nu[4] - qz * nu[5]);
    float rhs[6];
    // rhs = u - D nu, D = -Dl - nu0 Dlf - diag(dq |nu|)
#pragma unroll
    for (int r = 0; r < 6; r++) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 6; c++) {
            float d = -P.Dl[6 * r + c] - nu[0] * P.Dlf[6 * r + c];
            if (r == c) d -= P.dq[r] * fabsf(nu[r]);
            acc = fmaf(d, nu[c], acc);
        }
        rhs[r] = u[r] - acc;
    }
    // - C nu, C = [[0, S12], [S12, S22]], S12 = -skew(M11 nu1 + M12 nu2), S22 = -skew(M21 nu1 + M22 nu2)
    float a1[3], a2[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < 6; c++) {
            s1 = fmaf(P.Mtot[6 * r + c], nu[c], s1);
            s2 = fmaf(P.Mtot[6 * (3 + r) + c], nu[c], s2);
        }
        a1[r] = s1;
        a2[r] = s2;
    }
    // -skew(a) v = v x a ... written out: (-skew(a) v)_0 = a2 v1 - a1 v2, _1 = -a2 v0 + a0 v2, _2 = a1 v0 - a0 v1
    const float c0 = a1[2] * nu[4] - a1[1] * nu[5], c1 = -a1[2] * nu[3] + a1[0] * nu[5], c2 = a1[1] * nu[3] - a1[0] * nu[4];
    const float d0 = (a1[2] * nu[1] - a1[1] * nu[2]) + (a2[2] * nu[4] - a2[1] * nu[5]);
    const float d1 = (-a1[2] * nu[0] + a1[0] * nu[2]) + (-a2[2] * nu[3] + a2[0] * nu[5]);
    const float d2 = (a1[1] * nu[0] - a1[0] * nu[1]) + (a2[1] * nu[3] - a2[0] * nu[4]);
    rhs[0] -= c0; rhs[1] -= c1; rhs[2] -= c2; rhs[3] -= d0; rhs[4] -= d1; rhs[5] -= d2;
    // - g, g = -[fbg + fbb; cog x fbg + cob x fbb], fbg = R^T (0, 0, fng), fbb = R^T (0, 0, fnb)
    const float fbg[3] = {R[6] * P.fng, R[7] * P.fng, R[8] * P.fng}, fbb[3] = {R[6] * P.fnb, R[7] * P.fnb, R[8] * P.fnb};
    rhs[0] += fbg[0] + fbb[0];
    rhs[1] += fbg[1] + fbb[1];
    rhs[2] += fbg[2] + fbb[2];
    rhs[3] += (P.cog[1] * fbg[2] - P.cog[2] * fbg[1]) + (P.cob[1] * fbb[2] - P.cob[2] * fbb[1]);
    rhs[4] += (P.cog[2] * fbg[0] - P.cog[0] * fbg[2]) + (P.cob[2] * fbb[0] - P.cob[0] * fbb[2]);
    rhs[5] += (P.cog[0] * fbg[1] - P.cog[1] * fbg[0]) + (P.cob[0] * fbb[1] - P.cob[1] * fbb[0]);
#pragma unroll
    for (int r = 0; r < 6; r++) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 6; c++) acc = fmaf(P.invM[6 * r + c], rhs[c], acc);
        xd[7 + r] = acc;
    }
}

// ---- the reference's learned AUV model as a dynamics functor (NNAUVModel, scripts/src/models/nn_model.py:181-304) ----
//   X = (concat(x[3:13], u) - Xmean) / Xstd ; h = relu(h W_l + b_l) per hidden layer ; d = (h W_o + b_o) Ystd + Ymean ; x <- x + d
// (prepare_data :289-293, Sequential of Dense layers :54-60, denormalizeY :295-297, next_state :303-304: a plain add, the
// quaternion is not renormalised).  Weights sit in shared memory (every thread reads the same address: broadcast LDS.128),
// activations in registers; fp32 FFMA throughout - about 3 kFLOP per sample-step of mostly 32 x 32 layers, too thin for
// the 128-row tensor-core tile the config-4 kernel is built around.
// Blob: [0,16) Xmean, [16,32) 1/Xstd, [32,48) Ystd, [48,64) Ymean, W_0 [16][32], b_0 [32], then per further hidden layer
// W [32][32], b [32], then W_out [32][16] (13 columns used), b_out [16].
constexpr int kNnW = 32, kNnIn = 16, kNnOutPad = 16, kNnMaxHidden = 4;
__host__ __device__ inline int nn_auv_blob_floats(int n_hidden)
{
    return 64 + kNnIn * kNnW + kNnW + (n_hidden - 1) * (kNnW * kNnW + kNnW) + kNnW * kNnOutPad + kNnOutPad;
}

template <int NIN, int NOUT>
__device__ __forceinline__ void nn_dense(const float *W, const float (&in)[kNnW], float (&out)[kNnW])
{
    const float *b = W + NIN * NOUT;
#pragma unroll
    for (int o = 0; o < NOUT; o += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(b + o);
        out[o] = v.x; out[o + 1] = v.y; out[o + 2] = v.z; out[o + 3] = v.w;
    }
#pragma unroll
    for (int q = 0; q < NIN; q++) {
        const float xq = in[q];
#pragma unroll
        for (int o = 0; o < NOUT; o += 4) {
            const float4 w = *reinterpret_cast<const float4 *>(W + q * NOUT + o);
            out[o] = fmaf(xq, w.x, out[o]); out[o + 1] = fmaf(xq, w.y, out[o + 1]);
            out[o + 2] = fmaf(xq, w.z, out[o + 2]); out[o + 3] = fmaf(xq, w.w, out[o + 3]);
        }
    }
}

static __device__ __noinline__ void nn_auv_step(const float *nn, int n_hidden, float (&x)[kAuvS], const float (&u)[kAuvA])
{
    float a0[kNnW], a1[kNnW];
#pragma unroll
    for (int j = 0; j < 10; j++) a0[j] = (x[3 + j] - nn[j]) * nn[16 + j];
#pragma unroll
    for (int j = 0; j < kAuvA; j++) a0[10 + j] = (u[j] - nn[10 + j]) * nn[26 + j];
    const float *W = nn + 64;
    nn_dense<kNnIn, kNnW>(W, a0, a1);
#pragma unroll
    for (int o = 0; o < kNnW; o++) a0[o] = fmaxf(a1[o], 0.f);
    W += kNnIn * kNnW + kNnW;
    for (int l = 1; l < n_hidden; l++) {            // grid-uniform trip count
        nn_dense<kNnW, kNnW>(W, a0, a1);
#pragma unroll
        for (int o = 0; o < kNnW; o++) a0[o] = fmaxf(a1[o], 0.f);
        W += kNnW * kNnW + kNnW;
    }
    nn_dense<kNnW, kNnOutPad>(W, a0, a1);
#pragma unroll
    for (int j = 0; j < kAuvS; j++) x[j] += fmaf(a1[j], nn[32 + j], nn[48 + j]);
}

// x <- step(x, u): explicit Euler / Heun / the reference's rk-4 variant, then quaternion normalisation.
// RK is a template parameter: the register budget of the rollout kernel is set by the integrator.
template <int RK>
__device__ __forceinline__ void auv_step(const AuvParams &P, float (&x)[kAuvS], const float (&u)[kAuvA])
{
    float k1[kAuvS], xs[kAuvS], acc[kAuvS];
    const float dt = P.dt;
    auv_state_dot(P, x, u, k1);
    if (RK == 2) {
        float k2[kAuvS];
#pragma unroll
        for (int j = 0; j < kAuvS; j++) xs[j] = fmaf(dt, k1[j], x[j]);
        auv_state_dot(P, xs, u, k2);
#pragma unroll
        for (int j = 0; j < kAuvS; j++) acc[j] = 0.5f * dt * (k1[j] + k2[j]);
    } else if (RK == 4) {
        float kk[kAuvS];
#pragma unroll
        for (int j = 0; j < kAuvS; j++) { xs[j] = fmaf(0.5f * dt, k1[j], x[j]); acc[j] = k1[j]; }
        auv_state_dot(P, xs, u, kk);
#pragma unroll
        for (int j = 0; j < kAuvS; j++) { xs[j] = fmaf(0.5f * dt, kk[j], x[j]); acc[j] = fmaf(2.f, kk[j], acc[j]); }
        auv_state_dot(P, xs, u, kk);
#pragma unroll
        for (int j = 0; j < kAuvS; j++) { xs[j] = fmaf(dt, kk[j], x[j]); acc[j] = fmaf(2.f, kk[j], acc[j]); }
        auv_state_dot(P, xs, u, kk);
#pragma unroll
        for (int j = 0; j < kAuvS; j++) acc[j] = (1.0f / 6.0f) * fmaf(kk[j], dt, acc[j]) * dt;   // k4 * dt inside, as the reference writes it (:300-301)
    } else {
#pragma unroll
        for (int j = 0; j < kAuvS; j++) acc[j] = k1[j] * dt;
    }
#pragma unroll
    for (int j = 0; j < kAuvS; j++) x[j] += acc[j];
    const float n2 = fmaf(x[3], x[3], fmaf(x[4], x[4], fmaf(x[5], x[5], x[6] * x[6])));
    const float inv = rsqrtf(fmaxf(n2, 1e-12f));
#pragma unroll
    for (int j = 3; j < 7; j++) x[j] *= inv;
}

// Hamilton product of quaternions stored (x, y, z, w)
__device__ __forceinline__ void quat_mul(const float *a, const float *b, float *o)
{
    o[0] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    o[1] = -a[0] * b[2] + a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    o[2] = a[0] * b[1] - a[1] * b[0] + a[2] * b[3] + a[3] * b[2];
    o[3] = -a[0] * b[0] - a[1] * b[1] - a[2] * b[2] + a[3] * b[3];
}

// ElipseCost3D.state_cost of one sample (scripts/src/costs/elipse_cost.py:156-246):
//   p' = q p q*, q' = q quat;  m_state (| (p'x/a)^2 + (p'y/b)^2 + p'z^2 - 1 | + angle(q_t, q')) + m_vel | |v|^2 - speed^2 |
// with q_t the shortest rotation from (1,0,0) to the ellipse tangent (-a/b p'y, b/a p'x, 0) and angle = 2 acos|q_t . q'|.
__device__ __forceinline__ float ellipse3d_cost(const float *ell, const float (&x)[kAuvS])
{
    const float q[4] = {ell[0], ell[1], ell[2], ell[3]}, qc[4] = {-ell[0], -ell[1], -ell[2], ell[3]};
    const float p4[4] = {x[0], x[1], x[2], 0.f}, xq[4] = {x[3], x[4], x[5], x[6]};
    float t[4], pf[4], qpf[4];
    quat_mul(q, p4, t);
    quat_mul(t, qc, pf);
    quat_mul(q, xq, qpf);
    const float dx = pf[0] * ell[4], dy = pf[1] * ell[5];
    const float pos = fabsf(fmaf(dx, dx, fmaf(dy, dy, pf[2] * pf[2])) - 1.0f);
    float tx = pf[1] * ell[6], ty = pf[0] * ell[7];
    const float inv = rsqrtf(fmaxf(fmaf(tx, tx, ty * ty), 1e-24f));
    tx *= inv;
    ty *= inv;
    // between_two_vectors_3d((1,0,0), (tx, ty, 0)) = normalise(0, 0, ty, 1 + tx); antiparallel: (0, 0, 1, 0)
    float rz = ty, rw = 1.0f + tx;
    if (rw < 1e-6f) { rz = 1.0f; rw = 0.f; }
    const float ir = rsqrtf(fmaxf(fmaf(rz, rz, rw * rw), 1e-12f));
    const float dot = fminf(fabsf(fmaf(rz * ir, qpf[2], (rw * ir) * qpf[3])), 1.0f);
    const float ori = 2.0f * acosf(dot);
    const float dv = fabsf(fmaf(x[7], x[7], fmaf(x[8], x[8], x[9] * x[9])) - ell[8]);
    return fmaf(ell[9], pos + ori, ell[10] * dv);
}

// State costs on the 13-dimensional state.
//   kind 0  StaticCost      sum_i q_i (x_i - g_i)^2                                   (static_cost.py:40-63, diagonal Q)
//   kind 2  StaticQuatCost  d = (p - g_p, 2 acos(q . g_q), nu - g_nu), sum_i q_i d_i^2  (static_cost.py:116-159, diagonal Q[10])
//   kind 3  ElipseCost3D    (above)
__device__ __forceinline__ float auv_state_cost(int kind, const float *q, const float *g, const float *ell, const float (&x)[kAuvS])
{
    float c = 0.f;
    if (kind == 3) return ellipse3d_cost(ell, x);
    if (kind == 2) {
#pragma unroll
        for (int i = 0; i < 3; i++) { const float d = x[i] - g[i]; c = fmaf(q[i] * d, d, c); }
        const float dot = fmaf(x[3], g[3], fmaf(x[4], g[4], fmaf(x[5], g[5], x[6] * g[6])));
        const float th = 2.0f * acosf(fminf(fmaxf(dot, -1.0f), 1.0f));
        c = fmaf(q[3] * th, th, c);
#pragma unroll
        for (int i = 0; i < 6; i++) { const float d = x[7 + i] - g[7 + i]; c = fmaf(q[4 + i] * d, d, c); }
    } else {
#pragma unroll
        for (int i = 0; i < kAuvS; i++) { const float d = x[i] - g[i]; c = fmaf(q[i] * d, d, c); }
    }
    return c;
}

}  // namespace mppi
