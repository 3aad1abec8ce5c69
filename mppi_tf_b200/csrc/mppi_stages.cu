// Stage kernels: the reference's graph-builder methods as stand-alone CUDA kernels on plain buffers,
// so that each unit KAT of /root/reference/test/*.cpp can be replayed against device code.
// They share the arithmetic (PointMass, weight_exp, Philox) with the fused kernels.
#include "mppi_device.cuh"
#include "mppi_internal.h"

namespace mppi {

// ModelBase::mBuildFreeStepGraph / mBuildActionStepGraph / mBuildModelStepGraph
// (src/model_base.cpp:53-82).  mode 0: out[kst][s] = A state; 1: out[k][s] = (B/m) action;
// 2: out[k][s] = A state + (B/m) action with state broadcast when kst == 1.
__global__ void model_step_kernel(float dt, float c_pu, float c_vu, int s, int a, int kst, int k,
                                  const float *state, const float *action, float *out, int mode)
{
    const int n = (mode == 0 ? kst : k) * a;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int smp = i / a, ax = i - smp * a;
        float p = 0.f, v = 0.f, u = 0.f;
        if (mode != 1) {
            const float *st = state + (size_t)(kst == 1 ? 0 : smp) * s;
            p = st[2 * ax];
            v = st[2 * ax + 1];
        }
        if (mode != 0) u = action[(size_t)smp * a + ax];
        // free part first, then the action part, then the add: the order of src/model_base.cpp:53-57
        const float fp = fmaf(dt, v, p), fv = v;
        const float ap = c_pu * u, av = c_vu * u;
        out[(size_t)smp * s + 2 * ax] = (mode == 0) ? fp : (mode == 1 ? ap : fp + ap);
        out[(size_t)smp * s + 2 * ax + 1] = (mode == 0) ? fv : (mode == 1 ? av : fv + av);
    }
}

// CostBase (src/cost_base.cpp:43-68).  mode 0: state cost; 1: action cost; 2: step cost.
__global__ void cost_kernel(int k, int s, int a, float lambda, const float *inv_sigma, const float *goal,
                            const float *q, const float *state, const float *action, const float *noise,
                            float *out, int mode)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
        float c = 0.f;
        if (mode != 1) {
            for (int j = 0; j < s; j++) {
                const float d = state[(size_t)i * s + j] - goal[j];
                c = fmaf(q[j] * d, d, c);
            }
        }
        if (mode != 0) {
            float ac = 0.f;
            for (int r = 0; r < a; r++) {
                float nc = 0.f;
                for (int l = 0; l < a; l++) nc = fmaf(inv_sigma[r * a + l], noise[(size_t)i * a + l], nc);
                ac = fmaf(action[r], nc, ac);
            }
            c += lambda * ac;
        }
        out[i] = c;
    }
}

// CostBase.action_cost of the Python twin (scripts/src/costs/cost_base.py:114-170):
//   0.5 [gamma (u^T S^-1 u + 2 u^T S^-1 eps) + lambda (1 - 1/upsilon) eps^T S^-1 eps]
__global__ void action_cost_py_kernel(int k, int a, float lambda, float gamma, float upsilon, const float *inv_sigma,
                                      const float *action, const float *noise, float *out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
        float uu = 0.f, ue = 0.f, ee = 0.f;
        for (int r = 0; r < a; r++) {
            float su = 0.f, se = 0.f;
            for (int l = 0; l < a; l++) {
                su = fmaf(inv_sigma[r * a + l], action[l], su);
                se = fmaf(inv_sigma[r * a + l], noise[(size_t)i * a + l], se);
            }
            uu = fmaf(action[r], su, uu);
            ue = fmaf(action[r], se, ue);
            ee = fmaf(noise[(size_t)i * a + r], se, ee);
        }
        out[i] = 0.5f * (gamma * (uu + 2.0f * ue) + lambda * (1.0f - 1.0f / upsilon) * ee);
    }
}

// ElipseCost.state_cost (scripts/src/costs/elipse_cost.py:46-79): state [k][4] = (x, vx, y, vy)
__global__ void ellipse_cost_kernel(int k, const float *state, float ia, float ib, float cx, float cy, float speed,
                                    float m_state, float m_vel, float *out)
{
    const float ell[8] = {ia, ib, cx, cy, speed, m_state, m_vel, 0.f};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x)
        out[i] = ellipse_cost(ell, state[4 * i], state[4 * i + 1], state[4 * i + 2], state[4 * i + 3]);
}

// ControllerBase::mPrepareNoise (src/controller_base.cpp:210-213): noise[:, t] -> [k][a]
__global__ void prepare_noise_kernel(int k, int T, int a, const float *noise, int t, float *out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k * a; i += gridDim.x * blockDim.x) {
        const int smp = i / a, j = i - smp * a;
        out[i] = noise[((size_t)smp * T + t) * a + j];
    }
}

// mBeta, mExpArg, mExp, mNabla, mWeights, mWeightedNoise (src/controller_base.cpp:166-192), one CTA.
__global__ void __launch_bounds__(1024) update_stages_kernel(int k, int T, int a, float lambda, const float *cost,
                                                             const float *noise, float *scal, float *exp_arg,
                                                             float *exp_out, float *weights, float *wn)
{
    __shared__ float sRed[32];
    __shared__ float sBeta, sNabla;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    float b = kInf;
    for (int i = tid; i < k; i += blockDim.x) b = fminf(b, cost[i]);
    b = warp_min(b);
    if (lane == 0) sRed[warp] = b;
    __syncthreads();
    if (tid == 0) {
        float m = sRed[0];
        for (int w = 1; w < nw; w++) m = fminf(m, sRed[w]);
        sBeta = m;
    }
    __syncthreads();
    const float beta = sBeta;
    const float neg_inv_lambda = -1.0f / lambda;
    float e_sum = 0.f;
    for (int i = tid; i < k; i += blockDim.x) {
        const float arg = neg_inv_lambda * (cost[i] - beta);
        const float e = expf(arg);
        if (exp_arg) exp_arg[i] = arg;
        exp_out[i] = e;
        e_sum += e;
    }
    e_sum = warp_sum(e_sum);
    if (lane == 0) sRed[warp] = e_sum;
    __syncthreads();
    if (tid == 0) {
        float n = 0.f;
        for (int w = 0; w < nw; w++) n += sRed[w];
        sNabla = n;
        scal[0] = beta;
        scal[1] = n;
    }
    __syncthreads();
    const float nabla = sNabla;
    for (int i = tid; i < k; i += blockDim.x) weights[i] = exp_out[i] / nabla;
    __syncthreads();
    for (int j = tid; j < T * a; j += blockDim.x) {
        float acc = 0.f;
        for (int i = 0; i < k; i++) acc = fmaf(weights[i], noise[(size_t)i * T * a + j], acc);
        wn[j] = acc;
    }
}

// One stage at a time (mBeta / mExpArg / mExp / mNabla / mWeights), single CTA.
__global__ void __launch_bounds__(1024) vector_op_kernel(int op, int k, const float *in, float s0, float s1, float *out)
{
    __shared__ float sRed[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    if (op == MPPI_OP_MIN || op == MPPI_OP_SUM) {
        float v = (op == MPPI_OP_MIN) ? kInf : 0.f;
        for (int i = tid; i < k; i += blockDim.x) v = (op == MPPI_OP_MIN) ? fminf(v, in[i]) : v + in[i];
        v = (op == MPPI_OP_MIN) ? warp_min(v) : warp_sum(v);
        if (lane == 0) sRed[warp] = v;
        __syncthreads();
        if (tid == 0) {
            float r = sRed[0];
            for (int w = 1; w < nw; w++) r = (op == MPPI_OP_MIN) ? fminf(r, sRed[w]) : r + sRed[w];
            out[0] = r;
        }
        return;
    }
    const float neg_inv = -1.0f / s1;
    for (int i = tid; i < k; i += blockDim.x) {
        const float x = in[i];
        out[i] = (op == MPPI_OP_EXP_ARG) ? neg_inv * (x - s0) : (op == MPPI_OP_EXP ? expf(x) : x / s0);
    }
}

// mWeightedNoise (src/controller_base.cpp:188-192): out[j] = sum_k w_k noise[k][j]
__global__ void weighted_noise_kernel(int k, int TA, const float *weights, const float *noise, float *out)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < TA; j += gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int i = 0; i < k; i++) acc = fmaf(weights[i], noise[(size_t)i * TA + j], acc);
        out[j] = acc;
    }
}

__global__ void philox_raw_kernel(uint32_t key0, uint32_t key1, uint32_t call0, uint32_t sample, uint32_t update,
                                  uint32_t stream, int n_calls, int rounds, uint32_t *out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_calls; i += gridDim.x * blockDim.x) {
        const uint4 x = philox4x32_10(call0 + (uint32_t)i, sample, update, stream, key0, key1, rounds);
        out[4 * i] = x.x; out[4 * i + 1] = x.y; out[4 * i + 2] = x.z; out[4 * i + 3] = x.w;
    }
}

static inline int blocks_for(long long n, int threads)
{
    long long b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > 148 * 8) b = 148 * 8;
    return (int)b;
}

cudaError_t launch_model_step(float mass, float dt, int s, int a, int kst, int k, const float *state,
                              const float *action, float *out, int mode, cudaStream_t st)
{
    const float c_pu = (dt * dt) / 2.0f / mass, c_vu = dt / mass;
    const long long n = (long long)(mode == 0 ? kst : k) * a;
    model_step_kernel<<<blocks_for(n, 256), 256, 0, st>>>(dt, c_pu, c_vu, s, a, kst, k, state, action, out, mode);
    return cudaGetLastError();
}

cudaError_t launch_cost(int k, int s, int a, float lambda, const float *inv_sigma, const float *goal,
                        const float *q, const float *state, const float *action, const float *noise,
                        float *out, int mode, cudaStream_t st)
{
    cost_kernel<<<blocks_for(k, 256), 256, 0, st>>>(k, s, a, lambda, inv_sigma, goal, q, state, action, noise, out, mode);
    return cudaGetLastError();
}

cudaError_t launch_action_cost_py(int k, int a, float lambda, float gamma, float upsilon, const float *inv_sigma,
                                  const float *action, const float *noise, float *out, cudaStream_t st)
{
    action_cost_py_kernel<<<blocks_for(k, 256), 256, 0, st>>>(k, a, lambda, gamma, upsilon, inv_sigma, action, noise, out);
    return cudaGetLastError();
}

cudaError_t launch_ellipse_cost(int k, const float *state, const float *ell, float *out, cudaStream_t st)
{
    ellipse_cost_kernel<<<blocks_for(k, 256), 256, 0, st>>>(k, state, ell[0], ell[1], ell[2], ell[3], ell[4], ell[5], ell[6], out);
    return cudaGetLastError();
}

cudaError_t launch_prepare_noise(int k, int T, int a, const float *noise, int t, float *out, cudaStream_t st)
{
    prepare_noise_kernel<<<blocks_for((long long)k * a, 256), 256, 0, st>>>(k, T, a, noise, t, out);
    return cudaGetLastError();
}

cudaError_t launch_update_stages(int k, int T, int a, float lambda, const float *cost, const float *noise,
                                 float *scal, float *exp_arg, float *exp_out, float *weights,
                                 float *weighted_noise, cudaStream_t st)
{
    update_stages_kernel<<<1, 1024, 0, st>>>(k, T, a, lambda, cost, noise, scal, exp_arg, exp_out, weights,
                                             weighted_noise);
    return cudaGetLastError();
}

cudaError_t launch_vector_op(int op, int k, const float *in, float s0, float s1, float *out, cudaStream_t st)
{
    vector_op_kernel<<<1, 1024, 0, st>>>(op, k, in, s0, s1, out);
    return cudaGetLastError();
}

cudaError_t launch_weighted_noise(int k, int TA, const float *weights, const float *noise, float *out, cudaStream_t st)
{
    weighted_noise_kernel<<<blocks_for(TA, 128), 128, 0, st>>>(k, TA, weights, noise, out);
    return cudaGetLastError();
}

cudaError_t launch_philox_raw(uint64_t seed, uint32_t call0, uint32_t sample, uint32_t update, uint32_t stream,
                              int n_calls, int rounds, uint32_t *out, cudaStream_t st)
{
    philox_raw_kernel<<<blocks_for(n_calls, 256), 256, 0, st>>>((uint32_t)seed, (uint32_t)(seed >> 32), call0, sample,
                                                               update, stream, n_calls, rounds, out);
    return cudaGetLastError();
}

bool invert_matrix(const float *m, int n, float *inv)
{
    double w[kMaxA][2 * kMaxA];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            w[i][j] = m[i * n + j];
            w[i][n + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < n; c++) {
        int piv = c;
        for (int r = c + 1; r < n; r++)
            if (fabs(w[r][c]) > fabs(w[piv][c])) piv = r;
        if (w[piv][c] == 0.0) return false;
        if (piv != c)
            for (int j = 0; j < 2 * n; j++) { const double t = w[c][j]; w[c][j] = w[piv][j]; w[piv][j] = t; }
        const double d = w[c][c];
        for (int j = 0; j < 2 * n; j++) w[c][j] /= d;
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            const double f = w[r][c];
            if (f == 0.0) continue;
            for (int j = 0; j < 2 * n; j++) w[r][j] -= f * w[c][j];
        }
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) inv[i * n + j] = (float)w[i][n + j];
    return true;
}

}  // namespace mppi
