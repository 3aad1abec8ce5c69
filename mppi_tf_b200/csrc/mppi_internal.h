// Host-side declarations shared between the translation units of libmppi_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mppi {

struct RolloutParams;

// mppi_rollout.cu
cudaError_t launch_rollout_philox(RolloutParams p, int a, int num_sms, size_t smem_limit, cudaStream_t st, int *grid_x_out);
cudaError_t launch_rollout_injected(RolloutParams p, int a, int num_sms, size_t smem_limit, cudaStream_t st,
                                    int *grid_x_out);
cudaError_t launch_finish(RolloutParams p, int a, bool philox, const float *gathered, cudaStream_t st);
cudaError_t launch_dump_noise(RolloutParams p, int a, float *out_dev, cudaStream_t st);
int max_grid_x(int K_local, int n_ctrl, int num_sms);
int philox_grid_x(int K_local, int n_ctrl, int num_sms);

// mppi_rollout_fast.cu: the superposition kernels (p.fast set by the host).  variant: 0 = pick, 1 = regenerating kernel,
// 2 = resident-tile kernel (cudaErrorInvalidConfiguration when the rows are too long for it)
cudaError_t launch_rollout_philox_fast(RolloutParams p, int a, int variant, int num_sms, size_t smem_sm, size_t smem_limit,
                                       cudaStream_t st, int *grid_x_out);
bool resident_geometry(int A, int T, int TA, int K_local, int n_ctrl, int num_sms, size_t smem_sm, size_t smem_cta_limit, int *nw_out,
                       int *rs_out, int *cw_out, int *grid_x_out, size_t *smem_out, bool *latency_out);

#define MPPI_DISPATCH_A(a, ...)                  \
    switch (a) {                                 \
        case 1: { constexpr int A_ = 1; __VA_ARGS__; } break; \
        case 2: { constexpr int A_ = 2; __VA_ARGS__; } break; \
        case 3: { constexpr int A_ = 3; __VA_ARGS__; } break; \
        case 4: { constexpr int A_ = 4; __VA_ARGS__; } break; \
        case 5: { constexpr int A_ = 5; __VA_ARGS__; } break; \
        case 6: { constexpr int A_ = 6; __VA_ARGS__; } break; \
        case 7: { constexpr int A_ = 7; __VA_ARGS__; } break; \
        case 8: { constexpr int A_ = 8; __VA_ARGS__; } break; \
        default: return cudaErrorInvalidValue;   \
    }
bool injected_geometry(int A, int T, int TA, int K_local, int n_ctrl, int num_sms, size_t smem_limit,
                       int *ng_out, int *c_out, int *nbuf_out, int *grid_x_out, size_t *smem_out);

// mppi_mlp.cu
struct MlpParams;
cudaError_t launch_mlp_predict(const MlpParams &mp, int kst, int k, const float *state, const float *action, float *out,
                               cudaStream_t st);
cudaError_t launch_rollout_mlp(RolloutParams p, const MlpParams &mp, int a, bool philox, int num_sms, cudaStream_t st,
                               int *grid_x_out);
void mlp_pack_weights(int s, int a, const float *W1, const float *b1, const float *W2, const float *b2,
                      const float *W3, const float *b3, void *blob_host);

// mppi_auv.cu
struct AuvParams;
cudaError_t launch_rollout_auv(RolloutParams p, const AuvParams &P, bool philox, int num_sms, cudaStream_t st, int *grid_x_out);
cudaError_t launch_auv_predict(const AuvParams &P, int kst, int k, const float *state, const float *action, float *out,
                               cudaStream_t st);
cudaError_t launch_auv_cost(int kind, int k, const float *q /*[13] host*/, const float *goal /*[13] host*/,
                            const float *ell /*[12] host*/, const float *state, float *out, cudaStream_t st);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) costs a microsecond or two per call: remember, per kernel
// instantiation and device, the largest size already granted and only raise it.
template <auto kernel>
inline cudaError_t ensure_dyn_smem(size_t smem)
{
    static size_t granted[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (smem <= granted[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) granted[dev] = smem;
    return e;
}

// mppi_train.cu  (device pointers)
size_t mlp_param_count(int s, int a);
size_t train_work_floats(int s, int a, int n);
cudaError_t launch_pack_blob(int s, int a, const float *params, void *blob, cudaStream_t st);
cudaError_t launch_train_step(int s, int a, int n, const float *x, const float *u, const float *xnext, const float *norm,
                              float *params, float *adam_m, float *adam_v, float lr_t, float b1, float b2, float eps,
                              float *work, float *loss_dev, cudaStream_t st);

cudaError_t launch_augment(int n, int samples, int s, int a, const float *x, const float *u, const float *xnext, const float *norm,
                           float sigma, uint64_t seed, uint32_t epoch, float *xo, float *uo, float *xno, cudaStream_t st);

// mppi_stages.cu  (device pointers)
cudaError_t launch_model_step(float mass, float dt, int s, int a, int kst, int k, const float *state,
                              const float *action, float *out, int mode, cudaStream_t st);
cudaError_t launch_cost(int k, int s, int a, float lambda, const float *inv_sigma, const float *goal,
                        const float *q, const float *state, const float *action, const float *noise,
                        float *out, int mode, cudaStream_t st);
cudaError_t launch_action_cost_py(int k, int a, float lambda, float gamma, float upsilon, const float *inv_sigma,
                                  const float *action, const float *noise, float *out, cudaStream_t st);
cudaError_t launch_ellipse_cost(int k, const float *state, const float *ell /*[7] host*/, float *out, cudaStream_t st);
cudaError_t launch_prepare_noise(int k, int T, int a, const float *noise, int t, float *out, cudaStream_t st);
cudaError_t launch_update_stages(int k, int T, int a, float lambda, const float *cost, const float *noise,
                                 float *scal /*beta,nabla*/, float *exp_arg, float *exp_out, float *weights,
                                 float *weighted_noise, cudaStream_t st);
cudaError_t launch_vector_op(int op, int k, const float *in, float s0, float s1, float *out, cudaStream_t st);
cudaError_t launch_weighted_noise(int k, int TA, const float *weights, const float *noise, float *out, cudaStream_t st);
cudaError_t launch_philox_raw(uint64_t seed, uint32_t call0, uint32_t sample, uint32_t update, uint32_t stream,
                              int n_calls, int rounds, uint32_t *out, cudaStream_t st);

// Host helper: inverse of an a x a matrix (Gauss-Jordan, double); returns false when singular.
bool invert_matrix(const float *m, int n, float *inv);

}  // namespace mppi
