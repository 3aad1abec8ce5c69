// C-ABI of libmppi_b200.so (see include/mppi_b200.h for the contract and the reference
// methods each entry point replaces).  Host-side only: owns the device buffers, builds the kernel
// parameter block and launches the kernels of mppi_rollout.cu / mppi_stages.cu.
#include <dlfcn.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "mppi_device.cuh"
#include "mppi_internal.h"
#include "mppi_auv.cuh"
#include "mppi_mlp.cuh"
#include "mppi_linear.cuh"
#include "mppi_update.cuh"

using namespace mppi;

// ---- minimal NCCL surface, resolved with dlopen so single-GPU users need no NCCL at all -----------
namespace {
struct NcclUniqueId { char internal[128]; };
typedef void *NcclComm;
typedef int (*fn_ncclGetUniqueId)(NcclUniqueId *);
typedef int (*fn_ncclCommInitRank)(NcclComm *, int, NcclUniqueId, int);
typedef int (*fn_ncclAllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t);
typedef int (*fn_ncclCommDestroy)(NcclComm);
typedef const char *(*fn_ncclGetErrorString)(int);
struct NcclApi {
    void *lib = nullptr;
    fn_ncclGetUniqueId GetUniqueId = nullptr;
    fn_ncclCommInitRank CommInitRank = nullptr;
    fn_ncclAllGather AllGather = nullptr;
    fn_ncclCommDestroy CommDestroy = nullptr;
    fn_ncclGetErrorString GetErrorString = nullptr;
};
NcclApi g_nccl;
std::string g_last_error;   // errors without a handle (create / stateless stages)

bool load_nccl(std::string &err)
{
    if (g_nccl.lib) return true;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = std::string("dlopen libnccl.so.2 failed: ") + dlerror(); return false; }
    g_nccl.GetUniqueId = (fn_ncclGetUniqueId)dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (fn_ncclCommInitRank)dlsym(lib, "ncclCommInitRank");
    g_nccl.AllGather = (fn_ncclAllGather)dlsym(lib, "ncclAllGather");
    g_nccl.CommDestroy = (fn_ncclCommDestroy)dlsym(lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (fn_ncclGetErrorString)dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) {
        err = "libnccl.so.2 lacks a required symbol";
        return false;
    }
    g_nccl.lib = lib;
    return true;
}
}  // namespace

struct mppi_handle {
    int k = 0, tau = 0, s = 0, a = 0, n_ctrl = 1, rank = 0, world = 1;
    int K_local = 0, k_offset = 0, TA = 0, stride = 0;
    float dt = 0, mass = 1, lambda = 1;
    float sigma[kMaxA * kMaxA], lam_inv_sigma_T[kMaxA * kMaxA], q[kMaxS];
    int sigma_diag = 1, goal_per_ctrl = 0;
    // Python-twin extras (mppi_set_action_cost / mppi_set_normalize_cost)
    int cost_form = MPPI_ACTION_COST_CPP;
    float gamma = 1, upsilon = 1;
    float inv_sigma[kMaxA * kMaxA];
    bool normalize = false;
    float *d_norm = nullptr;
    bool clip = false;            // mppi_set_action_limits
    float act_min[kMaxA], act_max[kMaxA];
    int cost_kind = 0;            // 0 StaticCost, 1 ElipseCost (mppi_set_ellipse_cost), 2 StaticQuatCost, 3 ElipseCost3D
    float q10[kMaxS] = {0};       // StaticQuatCost weights, kept apart from q so that mppi_set_static_cost finds q untouched
    float ell[12] = {0};
    uint64_t seed = 1;
    int rounds = 10;              // Philox4x32-R (mppi_config.philox_rounds)
    int kernel_variant = 0;       // developer knob MPPI_PHILOX_KERNEL: 0 pick, 1 regenerating superposition kernel, 2 resident-tile
                                  // kernel, 3 direct-form kernels only
    bool last_fast = false;       // the last update ran in the superposition form: costs / beta on the device are relative to d_cost_base
    float *d_cost_base = nullptr;
    size_t smem_sm = 0;
    unsigned long long *d_trace = nullptr;      // developer knob: phase time stamps (mppi_debug_trace)
    size_t trace_ctas = 0;
    uint32_t update_counter = 0, last_update = 0;
    bool have_philox_update = false, last_philox = true, pending_finish = false;
    int device = 0, num_sms = 148;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int max_gx = 1, last_gx = 0, max_groups = 1;
    float *d_partials2 = nullptr;
    // device buffers
    float *d_x = nullptr, *d_goal = nullptr, *d_U = nullptr, *d_Unew = nullptr, *d_next = nullptr;
    float *d_costs = nullptr, *d_partials = nullptr, *d_payload = nullptr, *d_gather = nullptr, *d_stats = nullptr;
    unsigned int *d_counters = nullptr;
    float *d_eps_tmp = nullptr;
    bool ext_exchange = false;
    // pinned host staging
    float *h_x = nullptr, *h_next = nullptr;
    // batched handles copy the states H2D asynchronously: two staging buffers, each guarded by the event of its last copy,
    // so an asynchronous caller (set_state / enqueue_update in a loop) never overwrites a buffer the DMA still reads
    float *h_xs[2] = {nullptr, nullptr};
    cudaEvent_t x_ev[2] = {nullptr, nullptr};
    bool x_ev_armed[2] = {false, false};
    int x_idx = 0;
    // zero-copy result path (n_ctrl == 1): mapped pinned [a floats | pad | completion word]
    float *h_zc = nullptr, *d_zc = nullptr;
    unsigned int zc_epoch = 0;
    bool zc_armed = false, zc_request = false;   // requested by the synchronous next(): async callers do not pay the host write
    bool x_staged = false;
    NcclComm comm = nullptr;
    // fused exchange over peer memory (mppi_peer_handle / mppi_peer_attach)
    void *d_mailbox = nullptr;            // [2][world][n_ctrl][stride] floats + [2][world][n_ctrl] epochs
    size_t mail_floats = 0;
    size_t mail_ll_off = 0;               // byte offset of the tagged-word region inside the mailbox
    void *peer_base[kMaxWorld] = {nullptr};
    bool peer_opened[kMaxWorld] = {false};
    bool peer_on = false;
    uint32_t epoch = 0;
    unsigned int *d_peer_status = nullptr, *h_peer_status = nullptr;
    // AUV (Fossen) dynamics (cfg.model = MPPI_MODEL_AUV + mppi_set_auv_model)
    bool auv = false, auv_ready = false;
    AuvParams auv_prm;
    float *d_nn_blob = nullptr;   // learned AUV model (mppi_set_nn_auv_model)
    // learned-MLP dynamics (mppi_set_mlp)
    bool mlp = false;
    int mlp_hidden = 128;         // logical width; the device always runs 128-wide tiles (zero padded)
    void *d_wblob = nullptr;
    float *d_fvec = nullptr;
    // learner side (mppi_mlp_train_step): fp32 master weights + Adam state, W1 b1 W2 b2 W3 b3 back to back
    float *d_params = nullptr, *d_adam_m = nullptr, *d_adam_v = nullptr, *d_train_work = nullptr, *d_loss = nullptr;
    size_t train_work_cap = 0;
    long long adam_t = 0;
    float adam_b1 = 0.9f, adam_b2 = 0.999f, adam_eps = 1e-7f;     // tf.optimizers.Adam defaults (learner_base.py:325)
    std::string err;
};

namespace {

int fail(mppi_handle *h, int code, const std::string &msg)
{
    if (h) h->err = msg;
    g_last_error = msg;
    return code;
}

#define CU_TRY(h, expr)                                                                          \
    do {                                                                                         \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail(h, MPPI_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));   \
    } while (0)

int select_device(mppi_handle *h, int device, int *dev_out, cudaDeviceProp *prop)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(h, MPPI_ERR_CUDA,
                    std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                        " (libmppi_b200 has no CPU fallback)");
    int dev = device;
    if (dev < 0) CU_TRY(h, cudaGetDevice(&dev));
    if (dev >= n) return fail(h, MPPI_ERR_BAD_ARG, "device ordinal out of range");
    CU_TRY(h, cudaSetDevice(dev));
    CU_TRY(h, cudaGetDeviceProperties(prop, dev));
    if (prop->major != 10)
        return fail(h, MPPI_ERR_CUDA, std::string("device ") + prop->name + " is sm_" + std::to_string(prop->major) +
                                          std::to_string(prop->minor) + "; this library carries sm_100a code only");
    *dev_out = dev;
    return MPPI_OK;
}

void derive_sigma(mppi_handle *h)
{
    const int a = h->a;
    float inv[kMaxA * kMaxA];
    invert_matrix(h->sigma, a, inv);
    h->sigma_diag = 1;
    for (int i = 0; i < a; i++)
        for (int j = 0; j < a; j++) {
            if (i != j && h->sigma[i * a + j] != 0.f) h->sigma_diag = 0;
            // linear action-cost weight: lambda (src/cost_base.cpp:63-68) or gamma (cost_base.py:141-158)
            const float lin = h->cost_form == MPPI_ACTION_COST_PYTHON ? h->gamma : h->lambda;
            h->lam_inv_sigma_T[i * a + j] = lin * inv[j * a + i];   // lin * (Sigma^-1)^T
            h->inv_sigma[i * a + j] = inv[i * a + j];
        }
}

// Superposition form (mppi_linear.cuh): point-mass model, diagonal Sigma, q > 0, StaticCost, no noise-quadratic term.
bool fast_eligible(const mppi_handle *h, const float *eps_dev)
{
    if (h->auv || h->mlp || h->kernel_variant == 3) return false;
    if (h->cost_kind != 0 || h->tau > kFastMaxT) return false;
    if (!eps_dev && !h->sigma_diag) return false; // Philox mode folds the (diagonal) scale into the noise response; injected
                                                  // noise arrives scaled, any Sigma
    if (h->cost_form == MPPI_ACTION_COST_PYTHON && h->upsilon != 1.0f && h->lambda != 0.f) return false;   // kappa != 0
    for (int i = 0; i < h->s; i++)
        if (!(h->q[i] > 0.f)) return false;
    return true;
}

RolloutParams make_params(const mppi_handle *h, const float *eps_dev)
{
    RolloutParams p;
    memset(&p, 0, sizeof(p));
    p.K_local = h->K_local;
    p.k_offset = h->k_offset;
    p.T = h->tau;
    p.TA = h->TA;
    p.n_ctrl = h->n_ctrl;
    p.n_iter = 1;
    p.world = h->world;
    p.max_parts = h->max_gx;
    p.dt = h->dt;
    p.c_pu = (h->dt * h->dt) / 2.0f / h->mass;
    p.c_vu = h->dt / h->mass;
    p.lambda = h->lambda;
    p.neg_inv_lambda_log2e = -kLog2e / h->lambda;
    memcpy(p.q, h->cost_kind == 2 ? h->q10 : h->q, sizeof(p.q));
    for (int i = 0; i < kMaxS; i++) p.sqrt_q[i] = sqrtf(p.q[i] > 0.f ? p.q[i] : 0.f);
    const int a = h->a;
    const bool py = h->cost_form == MPPI_ACTION_COST_PYTHON;
    for (int i = 0; i < a * a; i++) p.sigma[i] = h->upsilon * h->sigma[i];   // sampling scale: eps = (upsilon Sigma) z
    memcpy(p.lam_inv_sigma_T, h->lam_inv_sigma_T, sizeof(p.lam_inv_sigma_T));
    memcpy(p.inv_sigma, h->inv_sigma, sizeof(p.inv_sigma));
    p.sigma_diag = h->sigma_diag;
    p.w_scale = (py ? h->gamma : h->lambda) * h->upsilon;    // u^T S^-1 (upsilon S z) = upsilon u^T z
    p.c0_scale = py ? 0.5f * h->gamma : 0.f;
    const float kappa = py ? 0.5f * h->lambda * (1.0f - 1.0f / h->upsilon) : 0.f;
    p.quad = (kappa != 0.f);
    if (p.quad) {
        // n = eps (injected): kappa * Sigma^-1 ; n = z (Philox): kappa * (uS)^T Sigma^-1 (uS); symmetrised
        float M[kMaxA * kMaxA];
        for (int i = 0; i < a; i++)
            for (int j = 0; j < a; j++) {
                if (eps_dev) {
                    M[i * a + j] = h->inv_sigma[i * a + j];
                } else {
                    float acc = 0.f;                       // (S^T S^-1 S)_ij = sum_l S_li (S^-1 S)_lj = S_ji ... kept general
                    for (int l = 0; l < a; l++)
                        for (int m = 0; m < a; m++) acc += h->sigma[l * a + i] * h->inv_sigma[l * a + m] * h->sigma[m * a + j];
                    M[i * a + j] = h->upsilon * h->upsilon * acc;
                }
            }
        for (int i = 0; i < a; i++)
            for (int j = 0; j < a; j++) p.quadm[i * a + j] = kappa * 0.5f * (M[i * a + j] + M[j * a + i]);
    }
    p.norm_mode = 0;
    p.norm = h->d_norm;
    p.clip = h->clip ? 1 : 0;
    if (h->clip) { memcpy(p.act_min, h->act_min, sizeof(p.act_min)); memcpy(p.act_max, h->act_max, sizeof(p.act_max)); }
    p.cost_kind = h->cost_kind;
    memcpy(p.ell, h->ell, sizeof(p.ell));
    p.goal_per_ctrl = h->goal_per_ctrl;
    p.key0 = (uint32_t)h->seed;
    p.key1 = (uint32_t)(h->seed >> 32);
    p.update = h->update_counter;
    for (int r = 0; r < 10; r++) {
        p.rk0[r] = p.key0 + (uint32_t)r * kPhiloxW0;
        p.rk1[r] = p.key1 + (uint32_t)r * kPhiloxW1;
    }
    p.rounds = h->rounds;
    p.fast = fast_eligible(h, eps_dev) ? 1 : 0;
    p.z_scale = 1.0f;
    p.cost_base = h->d_cost_base;
    if (p.fast) {
        p.z_scale = eps_dev ? 1.0f : kZScale;     // injected mode: n = eps, nothing folded
        for (int j = 0; j < a; j++) {
            const float sqp = p.sqrt_q[2 * j], sqv = p.sqrt_q[2 * j + 1], sg = eps_dev ? 1.0f : p.sigma[j * a + j];
            p.fa1[j] = p.dt * sqp / sqv;
            p.fb1[j] = p.z_scale * sqp * p.c_pu * sg;
            p.fb2[j] = p.z_scale * sqv * p.c_vu * sg;
        }
    }
    p.trace = h->d_trace;
    p.x_inline = (h->n_ctrl == 1);
    if (p.x_inline) memcpy(p.x0, h->h_x, sizeof(float) * h->s);
    p.x = h->d_x;
    p.goal = h->d_goal;
    p.U = h->d_U;
    p.U_new = h->d_Unew;
    p.next = h->d_next;
    p.costs = h->d_costs;
    p.partials = h->d_partials;
    p.partials2 = h->d_partials2;
    p.max_groups = h->max_groups;
    p.payload = h->d_payload;
    p.stats = h->d_stats;
    p.counters = h->d_counters;
    p.eps = eps_dev;
    if (h->d_zc && h->zc_request) {
        p.next_host = h->d_zc;
        p.done_host = nullptr;
        p.done_epoch = h->zc_epoch;
    }
    p.peer_on = h->peer_on ? 1 : 0;
    p.rank = h->rank;
    p.epoch = h->epoch;
    if (h->peer_on) {
        for (int r = 0; r < h->world; r++) {
            p.peer_mail[r] = static_cast<float *>(h->peer_base[r]);
            p.peer_flag[r] = reinterpret_cast<uint32_t *>(static_cast<float *>(h->peer_base[r]) + h->mail_floats);
            p.peer_ll[r] = reinterpret_cast<uint2 *>(static_cast<char *>(h->peer_base[r]) + h->mail_ll_off);
        }
        p.peer_status = h->d_peer_status;
    }
    return p;
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 4); }
    template <class T> T *as() { return static_cast<T *>(p); }
};

}  // namespace

// =================================================================================================
extern "C" {

const char *mppi_version(void) { return "mppi_b200 0.1 sm_100a"; }

void mppi_config_default(mppi_config *cfg, int k, int tau, float dt, float mass, int s_dim, int a_dim)
{
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->k = k; cfg->tau = tau; cfg->dt = dt; cfg->mass = mass; cfg->s_dim = s_dim; cfg->a_dim = a_dim;
    cfg->lambda = 1.0f;       // src/controller_base.cpp:40
    cfg->seed = 1;            // RandomNormal::Seed(1), src/controller_base.cpp:199
    cfg->philox_rounds = 10;  // Philox4x32-10, the generator behind TensorFlow's RandomNormal
    cfg->device = -1;
    cfg->rank = 0; cfg->world = 1; cfg->n_controllers = 1;
}

const char *mppi_last_error(const mppi_handle *h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int mppi_create(const mppi_config *cfg, mppi_handle **out)
{
    if (!cfg || !out) return fail(nullptr, MPPI_ERR_BAD_ARG, "null cfg/out");
    *out = nullptr;
    const int n_ctrl = cfg->n_controllers > 0 ? cfg->n_controllers : 1;
    const int world = cfg->world > 0 ? cfg->world : 1;
    if (cfg->k <= 0 || cfg->tau <= 0 || cfg->a_dim <= 0 || cfg->s_dim <= 0)
        return fail(nullptr, MPPI_ERR_BAD_ARG, "k, tau, s_dim, a_dim must be positive");
    if (cfg->model != MPPI_MODEL_POINT_MASS && cfg->model != MPPI_MODEL_AUV)
        return fail(nullptr, MPPI_ERR_BAD_ARG, "cfg.model must be MPPI_MODEL_POINT_MASS or MPPI_MODEL_AUV (the MLP is selected by mppi_set_mlp)");
    const bool auv = cfg->model == MPPI_MODEL_AUV;
    if (auv && (cfg->s_dim != kAuvS || cfg->a_dim != kAuvA))
        return fail(nullptr, MPPI_ERR_BAD_ARG, "AUV model needs s_dim == 13, a_dim == 6");
    if (!auv && cfg->s_dim != 2 * cfg->a_dim)
        return fail(nullptr, MPPI_ERR_BAD_ARG, "point-mass model needs s_dim == 2 * a_dim");
    if (cfg->a_dim > MPPI_MAX_A) return fail(nullptr, MPPI_ERR_UNSUPPORTED, "a_dim > MPPI_MAX_A");
    if ((long long)cfg->tau * cfg->a_dim > MPPI_MAX_TA) return fail(nullptr, MPPI_ERR_UNSUPPORTED, "tau*a_dim > MPPI_MAX_TA");
    if (!(cfg->lambda > 0.f) || !(cfg->mass != 0.f)) return fail(nullptr, MPPI_ERR_BAD_ARG, "lambda must be > 0, mass != 0");
    if (cfg->rank < 0 || cfg->rank >= world) return fail(nullptr, MPPI_ERR_BAD_ARG, "rank outside [0, world)");
    if (cfg->k < world) return fail(nullptr, MPPI_ERR_BAD_ARG, "fewer samples than ranks");

    mppi_handle *h = new (std::nothrow) mppi_handle();
    if (!h) return fail(nullptr, MPPI_ERR_CUDA, "out of host memory");
    cudaDeviceProp prop;
    int rc = select_device(h, cfg->device, &h->device, &prop);
    if (rc != MPPI_OK) { g_last_error = h->err; delete h; return rc; }
    h->num_sms = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    h->smem_sm = prop.sharedMemPerMultiprocessor;
    if (cfg->philox_rounds != 0 && cfg->philox_rounds != 7 && cfg->philox_rounds != 10) {
        delete h;
        return fail(nullptr, MPPI_ERR_BAD_ARG, "philox_rounds must be 0 (= 10), 7 or 10");
    }
    h->rounds = cfg->philox_rounds ? cfg->philox_rounds : 10;
    if (const char *kv = getenv("MPPI_PHILOX_KERNEL")) {
        if (!strcmp(kv, "regen")) h->kernel_variant = 1;
        else if (!strcmp(kv, "resident")) h->kernel_variant = 2;
        else if (!strcmp(kv, "direct")) h->kernel_variant = 3;
    }

    h->k = cfg->k; h->tau = cfg->tau; h->s = cfg->s_dim; h->a = cfg->a_dim;
    h->n_ctrl = n_ctrl; h->rank = cfg->rank; h->world = world;
    h->dt = cfg->dt; h->mass = cfg->mass; h->lambda = cfg->lambda; h->seed = cfg->seed;
    h->TA = cfg->tau * cfg->a_dim;
    h->stride = partial_stride(h->TA);
    // rank r owns samples [r*k/world, (r+1)*k/world)  
    mppi_shard_range(cfg->k, cfg->rank, world, &h->k_offset, &h->K_local);
    h->goal_per_ctrl = cfg->goal_per_controller ? 1 : 0;
    h->auv = auv;
    memset(&h->auv_prm, 0, sizeof(h->auv_prm));

    const int a = h->a, s = h->s;
    memset(h->sigma, 0, sizeof(h->sigma));
    memset(h->q, 0, sizeof(h->q));
    for (int i = 0; i < a; i++)
        for (int j = 0; j < a; j++) h->sigma[i * a + j] = cfg->sigma ? cfg->sigma[i * a + j] : (i == j ? 1.f : 0.f);
    for (int i = 0; i < s; i++) h->q[i] = cfg->q ? cfg->q[i] : 1.f;
    for (int i = 0; i < s; i++)
        if (!(h->q[i] >= 0.f)) { delete h; return fail(nullptr, MPPI_ERR_BAD_ARG, "q must be non-negative (diag of a PSD Q)"); }
    {
        float inv[kMaxA * kMaxA];
        if (!invert_matrix(h->sigma, a, inv)) { delete h; return fail(nullptr, MPPI_ERR_BAD_ARG, "sigma is singular"); }
    }
    derive_sigma(h);

    auto bail = [&](int code, const std::string &msg) {
        g_last_error = msg;
        mppi_destroy(h);
        return code;
    };
#define CU_TRY_C(expr)                                                                             \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) return bail(MPPI_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

    if (cfg->stream) { h->stream = (cudaStream_t)cfg->stream; h->own_stream = false; }
    else { CU_TRY_C(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }

    h->max_gx = max_grid_x(h->K_local, n_ctrl, h->num_sms);
    const size_t n_goal = (size_t)(h->goal_per_ctrl ? n_ctrl : 1) * s;
    CU_TRY_C(cudaMalloc(&h->d_x, sizeof(float) * n_ctrl * s));
    CU_TRY_C(cudaMalloc(&h->d_goal, sizeof(float) * n_goal));
    CU_TRY_C(cudaMalloc(&h->d_U, sizeof(float) * n_ctrl * h->TA));
    CU_TRY_C(cudaMalloc(&h->d_Unew, sizeof(float) * n_ctrl * h->TA));
    CU_TRY_C(cudaMalloc(&h->d_next, sizeof(float) * n_ctrl * a));
    CU_TRY_C(cudaMalloc(&h->d_costs, sizeof(float) * (size_t)n_ctrl * h->K_local));
    h->max_groups = (h->max_gx + kMergeGroup - 1) / kMergeGroup;
    CU_TRY_C(cudaMalloc(&h->d_partials, sizeof(float) * (size_t)n_ctrl * h->max_gx * h->stride));
    CU_TRY_C(cudaMalloc(&h->d_partials2, sizeof(float) * (size_t)n_ctrl * h->max_groups * h->stride));
    CU_TRY_C(cudaMalloc(&h->d_payload, sizeof(float) * (size_t)n_ctrl * h->stride));
    CU_TRY_C(cudaMalloc(&h->d_gather, sizeof(float) * (size_t)world * n_ctrl * h->stride));
    CU_TRY_C(cudaMalloc(&h->d_stats, sizeof(float) * 2 * n_ctrl));
    CU_TRY_C(cudaMalloc(&h->d_norm, sizeof(float) * 2 * n_ctrl));
    CU_TRY_C(cudaMalloc(&h->d_cost_base, sizeof(float) * n_ctrl));
    CU_TRY_C(cudaMemset(h->d_cost_base, 0, sizeof(float) * n_ctrl));
    CU_TRY_C(cudaMalloc(&h->d_counters, sizeof(unsigned int) * (size_t)n_ctrl * (1 + h->max_groups)));
    CU_TRY_C(cudaMemset(h->d_counters, 0, sizeof(unsigned int) * (size_t)n_ctrl * (1 + h->max_groups)));
    CU_TRY_C(cudaMemset(h->d_U, 0, sizeof(float) * n_ctrl * h->TA));
    CU_TRY_C(cudaMemset(h->d_Unew, 0, sizeof(float) * n_ctrl * h->TA));
    CU_TRY_C(cudaMemset(h->d_stats, 0, sizeof(float) * 2 * n_ctrl));
    CU_TRY_C(cudaMemset(h->d_x, 0, sizeof(float) * n_ctrl * s));
    CU_TRY_C(cudaMallocHost(&h->h_xs[0], sizeof(float) * n_ctrl * s));
    CU_TRY_C(cudaMallocHost(&h->h_xs[1], sizeof(float) * n_ctrl * s));
    CU_TRY_C(cudaEventCreateWithFlags(&h->x_ev[0], cudaEventDisableTiming));
    CU_TRY_C(cudaEventCreateWithFlags(&h->x_ev[1], cudaEventDisableTiming));
    h->h_x = h->h_xs[0];
    CU_TRY_C(cudaMallocHost(&h->h_next, sizeof(float) * n_ctrl * a));
    if (n_ctrl == 1) {
        if (cudaHostAlloc(&h->h_zc, 128, cudaHostAllocMapped) == cudaSuccess &&
            cudaHostGetDevicePointer(&h->d_zc, h->h_zc, 0) == cudaSuccess) {
            memset(h->h_zc, 0, 128);
        } else {
            (void)cudaGetLastError();
            if (h->h_zc) cudaFreeHost(h->h_zc);
            h->h_zc = nullptr;
            h->d_zc = nullptr;
        }
    }
    memset(h->h_xs[0], 0, sizeof(float) * n_ctrl * s);
    memset(h->h_xs[1], 0, sizeof(float) * n_ctrl * s);

    std::vector<float> goal(n_goal);
    for (size_t i = 0; i < n_goal; i++)
        goal[i] = cfg->goal ? cfg->goal[i]
                  : auv     ? ((i % s) == 6 ? 1.f : 0.f)                         // origin, identity attitude, at rest
                            : ((i % s) % 2 == 0 ? 1.f : 0.f);                    // (1,0,1,0,..), :43-46
    CU_TRY_C(cudaMemcpy(h->d_goal, goal.data(), sizeof(float) * n_goal, cudaMemcpyHostToDevice));
#undef CU_TRY_C
    *out = h;
    return MPPI_OK;
}

int mppi_destroy(mppi_handle *h)
{
    if (!h) return MPPI_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    for (int r = 0; r < kMaxWorld; r++)
        if (h->peer_opened[r]) cudaIpcCloseMemHandle(h->peer_base[r]);
    cudaFree(h->d_mailbox); cudaFree(h->d_peer_status);
    if (h->h_peer_status) cudaFreeHost(h->h_peer_status);
    cudaFree(h->d_x); cudaFree(h->d_goal); cudaFree(h->d_U); cudaFree(h->d_Unew); cudaFree(h->d_next);
    cudaFree(h->d_costs); cudaFree(h->d_partials); cudaFree(h->d_partials2); cudaFree(h->d_stats); cudaFree(h->d_counters);
    cudaFree(h->d_norm);
    cudaFree(h->d_cost_base);
    cudaFree(h->d_trace);
    cudaFree(h->d_eps_tmp);
    cudaFree(h->d_wblob);
    cudaFree(h->d_nn_blob);
    cudaFree(h->d_fvec);
    cudaFree(h->d_params); cudaFree(h->d_adam_m); cudaFree(h->d_adam_v); cudaFree(h->d_train_work); cudaFree(h->d_loss);
    if (!h->ext_exchange) { cudaFree(h->d_payload); cudaFree(h->d_gather); }
    for (int i = 0; i < 2; i++) {
        if (h->h_xs[i]) cudaFreeHost(h->h_xs[i]);
        if (h->x_ev[i]) cudaEventDestroy(h->x_ev[i]);
    }
    if (h->h_next) cudaFreeHost(h->h_next);
    if (h->h_zc) cudaFreeHost(h->h_zc);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MPPI_OK;
}

int mppi_k_local(const mppi_handle *h) { return h ? h->K_local : 0; }
int mppi_k_offset(const mppi_handle *h) { return h ? h->k_offset : 0; }
int mppi_exchange_stride(const mppi_handle *h) { return h ? h->stride : 0; }
int mppi_payload_stride(int tau_times_a) { return tau_times_a > 0 ? partial_stride(tau_times_a) : 0; }
int mppi_shard_range(int k, int rank, int world, int *k_offset, int *k_local)
{
    if (k <= 0 || world <= 0 || rank < 0 || rank >= world || k < world || !k_offset || !k_local)
        return fail(nullptr, MPPI_ERR_BAD_ARG, "bad shard_range argument");
    const int lo = (int)((long long)k * rank / world), hi = (int)((long long)k * (rank + 1) / world);
    *k_offset = lo;
    *k_local = hi - lo;
    return MPPI_OK;
}

// ---- update step --------------------------------------------------------------------------------
int mppi_set_state(mppi_handle *h, const float *x_host)
{
    if (!h || !x_host) return fail(h, MPPI_ERR_BAD_ARG, "null handle/state");
    CU_TRY(h, cudaSetDevice(h->device));
    if (h->n_ctrl > 1) {
        const int i = h->x_idx;
        if (h->x_ev_armed[i]) CU_TRY(h, cudaEventSynchronize(h->x_ev[i]));     // the copy that last read this buffer is done
        h->h_x = h->h_xs[i];
        memcpy(h->h_x, x_host, sizeof(float) * h->n_ctrl * h->s);
        CU_TRY(h, cudaMemcpyAsync(h->d_x, h->h_x, sizeof(float) * h->n_ctrl * h->s, cudaMemcpyHostToDevice, h->stream));
        CU_TRY(h, cudaEventRecord(h->x_ev[i], h->stream));
        h->x_ev_armed[i] = true;
        h->x_idx ^= 1;
    } else {
        memcpy(h->h_x, x_host, sizeof(float) * h->s);      // travels by value in the kernel parameter block
    }
    h->x_staged = true;
    return MPPI_OK;
}

int mppi_enqueue_update(mppi_handle *h, const float *eps_dev)
{
    if (!h) return fail(h, MPPI_ERR_BAD_ARG, "null handle");
    if (!h->x_staged) return fail(h, MPPI_ERR_STATE, "mppi_set_state must precede mppi_enqueue_update");
    if (h->auv && !h->auv_ready) return fail(h, MPPI_ERR_STATE, "AUV handle: mppi_set_auv_model has not been called");
    CU_TRY(h, cudaSetDevice(h->device));
    if (h->d_zc && h->zc_request) { h->zc_epoch++; h->zc_armed = true; }
    RolloutParams p = make_params(h, eps_dev);
    int gx = 0;
    // cost normalisation (controller_base.py:468-474) needs max_k(S_k - beta) before any weight: two launches,
    // pass 1 = costs + global (min, max), pass 2 = weights and update from the stored costs
    const int npass = h->normalize ? 2 : 1;
    for (int pass = 1; pass <= npass; pass++) {
        p.norm_mode = h->normalize ? pass : 0;
        if (h->peer_on) p.epoch = ++h->epoch;       // one exchange per launch; the same count on every rank
        if (h->auv) {
            CU_TRY(h, launch_rollout_auv(p, h->auv_prm, eps_dev == nullptr, h->num_sms, h->stream, &gx));
        } else if (h->mlp) {
            MlpParams mp{h->d_wblob, h->d_fvec, h->s, h->a};
            CU_TRY(h, launch_rollout_mlp(p, mp, h->a, eps_dev == nullptr, h->num_sms, h->stream, &gx));
        } else if (eps_dev) {
            cudaError_t e = launch_rollout_injected(p, h->a, h->num_sms, h->smem_optin, h->stream, &gx);
            if (e == cudaErrorInvalidConfiguration)
                return fail(h, MPPI_ERR_UNSUPPORTED, "injected-noise mode: one 32-sample tile of tau*a_dim floats does not fit shared memory");
            CU_TRY(h, e);
        } else if (p.fast) {
            cudaError_t e = launch_rollout_philox_fast(p, h->a, h->kernel_variant, h->num_sms, h->smem_sm, h->smem_optin, h->stream, &gx);
            if (e == cudaErrorInvalidConfiguration)
                return fail(h, MPPI_ERR_UNSUPPORTED, "tau*a_dim is too large for the per-CTA shared-memory tables of the update kernel (about 2400 floats on this device)");
            CU_TRY(h, e);
        } else {
            cudaError_t e = launch_rollout_philox(p, h->a, h->num_sms, h->smem_optin, h->stream, &gx);
            if (e == cudaErrorInvalidConfiguration)
                return fail(h, MPPI_ERR_UNSUPPORTED, "tau*a_dim is too large for the per-CTA shared-memory tables of the update kernel (about 2400 floats on this device)");
            CU_TRY(h, e);
        }
    }
    h->last_philox = (eps_dev == nullptr);
    h->last_fast = p.fast != 0;
    h->last_gx = gx;
    if (!eps_dev) {
        h->have_philox_update = true;
        h->last_update = h->update_counter;
        h->update_counter++;
    }
    h->pending_finish = (h->world > 1) && !h->peer_on;
    return MPPI_OK;
}

int mppi_enqueue_finish(mppi_handle *h)
{
    if (!h) return fail(h, MPPI_ERR_BAD_ARG, "null handle");
    if (h->world <= 1 || h->peer_on) return MPPI_OK;
    if (!h->pending_finish) return fail(h, MPPI_ERR_STATE, "no update awaiting a finish");
    CU_TRY(h, cudaSetDevice(h->device));
    RolloutParams p = make_params(h, nullptr);
    CU_TRY(h, launch_finish(p, h->a, h->last_philox, h->d_gather, h->stream));
    h->pending_finish = false;
    return MPPI_OK;
}

static int exchange(mppi_handle *h)
{
    if (h->world <= 1 || h->peer_on) return MPPI_OK;
    if (!h->comm)
        return fail(h, MPPI_ERR_COMM, "world > 1 needs mppi_comm_init (in-library NCCL) or a caller-side all-gather "
                                      "between mppi_enqueue_update and mppi_enqueue_finish");
    const int rc = g_nccl.AllGather(h->d_payload, h->d_gather, (size_t)h->n_ctrl * h->stride, /*ncclFloat*/ 7, h->comm, h->stream);
    if (rc != 0)
        return fail(h, MPPI_ERR_COMM, std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    return MPPI_OK;
}

int mppi_enqueue_exchange(mppi_handle *h)
{
    if (!h) return fail(h, MPPI_ERR_BAD_ARG, "null handle");
    CU_TRY(h, cudaSetDevice(h->device));
    return exchange(h);
}

int mppi_fetch_action(mppi_handle *h, float *action_host)
{
    if (!h || !action_host) return fail(h, MPPI_ERR_BAD_ARG, "null handle/action");
    CU_TRY(h, cudaSetDevice(h->device));
    // zero-copy fast path: the kernel that applies the update has stored the action and then the update's epoch
    // into mapped pinned memory; wait for the word instead of copying and synchronising the stream
    const bool finish_in_kernel = (h->world <= 1) || h->peer_on;
    if (h->zc_armed && h->h_zc && finish_in_kernel && !h->pending_finish) {
        // slot i = {action_i, epoch} (i < a), slot 8 = {exchange status, epoch}: each an 8-byte store of the finishing CTA
        volatile unsigned int *zc = reinterpret_cast<volatile unsigned int *>(h->h_zc);
        auto all_in = [&]() {
            if (zc[2 * 8 + 1] != h->zc_epoch) return false;
            for (int i = 0; i < h->a; i++)
                if (zc[2 * i + 1] != h->zc_epoch) return false;
            return true;
        };
        bool seen = false;
        for (long long spin = 0; spin < (1LL << 34); spin++) {
            if (all_in()) { seen = true; break; }
            if ((spin & 0xfffff) == 0xfffff && cudaStreamQuery(h->stream) != cudaErrorNotReady) break;   // finished or failed
        }
        h->zc_armed = false;
        if (seen || all_in()) {
            std::atomic_thread_fence(std::memory_order_acquire);
            for (int i = 0; i < h->a; i++) {
                const unsigned int w = zc[2 * i];
                memcpy(action_host + i, &w, sizeof(float));
            }
            if (h->peer_on && zc[2 * 8] != 0u) {
                cudaMemsetAsync(h->d_peer_status, 0, sizeof(unsigned int), h->stream);    // reported: the next update starts clean
                return fail(h, MPPI_ERR_COMM, "fused exchange: a peer's payload did not arrive in time (is every rank calling the update?); the sequence was left unchanged");
            }
            return MPPI_OK;
        }
        CU_TRY(h, cudaStreamSynchronize(h->stream));    // surfaces the launch failure, if any; else fall through
    }
    h->zc_armed = false;
    CU_TRY(h, cudaMemcpyAsync(h->h_next, h->d_next, sizeof(float) * h->n_ctrl * h->a, cudaMemcpyDeviceToHost, h->stream));
    if (h->peer_on)
        CU_TRY(h, cudaMemcpyAsync(h->h_peer_status, h->d_peer_status, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    memcpy(action_host, h->h_next, sizeof(float) * h->n_ctrl * h->a);
    if (h->peer_on && *h->h_peer_status != 0u) {
        cudaMemsetAsync(h->d_peer_status, 0, sizeof(unsigned int), h->stream);            // reported: the next update starts clean
        return fail(h, MPPI_ERR_COMM, "fused exchange: a peer's payload did not arrive in time (is every rank calling the update?); the sequence was left unchanged");
    }
    return MPPI_OK;
}

int mppi_synchronize(mppi_handle *h)
{
    if (!h) return fail(h, MPPI_ERR_BAD_ARG, "null handle");
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return MPPI_OK;
}

static int next_common(mppi_handle *h, const float *x_host, const float *eps_dev, float *action_host)
{
    int rc = mppi_set_state(h, x_host);
    if (rc) return rc;
    h->zc_request = true;            // the caller waits for the action: let the kernel hand it over through mapped memory
    rc = mppi_enqueue_update(h, eps_dev);
    h->zc_request = false;
    if (rc) return rc;
    if (h->world > 1) {
        rc = exchange(h);
        if (rc) return rc;
        rc = mppi_enqueue_finish(h);
        if (rc) return rc;
    }
    return mppi_fetch_action(h, action_host);
}

int mppi_next(mppi_handle *h, const float *x_host, float *action_host)
{
    if (!h || !x_host || !action_host) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    return next_common(h, x_host, nullptr, action_host);
}

int mppi_next_with_noise_dev(mppi_handle *h, const float *x_host, const float *eps_dev, float *action_host)
{
    if (!h || !x_host || !eps_dev || !action_host) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    return next_common(h, x_host, eps_dev, action_host);
}

int mppi_next_with_noise(mppi_handle *h, const float *x_host, const float *eps_host, float *action_host)
{
    if (!h || !x_host || !eps_host || !action_host) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t bytes = sizeof(float) * (size_t)h->n_ctrl * h->K_local * h->TA;
    if (!h->d_eps_tmp) CU_TRY(h, cudaMalloc(&h->d_eps_tmp, bytes));
    CU_TRY(h, cudaMemcpyAsync(h->d_eps_tmp, eps_host, bytes, cudaMemcpyHostToDevice, h->stream));
    return next_common(h, x_host, h->d_eps_tmp, action_host);
}

// ---- controller state ------------------------------------------------------------------------------
static int d2h(mppi_handle *h, float *dst, const float *src, size_t n)
{
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaMemcpyAsync(dst, src, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return MPPI_OK;
}
static int h2d(mppi_handle *h, float *dst, const float *src, size_t n)
{
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaMemcpyAsync(dst, src, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return MPPI_OK;
}

int mppi_set_goal(mppi_handle *h, const float *goal_host)
{
    if (!h || !goal_host) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    return h2d(h, h->d_goal, goal_host, (size_t)(h->goal_per_ctrl ? h->n_ctrl : 1) * h->s);
}
int mppi_set_goal_n(mppi_handle *h, const float *goal_host, int n_rows)
{
    if (!h || !goal_host) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    if (n_rows == 1) {
        if (!h->goal_per_ctrl) return h2d(h, h->d_goal, goal_host, (size_t)h->s);
        std::vector<float> all((size_t)h->n_ctrl * h->s);
        for (int c = 0; c < h->n_ctrl; c++) memcpy(all.data() + (size_t)c * h->s, goal_host, sizeof(float) * h->s);
        return h2d(h, h->d_goal, all.data(), all.size());
    }
    if (n_rows == h->n_ctrl && h->goal_per_ctrl) return h2d(h, h->d_goal, goal_host, (size_t)h->n_ctrl * h->s);
    return fail(h, MPPI_ERR_BAD_ARG, "mppi_set_goal_n: n_rows must be 1, or n_controllers on a goal_per_controller handle");
}
int mppi_set_lambda(mppi_handle *h, float lambda)
{
    if (!h || !(lambda > 0.f)) return fail(h, MPPI_ERR_BAD_ARG, "lambda must be > 0");
    h->lambda = lambda;
    derive_sigma(h);
    return MPPI_OK;
}
int mppi_set_action_cost(mppi_handle *h, int form, float gamma, float upsilon)
{
    if (!h) return fail(h, MPPI_ERR_BAD_ARG, "null handle");
    if (form != MPPI_ACTION_COST_CPP && form != MPPI_ACTION_COST_PYTHON) return fail(h, MPPI_ERR_BAD_ARG, "unknown action-cost form");
    if (!(upsilon > 0.f)) return fail(h, MPPI_ERR_BAD_ARG, "upsilon must be > 0");
    h->cost_form = form;
    h->gamma = gamma;
    h->upsilon = upsilon;
    derive_sigma(h);
    return MPPI_OK;
}
static bool ellipse_params(float a, float b, float cx, float cy, float speed, float m_state, float m_vel, float *ell)
{
    if (!(a != 0.f) || !(b != 0.f)) return false;
    ell[0] = 1.0f / a; ell[1] = 1.0f / b; ell[2] = cx; ell[3] = cy; ell[4] = speed; ell[5] = m_state; ell[6] = m_vel; ell[7] = 0.f;
    return true;
}
int mppi_set_ellipse_cost(mppi_handle *h, float a, float b, float center_x, float center_y, float speed, float m_state,
                          float m_vel)
{
    if (!h) return fail(h, MPPI_ERR_BAD_ARG, "null handle");
    if (h->s != 4 || h->a != 2) return fail(h, MPPI_ERR_UNSUPPORTED, "ElipseCost is defined on the point_mass2d state (x, vx, y, vy): s_dim = 4, a_dim = 2");
    if (!ellipse_params(a, b, center_x, center_y, speed, m_state, m_vel, h->ell)) return fail(h, MPPI_ERR_BAD_ARG, "ellipse axes must be non-zero");
    h->cost_kind = 1;
    return MPPI_OK;
}
int mppi_set_static_cost(mppi_handle *h)
{
    if (!h) return fail(h, MPPI_ERR_BAD_ARG, "null handle");
    h->cost_kind = 0;
    return MPPI_OK;
}
int mppi_set_normalize_cost(mppi_handle *h, int on)
{
    if (!h) return fail(h, MPPI_ERR_BAD_ARG, "null handle");
    if (on && h->world > 1 && !h->peer_on)
        return fail(h, MPPI_ERR_UNSUPPORTED, "cost normalisation needs the global cost range before any weight: with world > 1 it "
                                              "runs over the fused peer-memory exchange only (mppi_peer_attach first)");
    h->normalize = on != 0;
    return MPPI_OK;
}
int mppi_set_action_limits(mppi_handle *h, int on, int n, const float *act_min, const float *act_max)
{
    if (!h) return fail(h, MPPI_ERR_BAD_ARG, "null handle");
    if (!on) { h->clip = false; return MPPI_OK; }
    if (!act_min || !act_max || (n != 1 && n != h->a)) return fail(h, MPPI_ERR_BAD_ARG, "limits: n must be 1 or a_dim");
    for (int j = 0; j < h->a; j++) {
        const float lo = act_min[n == 1 ? 0 : j], hi = act_max[n == 1 ? 0 : j];
        if (!(lo <= hi)) return fail(h, MPPI_ERR_BAD_ARG, "limits: act_min must not exceed act_max");
        h->act_min[j] = lo;
        h->act_max[j] = hi;
    }
    h->clip = true;
    return MPPI_OK;
}
int mppi_set_sigma(mppi_handle *h, const float *sigma_host)
{
    if (!h || !sigma_host) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    float inv[kMaxA * kMaxA];
    if (!invert_matrix(sigma_host, h->a, inv)) return fail(h, MPPI_ERR_BAD_ARG, "sigma is singular");
    memcpy(h->sigma, sigma_host, sizeof(float) * h->a * h->a);
    derive_sigma(h);
    return MPPI_OK;
}
int mppi_set_q(mppi_handle *h, const float *q_host)
{
    if (!h || !q_host) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    for (int i = 0; i < h->s; i++)
        if (!(q_host[i] >= 0.f)) return fail(h, MPPI_ERR_BAD_ARG, "q must be non-negative (diag of a PSD Q)");
    memcpy(h->q, q_host, sizeof(float) * h->s);
    return MPPI_OK;
}
int mppi_set_mass(mppi_handle *h, float mass)
{
    if (!h || !(mass != 0.f)) return fail(h, MPPI_ERR_BAD_ARG, "mass must be non-zero");
    h->mass = mass;
    return MPPI_OK;
}
int mppi_set_sequence(mppi_handle *h, const float *U_host)
{
    if (!h || !U_host) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    return h2d(h, h->d_U, U_host, (size_t)h->n_ctrl * h->TA);
}
int mppi_get_sequence(mppi_handle *h, float *U_host)
{
    if (!h || !U_host) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    return d2h(h, U_host, h->d_U, (size_t)h->n_ctrl * h->TA);
}
int mppi_get_update(mppi_handle *h, float *U_new_host)
{
    if (!h || !U_new_host) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    return d2h(h, U_new_host, h->d_Unew, (size_t)h->n_ctrl * h->TA);
}
// On the superposition path the device keeps S_k - C and beta - C (C: the cost part all samples of a controller share):
// it is added back here, in fp32, exactly as the device would have rounded C + (S_k - C).
static int cost_bases(mppi_handle *h, std::vector<float> &base)
{
    base.assign((size_t)h->n_ctrl, 0.f);
    if (!h->last_fast) return MPPI_OK;
    return d2h(h, base.data(), h->d_cost_base, base.size());
}
int mppi_get_costs(mppi_handle *h, float *costs_host)
{
    if (!h || !costs_host) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    int rc = d2h(h, costs_host, h->d_costs, (size_t)h->n_ctrl * h->K_local);
    if (rc) return rc;
    std::vector<float> base;
    rc = cost_bases(h, base);
    if (rc) return rc;
    if (h->last_fast)
        for (int c = 0; c < h->n_ctrl; c++) {
            float *row = costs_host + (size_t)c * h->K_local;
            for (int k = 0; k < h->K_local; k++) row[k] = base[c] + row[k];
        }
    return MPPI_OK;
}
int mppi_get_weight_stats(mppi_handle *h, float *beta, float *eta)
{
    if (!h || !beta || !eta) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    std::vector<float> st(2 * (size_t)h->n_ctrl);
    int rc = d2h(h, st.data(), h->d_stats, st.size());
    if (rc) return rc;
    std::vector<float> base;
    rc = cost_bases(h, base);
    if (rc) return rc;
    for (int c = 0; c < h->n_ctrl; c++) { beta[c] = base[c] + st[2 * c]; eta[c] = st[2 * c + 1]; }
    return MPPI_OK;
}
int mppi_set_update_counter(mppi_handle *h, uint32_t counter)
{
    if (!h) return fail(h, MPPI_ERR_BAD_ARG, "null handle");
    h->update_counter = counter;
    return MPPI_OK;
}

// Developer knob: %globaltimer stamps of the update kernel's phases (see RolloutParams::trace), [n_controllers][grid.x][12]
// nanoseconds of the LAST update.  mppi_debug_trace(h, 1) arms it (a few stores per CTA), (h, 0) disarms.
int mppi_debug_trace(mppi_handle *h, int on)
{
    if (!h) return fail(h, MPPI_ERR_BAD_ARG, "null handle");
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->d_trace);
    h->d_trace = nullptr;
    h->trace_ctas = 0;
    if (on) {
        h->trace_ctas = (size_t)h->n_ctrl * h->max_gx;
        CU_TRY(h, cudaMalloc(&h->d_trace, sizeof(unsigned long long) * kTraceSlots * h->trace_ctas));
        CU_TRY(h, cudaMemset(h->d_trace, 0, sizeof(unsigned long long) * kTraceSlots * h->trace_ctas));
    }
    return MPPI_OK;
}
int mppi_debug_get_trace(mppi_handle *h, unsigned long long *out, int n_ctas)
{
    if (!h || !out || n_ctas <= 0) return fail(h, MPPI_ERR_BAD_ARG, "bad argument");
    if (!h->d_trace || (size_t)n_ctas > h->trace_ctas) return fail(h, MPPI_ERR_STATE, "trace not armed (mppi_debug_trace) or too many CTAs asked for");
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaMemcpyAsync(out, h->d_trace, sizeof(unsigned long long) * kTraceSlots * n_ctas, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return MPPI_OK;
}
int mppi_last_grid_x(const mppi_handle *h) { return h ? h->last_gx : 0; }

int mppi_dump_noise(mppi_handle *h, float *eps_host)
{
    if (!h || !eps_host) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    if (!h->have_philox_update) return fail(h, MPPI_ERR_STATE, "no Philox-mode update has run yet");
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t n = (size_t)h->n_ctrl * h->K_local * h->TA;
    DevBuf buf;
    CU_TRY(h, buf.alloc(sizeof(float) * n));
    RolloutParams p = make_params(h, nullptr);
    p.update = h->last_update;
    CU_TRY(h, launch_dump_noise(p, h->a, buf.as<float>(), h->stream));
    return d2h(h, eps_host, buf.as<float>(), n);
}

// ---- exchange ------------------------------------------------------------------------------------
int mppi_exchange_buffers(mppi_handle *h, void **send_dev, void **recv_dev)
{
    if (!h || !send_dev || !recv_dev) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    *send_dev = h->d_payload;
    *recv_dev = h->d_gather;
    return MPPI_OK;
}
int mppi_exchange_set_buffers(mppi_handle *h, void *send_dev, void *recv_dev)
{
    if (!h || !send_dev || !recv_dev) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    CU_TRY(h, cudaSetDevice(h->device));
    if (!h->ext_exchange) { cudaFree(h->d_payload); cudaFree(h->d_gather); }
    h->d_payload = (float *)send_dev;
    h->d_gather = (float *)recv_dev;
    h->ext_exchange = true;
    return MPPI_OK;
}
static int peer_alloc(mppi_handle *h)
{
    if (h->d_mailbox) return MPPI_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    h->mail_floats = (size_t)2 * h->world * h->n_ctrl * h->stride;
    // [float records + flags: the (min, max) exchange of a normalised update] [tagged words: the payload exchange]
    h->mail_ll_off = (sizeof(float) * h->mail_floats + sizeof(uint32_t) * 2 * h->world * h->n_ctrl + 15) & ~(size_t)15;
    const size_t bytes = h->mail_ll_off + sizeof(uint2) * h->mail_floats;
    CU_TRY(h, cudaMalloc(&h->d_mailbox, bytes));
    CU_TRY(h, cudaMemset(h->d_mailbox, 0, bytes));
    CU_TRY(h, cudaMalloc(&h->d_peer_status, sizeof(unsigned int)));
    CU_TRY(h, cudaMemset(h->d_peer_status, 0, sizeof(unsigned int)));
    CU_TRY(h, cudaMallocHost(&h->h_peer_status, sizeof(unsigned int)));
    *h->h_peer_status = 0u;
    CU_TRY(h, cudaDeviceSynchronize());
    return MPPI_OK;
}
int mppi_peer_handle(mppi_handle *h, void *handle64)
{
    if (!h || !handle64) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    if (h->world < 2 || h->world > kMaxWorld) return fail(h, MPPI_ERR_UNSUPPORTED, "fused exchange needs 2 <= world <= MPPI_MAX_PEERS");
    static_assert(sizeof(cudaIpcMemHandle_t) == MPPI_PEER_HANDLE_BYTES, "IPC handle size");
    int rc = peer_alloc(h);
    if (rc) return rc;
    cudaIpcMemHandle_t ipc;
    CU_TRY(h, cudaIpcGetMemHandle(&ipc, h->d_mailbox));
    memcpy(handle64, &ipc, sizeof(ipc));
    return MPPI_OK;
}
int mppi_peer_attach(mppi_handle *h, const void *handles)
{
    if (!h || !handles) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    if (h->world < 2 || h->world > kMaxWorld) return fail(h, MPPI_ERR_UNSUPPORTED, "fused exchange needs 2 <= world <= MPPI_MAX_PEERS");
    int rc = peer_alloc(h);
    if (rc) return rc;
    CU_TRY(h, cudaSetDevice(h->device));
    for (int r = 0; r < h->world; r++) {
        if (r == h->rank) { h->peer_base[r] = h->d_mailbox; continue; }
        if (h->peer_opened[r]) continue;
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, static_cast<const char *>(handles) + (size_t)r * MPPI_PEER_HANDLE_BYTES, sizeof(ipc));
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            return fail(h, MPPI_ERR_COMM, std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
        }
        h->peer_base[r] = ptr;
        h->peer_opened[r] = true;
    }
    if (!h->peer_on) h->epoch = 0;      // a re-attach keeps counting: flags of earlier epochs can never match a later one
    h->peer_on = true;
    return MPPI_OK;
}
int mppi_comm_unique_id(void *id128)
{
    if (!id128) return fail(nullptr, MPPI_ERR_BAD_ARG, "null id");
    std::string err;
    if (!load_nccl(err)) return fail(nullptr, MPPI_ERR_COMM, err);
    NcclUniqueId id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) return fail(nullptr, MPPI_ERR_COMM, "ncclGetUniqueId failed");
    memcpy(id128, &id, sizeof(id));
    return MPPI_OK;
}
int mppi_comm_init(mppi_handle *h, const void *id128)
{
    if (!h || !id128) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    std::string err;
    if (!load_nccl(err)) return fail(h, MPPI_ERR_COMM, err);
    CU_TRY(h, cudaSetDevice(h->device));
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    const int rc = g_nccl.CommInitRank(&h->comm, h->world, id, h->rank);
    if (rc != 0)
        return fail(h, MPPI_ERR_COMM, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    return MPPI_OK;
}

// ---- AUV (Fossen) model and quaternion cost (rows N3 / N4) ------------------------------------------
static bool invert6(const double *m, double *inv)
{
    double w[6][12];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) { w[i][j] = m[i * 6 + j]; w[i][6 + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 6; c++) {
        int piv = c;
        for (int r = c + 1; r < 6; r++)
            if (fabs(w[r][c]) > fabs(w[piv][c])) piv = r;
        if (w[piv][c] == 0.0) return false;
        if (piv != c)
            for (int j = 0; j < 12; j++) { const double t = w[c][j]; w[c][j] = w[piv][j]; w[piv][j] = t; }
        const double d = w[c][c];
        for (int j = 0; j < 12; j++) w[c][j] /= d;
        for (int r = 0; r < 6; r++) {
            if (r == c) continue;
            const double f = w[r][c];
            for (int j = 0; j < 12; j++) w[r][j] -= f * w[c][j];
        }
    }
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) inv[i * 6 + j] = w[i][6 + j];
    return true;
}

// Derived constants of AUVModel.__init__ (auv_model.py:146-283), in double, rounded once to fp32.
static bool derive_auv(const mppi_auv_params *prm, float dt, AuvParams *out)
{
    const double m = prm->mass, g = 9.81;
    const float *cog = prm->cog, *in = prm->inertia;
    // tf_skew_op (:23-40) stacks its rows as columns: the TRANSPOSED skew matrix of cog, kept as the reference has it
    const double S[9] = {0, cog[2], -cog[1], -cog[2], 0, cog[0], cog[1], -cog[0], 0};
    const double I[9] = {in[0], in[3], in[4], in[3], in[1], in[5], in[4], in[5], in[2]};
    double M[36], Minv[36];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            M[r * 6 + c] = (r == c) ? m : 0.0;
            M[r * 6 + 3 + c] = -(m * S[r * 3 + c]);
            M[(3 + r) * 6 + c] = m * S[r * 3 + c];
            M[(3 + r) * 6 + 3 + c] = I[r * 3 + c];
        }
    for (int i = 0; i < 36; i++) M[i] += prm->added_mass[i];
    if (!invert6(M, Minv)) return false;
    memset(out, 0, sizeof(*out));
    out->fng = (float)(-(m * g));
    out->fnb = (float)((double)prm->volume * (double)prm->density * g);
    for (int i = 0; i < 3; i++) { out->cog[i] = prm->cog[i]; out->cob[i] = prm->cob[i]; }
    for (int i = 0; i < 36; i++) {
        out->Mtot[i] = (float)M[i];
        out->invM[i] = (float)Minv[i];
        out->Dl[i] = prm->linear_damping[i];
        out->Dlf[i] = prm->linear_damping_forward_speed[i];
    }
    for (int i = 0; i < 6; i++) out->dq[i] = prm->quad_damping[i];
    out->dt = dt;
    out->rk = prm->rk;
    return true;
}

int mppi_set_auv_model(mppi_handle *h, const mppi_auv_params *prm)
{
    if (!h || !prm) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    if (!h->auv) return fail(h, MPPI_ERR_STATE, "handle was not created with cfg.model = MPPI_MODEL_AUV");
    if (prm->rk != 1 && prm->rk != 2 && prm->rk != 4) return fail(h, MPPI_ERR_BAD_ARG, "rk must be 1, 2 or 4");
    if (!derive_auv(prm, h->dt, &h->auv_prm)) return fail(h, MPPI_ERR_BAD_ARG, "total mass matrix is singular");
    h->auv_ready = true;
    return MPPI_OK;
}

int mppi_set_nn_auv_model(mppi_handle *h, int n_hidden, int hidden, const float *const *W, const float *const *b, const float *Xmean,
                          const float *Xstd, const float *Ymean, const float *Ystd)
{
    if (!h || !W || !b) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    if (!h->auv) return fail(h, MPPI_ERR_STATE, "handle was not created with cfg.model = MPPI_MODEL_AUV");
    if (n_hidden < 1 || n_hidden > kNnMaxHidden || hidden < 1 || hidden > kNnW)
        return fail(h, MPPI_ERR_UNSUPPORTED, "the learned AUV model takes 1..4 hidden layers of 1..32 units");
    for (int l = 0; l <= n_hidden; l++)
        if (!W[l] || !b[l]) return fail(h, MPPI_ERR_BAD_ARG, "null layer");
    CU_TRY(h, cudaSetDevice(h->device));
    std::vector<float> blob((size_t)nn_auv_blob_floats(n_hidden), 0.f);
    for (int i = 0; i < 16; i++) {
        blob[i] = Xmean ? Xmean[i] : 0.f;
        const float sd = Xstd ? Xstd[i] : 1.f;
        if (!(sd != 0.f)) return fail(h, MPPI_ERR_BAD_ARG, "Xstd must be non-zero");
        blob[16 + i] = 1.0f / sd;
    }
    for (int i = 0; i < kAuvS; i++) { blob[32 + i] = Ystd ? Ystd[i] : 1.f; blob[48 + i] = Ymean ? Ymean[i] : 0.f; }
    // zero padding of narrower layers is exact: a padded unit has zero weights and bias, outputs relu(0) = 0 and feeds nothing on
    float *dst = blob.data() + 64;
    int n_in = kNnIn, w_in = kNnIn;                           // logical / padded input width of the layer
    for (int l = 0; l <= n_hidden; l++) {
        const int n_out = (l == n_hidden) ? kAuvS : hidden, w_out = (l == n_hidden) ? kNnOutPad : kNnW;
        for (int q = 0; q < n_in; q++)
            for (int o = 0; o < n_out; o++) dst[(size_t)q * w_out + o] = W[l][(size_t)q * n_out + o];
        for (int o = 0; o < n_out; o++) dst[(size_t)w_in * w_out + o] = b[l][o];
        dst += (size_t)w_in * w_out + w_out;
        n_in = n_out;
        w_in = w_out;
    }
    if (h->d_nn_blob) { cudaFree(h->d_nn_blob); h->d_nn_blob = nullptr; }
    CU_TRY(h, cudaMalloc(&h->d_nn_blob, sizeof(float) * blob.size()));
    CU_TRY(h, cudaMemcpyAsync(h->d_nn_blob, blob.data(), sizeof(float) * blob.size(), cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    memset(&h->auv_prm, 0, sizeof(h->auv_prm));
    h->auv_prm.dt = h->dt;
    h->auv_prm.rk = 0;
    h->auv_prm.nn_blob = h->d_nn_blob;
    h->auv_prm.nn_hidden = n_hidden;
    h->auv_ready = true;
    return MPPI_OK;
}

int mppi_auv_predict(mppi_handle *h, int kst, int k, const float *state, const float *action, float *out)
{
    if (!h || !state || !action || !out || k <= 0 || (kst != 1 && kst != k)) return fail(h, MPPI_ERR_BAD_ARG, "bad auv_predict argument");
    if (!h->auv || !h->auv_ready) return fail(h, MPPI_ERR_STATE, "mppi_set_auv_model has not been called");
    CU_TRY(h, cudaSetDevice(h->device));
    DevBuf ds, da, dout;
    CU_TRY(h, ds.alloc(sizeof(float) * (size_t)kst * kAuvS));
    CU_TRY(h, da.alloc(sizeof(float) * (size_t)k * kAuvA));
    CU_TRY(h, dout.alloc(sizeof(float) * (size_t)k * kAuvS));
    CU_TRY(h, cudaMemcpyAsync(ds.p, state, sizeof(float) * (size_t)kst * kAuvS, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(da.p, action, sizeof(float) * (size_t)k * kAuvA, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, launch_auv_predict(h->auv_prm, kst, k, ds.as<float>(), da.as<float>(), dout.as<float>(), h->stream));
    CU_TRY(h, cudaMemcpyAsync(out, dout.p, sizeof(float) * (size_t)k * kAuvS, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return MPPI_OK;
}

int mppi_set_quat_cost(mppi_handle *h, const float *q10)
{
    if (!h || !q10) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    if (!h->auv) return fail(h, MPPI_ERR_UNSUPPORTED, "StaticQuatCost is defined on the 13-dimensional AUV state");
    for (int i = 0; i < 10; i++)
        if (!(q10[i] >= 0.f)) return fail(h, MPPI_ERR_BAD_ARG, "q must be non-negative (diag of a PSD Q)");
    memset(h->q10, 0, sizeof(h->q10));
    memcpy(h->q10, q10, sizeof(float) * 10);
    h->cost_kind = 2;
    return MPPI_OK;
}

int mppi_cost_state_quat(int device, int k, const float *state, const float *goal, const float *q10, float *out)
{
    if (k <= 0 || !state || !goal || !q10 || !out) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad cost_state_quat argument");
    cudaDeviceProp prop;
    int dev = 0;
    int rc = select_device(nullptr, device, &dev, &prop);
    if (rc != MPPI_OK) return rc;
    float q[kAuvS] = {0};
    memcpy(q, q10, sizeof(float) * 10);
    DevBuf ds, dout;
    CU_TRY(nullptr, ds.alloc(sizeof(float) * (size_t)k * kAuvS));
    CU_TRY(nullptr, dout.alloc(sizeof(float) * (size_t)k));
    CU_TRY(nullptr, cudaMemcpy(ds.p, state, sizeof(float) * (size_t)k * kAuvS, cudaMemcpyHostToDevice));
    CU_TRY(nullptr, launch_auv_cost(2, k, q, goal, nullptr, ds.as<float>(), dout.as<float>(), nullptr));
    CU_TRY(nullptr, cudaMemcpy(out, dout.p, sizeof(float) * (size_t)k, cudaMemcpyDeviceToHost));
    return MPPI_OK;
}

// ElipseCost3D.prepare_consts (elipse_cost.py:137-154) on the host: N = [aVec, normal x aVec, normal] (columns),
// R = inv(N)^T, q = from_rotation_matrix(R) (tensorflow_graphics' four-branch trace method), then the kernel constants.
static bool ellipse3d_params(const float *normal, const float *a_vec, const float *axis, float speed, float m_state, float m_vel,
                             float *ell)
{
    if (!(axis[0] != 0.f) || !(axis[1] != 0.f)) return false;
    const double n[3] = {normal[0], normal[1], normal[2]}, a[3] = {a_vec[0], a_vec[1], a_vec[2]};
    const double b[3] = {n[1] * a[2] - n[2] * a[1], n[2] * a[0] - n[0] * a[2], n[0] * a[1] - n[1] * a[0]};
    const double N[9] = {a[0], b[0], n[0], a[1], b[1], n[1], a[2], b[2], n[2]};
    const double det = N[0] * (N[4] * N[8] - N[5] * N[7]) - N[1] * (N[3] * N[8] - N[5] * N[6]) + N[2] * (N[3] * N[7] - N[4] * N[6]);
    if (det == 0.0) return false;
    double inv[9];
    inv[0] = (N[4] * N[8] - N[5] * N[7]) / det; inv[1] = (N[2] * N[7] - N[1] * N[8]) / det; inv[2] = (N[1] * N[5] - N[2] * N[4]) / det;
    inv[3] = (N[5] * N[6] - N[3] * N[8]) / det; inv[4] = (N[0] * N[8] - N[2] * N[6]) / det; inv[5] = (N[2] * N[3] - N[0] * N[5]) / det;
    inv[6] = (N[3] * N[7] - N[4] * N[6]) / det; inv[7] = (N[1] * N[6] - N[0] * N[7]) / det; inv[8] = (N[0] * N[4] - N[1] * N[3]) / det;
    double R[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) R[r * 3 + c] = inv[c * 3 + r];
    double q[4], sq;
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0.0) {
        sq = sqrt(tr + 1.0) * 2.0;
        q[0] = (R[7] - R[5]) / sq; q[1] = (R[2] - R[6]) / sq; q[2] = (R[3] - R[1]) / sq; q[3] = 0.25 * sq;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        sq = sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0;
        q[0] = 0.25 * sq; q[1] = (R[1] + R[3]) / sq; q[2] = (R[2] + R[6]) / sq; q[3] = (R[7] - R[5]) / sq;
    } else if (R[4] > R[8]) {
        sq = sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0;
        q[0] = (R[1] + R[3]) / sq; q[1] = 0.25 * sq; q[2] = (R[5] + R[7]) / sq; q[3] = (R[2] - R[6]) / sq;
    } else {
        sq = sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0;
        q[0] = (R[2] + R[6]) / sq; q[1] = (R[5] + R[7]) / sq; q[2] = 0.25 * sq; q[3] = (R[3] - R[1]) / sq;
    }
    for (int i = 0; i < 4; i++) {
        if (!(q[i] == q[i])) return false;
        ell[i] = (float)q[i];
    }
    const double ea = axis[0], eb = axis[1];
    ell[4] = (float)(1.0 / ea); ell[5] = (float)(1.0 / eb); ell[6] = (float)(-ea / eb); ell[7] = (float)(eb / ea);
    ell[8] = (float)((double)speed * speed); ell[9] = m_state; ell[10] = m_vel; ell[11] = 0.f;
    return true;
}

int mppi_set_ellipse3d_cost(mppi_handle *h, const float *normal, const float *a_vec, const float *axis, const float *center,
                            float speed, float m_state, float m_vel)
{
    if (!h || !normal || !a_vec || !axis) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    (void)center;                 // stored by the reference (self.t) and never used by its state_cost
    if (!h->auv) return fail(h, MPPI_ERR_UNSUPPORTED, "ElipseCost3D is defined on the 13-dimensional AUV state");
    if (!ellipse3d_params(normal, a_vec, axis, speed, m_state, m_vel, h->ell))
        return fail(h, MPPI_ERR_BAD_ARG, "ellipse axes must be non-zero and (aVec, normal x aVec, normal) must be a basis");
    h->cost_kind = 3;
    return MPPI_OK;
}

int mppi_cost_state_ellipse3d(int device, int k, const float *state, const float *normal, const float *a_vec, const float *axis,
                              const float *center, float speed, float m_state, float m_vel, float *out)
{
    if (k <= 0 || !state || !normal || !a_vec || !axis || !out) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad cost_state_ellipse3d argument");
    (void)center;
    float ell[12];
    if (!ellipse3d_params(normal, a_vec, axis, speed, m_state, m_vel, ell))
        return fail(nullptr, MPPI_ERR_BAD_ARG, "ellipse axes must be non-zero and (aVec, normal x aVec, normal) must be a basis");
    cudaDeviceProp prop;
    int dev = 0;
    int rc = select_device(nullptr, device, &dev, &prop);
    if (rc != MPPI_OK) return rc;
    DevBuf ds, dout;
    CU_TRY(nullptr, ds.alloc(sizeof(float) * (size_t)k * kAuvS));
    CU_TRY(nullptr, dout.alloc(sizeof(float) * (size_t)k));
    CU_TRY(nullptr, cudaMemcpy(ds.p, state, sizeof(float) * (size_t)k * kAuvS, cudaMemcpyHostToDevice));
    CU_TRY(nullptr, launch_auv_cost(3, k, nullptr, nullptr, ell, ds.as<float>(), dout.as<float>(), nullptr));
    CU_TRY(nullptr, cudaMemcpy(out, dout.p, sizeof(float) * (size_t)k, cudaMemcpyDeviceToHost));
    return MPPI_OK;
}

int mppi_set_mlp(mppi_handle *h, int hidden, const float *W1, const float *b1, const float *W2, const float *b2,
                 const float *W3, const float *b3, const float *Xmean, const float *Xstd, const float *Ymean,
                 const float *Ystd)
{
    if (!h || !W1 || !b1 || !W2 || !b2 || !W3 || !b3) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    if (h->auv) return fail(h, MPPI_ERR_UNSUPPORTED, "mppi_set_mlp needs a point-mass handle (s_dim = 2 a_dim)");
    if (hidden < 1 || hidden > kMlpH) return fail(h, MPPI_ERR_UNSUPPORTED, "hidden must be in 1..128");
    if (h->s + h->a + 1 > kMlpKin || h->a > 5) return fail(h, MPPI_ERR_UNSUPPORTED, "MLP path needs s + a <= 15 (a <= 5)");
    CU_TRY(h, cudaSetDevice(h->device));
    // Narrower networks (the reference's models use 32 units, nn_model.py:54-60) run on the 128-wide tensor-core tiles with
    // zero padding: a padded unit has zero weights and bias, so it outputs relu(0) = 0, feeds nothing forward, and its
    // gradients are exactly zero - the padded network IS the narrow one, in the rollout and under training.
    std::vector<float> pW1, pb1, pW2, pb2, pW3;
    if (hidden != kMlpH) {
        const int in_ = h->s + h->a, H = hidden;
        pW1.assign((size_t)in_ * kMlpH, 0.f); pb1.assign(kMlpH, 0.f); pW2.assign((size_t)kMlpH * kMlpH, 0.f);
        pb2.assign(kMlpH, 0.f); pW3.assign((size_t)kMlpH * h->s, 0.f);
        for (int i = 0; i < in_; i++) memcpy(&pW1[(size_t)i * kMlpH], W1 + (size_t)i * H, sizeof(float) * H);
        memcpy(pb1.data(), b1, sizeof(float) * H);
        for (int i = 0; i < H; i++) memcpy(&pW2[(size_t)i * kMlpH], W2 + (size_t)i * H, sizeof(float) * H);
        memcpy(pb2.data(), b2, sizeof(float) * H);
        memcpy(pW3.data(), W3, sizeof(float) * (size_t)H * h->s);
        W1 = pW1.data(); b1 = pb1.data(); W2 = pW2.data(); b2 = pb2.data(); W3 = pW3.data();
    }
    h->mlp_hidden = hidden;
    std::vector<uint8_t> blob(kWBlobBytes);
    mlp_pack_weights(h->s, h->a, W1, b1, W2, b2, W3, b3, blob.data());
    std::vector<float> fv(kFvecFloats, 0.f);
    const int in = h->s + h->a;
    for (int i = 0; i < 16; i++) { fv[16 + i] = 1.f; fv[32 + i] = 1.f; }
    for (int i = 0; i < in; i++) {
        fv[i] = Xmean ? Xmean[i] : 0.f;
        const float sd = Xstd ? Xstd[i] : 1.f;
        if (!(sd != 0.f)) return fail(h, MPPI_ERR_BAD_ARG, "Xstd must be non-zero");
        fv[16 + i] = 1.0f / sd;
    }
    for (int i = 0; i < h->s; i++) {
        fv[32 + i] = Ystd ? Ystd[i] : 1.f;
        fv[48 + i] = Ymean ? Ymean[i] : 0.f;
    }
    if (!h->d_wblob) CU_TRY(h, cudaMalloc(&h->d_wblob, kWBlobBytes));
    if (!h->d_fvec) CU_TRY(h, cudaMalloc(&h->d_fvec, sizeof(float) * kFvecFloats));
    CU_TRY(h, cudaMemcpyAsync(h->d_wblob, blob.data(), kWBlobBytes, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(h->d_fvec, fv.data(), sizeof(float) * kFvecFloats, cudaMemcpyHostToDevice, h->stream));
    // fp32 master copy for the learner, Adam state reset
    const size_t np = mlp_param_count(h->s, h->a);
    std::vector<float> params;
    params.reserve(np);
    params.insert(params.end(), W1, W1 + (size_t)in * kMlpH);
    params.insert(params.end(), b1, b1 + kMlpH);
    params.insert(params.end(), W2, W2 + (size_t)kMlpH * kMlpH);
    params.insert(params.end(), b2, b2 + kMlpH);
    params.insert(params.end(), W3, W3 + (size_t)kMlpH * h->s);
    params.insert(params.end(), b3, b3 + h->s);
    if (!h->d_params) {
        CU_TRY(h, cudaMalloc(&h->d_params, sizeof(float) * np));
        CU_TRY(h, cudaMalloc(&h->d_adam_m, sizeof(float) * np));
        CU_TRY(h, cudaMalloc(&h->d_adam_v, sizeof(float) * np));
        CU_TRY(h, cudaMalloc(&h->d_loss, sizeof(float)));
    }
    CU_TRY(h, cudaMemcpyAsync(h->d_params, params.data(), sizeof(float) * np, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemsetAsync(h->d_adam_m, 0, sizeof(float) * np, h->stream));
    CU_TRY(h, cudaMemsetAsync(h->d_adam_v, 0, sizeof(float) * np, h->stream));
    h->adam_t = 0;
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    h->mlp = true;
    return MPPI_OK;
}

int mppi_mlp_set_adam(mppi_handle *h, float beta1, float beta2, float epsilon)
{
    if (!h) return fail(h, MPPI_ERR_BAD_ARG, "null handle");
    if (!h->mlp) return fail(h, MPPI_ERR_STATE, "mppi_set_mlp has not been called");
    if (!(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f) || !(epsilon > 0.f))
        return fail(h, MPPI_ERR_BAD_ARG, "Adam needs 0 <= beta < 1 and epsilon > 0");
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t np = mlp_param_count(h->s, h->a);
    h->adam_b1 = beta1; h->adam_b2 = beta2; h->adam_eps = epsilon;
    h->adam_t = 0;
    CU_TRY(h, cudaMemsetAsync(h->d_adam_m, 0, sizeof(float) * np, h->stream));
    CU_TRY(h, cudaMemsetAsync(h->d_adam_v, 0, sizeof(float) * np, h->stream));
    return MPPI_OK;
}

int mppi_mlp_train_step(mppi_handle *h, int n, const float *state, const float *action, const float *next_state,
                        float learning_rate, float *loss_out)
{
    if (!h || !state || !action || !next_state || n <= 0) return fail(h, MPPI_ERR_BAD_ARG, "bad train_step argument");
    if (!h->mlp) return fail(h, MPPI_ERR_STATE, "mppi_set_mlp has not been called");
    CU_TRY(h, cudaSetDevice(h->device));
    const int s = h->s, a = h->a;
    const size_t need = train_work_floats(s, a, n) + (size_t)n * (2 * s + a);
    if (need > h->train_work_cap) {
        cudaFree(h->d_train_work);
        h->d_train_work = nullptr;
        h->train_work_cap = 0;
        CU_TRY(h, cudaMalloc(&h->d_train_work, sizeof(float) * need));
        h->train_work_cap = need;
    }
    float *dx = h->d_train_work, *du = dx + (size_t)n * s, *dxn = du + (size_t)n * a, *work = dxn + (size_t)n * s;
    CU_TRY(h, cudaMemcpyAsync(dx, state, sizeof(float) * (size_t)n * s, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(du, action, sizeof(float) * (size_t)n * a, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(dxn, next_state, sizeof(float) * (size_t)n * s, cudaMemcpyHostToDevice, h->stream));
    h->adam_t++;
    const double t = (double)h->adam_t;
    const float lr_t = (float)((double)learning_rate * sqrt(1.0 - pow((double)h->adam_b2, t)) / (1.0 - pow((double)h->adam_b1, t)));
    CU_TRY(h, launch_train_step(s, a, n, dx, du, dxn, h->d_fvec, h->d_params, h->d_adam_m, h->d_adam_v, lr_t, h->adam_b1,
                                h->adam_b2, h->adam_eps, work, h->d_loss, h->stream));
    CU_TRY(h, launch_pack_blob(s, a, h->d_params, h->d_wblob, h->stream));     // the rollout sees the new weights
    float loss = 0.f;
    CU_TRY(h, cudaMemcpyAsync(&loss, h->d_loss, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    if (loss_out) *loss_out = loss;
    return MPPI_OK;
}

// The learner's loop (LearnerBase.train, learner_base.py:324-358): `epochs` times { optional augment_data (:455-467), then
// Adam steps on the normalised MSE (_train_step, :469-496) }.  batch_size <= 0 is the reference's behaviour - ONE full-batch
// step per epoch (its batchSize argument is accepted and never used, :146-153,470); batch_size > 0 walks the epoch's data
// in consecutive minibatches (an extension).  The data stay on the device for the whole call; losses_out (may be NULL)
// receives the loss before every step, [epochs * steps_per_epoch].
int mppi_mlp_train(mppi_handle *h, int n, const float *state, const float *action, const float *next_state, int epochs,
                   int batch_size, float learning_rate, int augment_samples, float augment_sigma, uint64_t seed, float *losses_out,
                   int *n_steps_out)
{
    if (!h || !state || !action || !next_state || n <= 0 || epochs <= 0) return fail(h, MPPI_ERR_BAD_ARG, "bad mlp_train argument");
    if (!h->mlp) return fail(h, MPPI_ERR_STATE, "mppi_set_mlp has not been called");
    if (augment_samples < 0 || augment_sigma < 0.f) return fail(h, MPPI_ERR_BAD_ARG, "augmentation needs samples >= 0, sigma >= 0");
    CU_TRY(h, cudaSetDevice(h->device));
    const int s = h->s, a = h->a;
    const bool aug = augment_samples > 0;
    const long long n_ep = aug ? (long long)n * augment_samples : n;
    if (n_ep > (1LL << 28)) return fail(h, MPPI_ERR_UNSUPPORTED, "epoch data too large");
    const int bs = (batch_size > 0 && batch_size < n_ep) ? batch_size : (int)n_ep;
    const int steps_per_epoch = (int)((n_ep + bs - 1) / bs);
    const size_t need = train_work_floats(s, a, bs);
    if (need > h->train_work_cap) {
        cudaFree(h->d_train_work);
        h->d_train_work = nullptr;
        h->train_work_cap = 0;
        CU_TRY(h, cudaMalloc(&h->d_train_work, sizeof(float) * need));
        h->train_work_cap = need;
    }
    DevBuf raw, augb, losses;
    const size_t row = (size_t)(2 * s + a);
    CU_TRY(h, raw.alloc(sizeof(float) * row * n));
    if (aug) CU_TRY(h, augb.alloc(sizeof(float) * row * (size_t)n_ep));
    CU_TRY(h, losses.alloc(sizeof(float) * (size_t)epochs * steps_per_epoch));
    float *dx = raw.as<float>(), *du = dx + (size_t)n * s, *dxn = du + (size_t)n * a;
    CU_TRY(h, cudaMemcpyAsync(dx, state, sizeof(float) * (size_t)n * s, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(du, action, sizeof(float) * (size_t)n * a, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(dxn, next_state, sizeof(float) * (size_t)n * s, cudaMemcpyHostToDevice, h->stream));
    float *ex = dx, *eu = du, *exn = dxn;
    if (aug) {
        ex = augb.as<float>();
        eu = ex + (size_t)n_ep * s;
        exn = eu + (size_t)n_ep * a;
    }
    for (int e = 0; e < epochs; e++) {
        if (aug)
            CU_TRY(h, launch_augment(n, augment_samples, s, a, dx, du, dxn, h->d_fvec, augment_sigma, seed, (uint32_t)e, ex, eu, exn, h->stream));
        for (int b = 0; b < steps_per_epoch; b++) {
            const long long lo = (long long)b * bs;
            const int nb = (int)((n_ep - lo) < bs ? (n_ep - lo) : bs);
            h->adam_t++;
            const double t = (double)h->adam_t;
            const float lr_t = (float)((double)learning_rate * sqrt(1.0 - pow((double)h->adam_b2, t)) / (1.0 - pow((double)h->adam_b1, t)));
            CU_TRY(h, launch_train_step(s, a, nb, ex + lo * s, eu + lo * a, exn + lo * s, h->d_fvec, h->d_params, h->d_adam_m, h->d_adam_v,
                                        lr_t, h->adam_b1, h->adam_b2, h->adam_eps, h->d_train_work,
                                        losses.as<float>() + (size_t)e * steps_per_epoch + b, h->stream));
        }
    }
    CU_TRY(h, launch_pack_blob(s, a, h->d_params, h->d_wblob, h->stream));     // the rollout sees the new weights
    if (n_steps_out) *n_steps_out = epochs * steps_per_epoch;
    if (losses_out) return d2h(h, losses_out, losses.as<float>(), (size_t)epochs * steps_per_epoch);
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return MPPI_OK;
}

int mppi_mlp_get_weights(mppi_handle *h, float *W1, float *b1, float *W2, float *b2, float *W3, float *b3)
{
    if (!h || !W1 || !b1 || !W2 || !b2 || !W3 || !b3) return fail(h, MPPI_ERR_BAD_ARG, "null argument");
    if (!h->mlp) return fail(h, MPPI_ERR_STATE, "mppi_set_mlp has not been called");
    const int in = h->s + h->a;
    std::vector<float> params(mlp_param_count(h->s, h->a));
    int rc = d2h(h, params.data(), h->d_params, params.size());
    if (rc) return rc;
    const float *p = params.data();
    const int H = h->mlp_hidden;                    // the caller's buffers have the logical width
    for (int i = 0; i < in; i++) memcpy(W1 + (size_t)i * H, p + (size_t)i * kMlpH, sizeof(float) * H);
    p += (size_t)in * kMlpH;
    memcpy(b1, p, sizeof(float) * H); p += kMlpH;
    for (int i = 0; i < H; i++) memcpy(W2 + (size_t)i * H, p + (size_t)i * kMlpH, sizeof(float) * H);
    p += (size_t)kMlpH * kMlpH;
    memcpy(b2, p, sizeof(float) * H); p += kMlpH;
    memcpy(W3, p, sizeof(float) * H * h->s); p += (size_t)kMlpH * h->s;
    memcpy(b3, p, sizeof(float) * h->s);
    return MPPI_OK;
}

int mppi_mlp_predict(mppi_handle *h, int kst, int k, const float *state, const float *action, float *out)
{
    if (!h || !state || !action || !out || k <= 0 || (kst != 1 && kst != k)) return fail(h, MPPI_ERR_BAD_ARG, "bad mlp_predict argument");
    if (!h->mlp) return fail(h, MPPI_ERR_STATE, "mppi_set_mlp has not been called");
    CU_TRY(h, cudaSetDevice(h->device));
    DevBuf ds, da, dout;
    CU_TRY(h, ds.alloc(sizeof(float) * (size_t)kst * h->s));
    CU_TRY(h, da.alloc(sizeof(float) * (size_t)k * h->a));
    CU_TRY(h, dout.alloc(sizeof(float) * (size_t)k * h->s));
    CU_TRY(h, cudaMemcpyAsync(ds.p, state, sizeof(float) * (size_t)kst * h->s, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(da.p, action, sizeof(float) * (size_t)k * h->a, cudaMemcpyHostToDevice, h->stream));
    MlpParams mp{h->d_wblob, h->d_fvec, h->s, h->a};
    CU_TRY(h, launch_mlp_predict(mp, kst, k, ds.as<float>(), da.as<float>(), dout.as<float>(), h->stream));
    return d2h(h, out, dout.as<float>(), (size_t)k * h->s);
}

// ---- stateless stages -------------------------------------------------------------------------------
int mppi_block_diag(const float *in, int rows, int cols, int nb, float *out)
{
    if (!in || !out || rows <= 0 || cols <= 0 || nb <= 0) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad block_diag argument");
    const int R = rows * nb, C = cols * nb;
    for (int i = 0; i < R * C; i++) out[i] = 0.f;
    for (int b = 0; b < nb; b++)
        for (int r = 0; r < rows; r++)
            memcpy(out + (size_t)(b * rows + r) * C + b * cols, in + (size_t)r * cols, sizeof(float) * cols);
    return MPPI_OK;
}

static int stage_device(int device)
{
    cudaDeviceProp prop;
    int dev;
    return select_device(nullptr, device, &dev, &prop);
}
#define CU_TRY_S(expr) CU_TRY((mppi_handle *)nullptr, expr)

static int model_stage(int device, float mass, float dt, int s, int a, int kst, int k, const float *state,
                       const float *action, float *out, int mode)
{
    if (!out || s != 2 * a || a <= 0 || (mode != 1 && (!state || kst <= 0)) || (mode != 0 && (!action || k <= 0)))
        return fail(nullptr, MPPI_ERR_BAD_ARG, "bad model stage argument");
    if (mode == 2 && kst != 1 && kst != k) return fail(nullptr, MPPI_ERR_BAD_ARG, "state batch must be 1 or k");
    int rc = stage_device(device);
    if (rc) return rc;
    const int n_out = (mode == 0 ? kst : k);
    DevBuf ds, da, dout;
    CU_TRY_S(ds.alloc(sizeof(float) * (size_t)(mode != 1 ? kst : 1) * s));
    CU_TRY_S(da.alloc(sizeof(float) * (size_t)(mode != 0 ? k : 1) * a));
    CU_TRY_S(dout.alloc(sizeof(float) * (size_t)n_out * s));
    if (mode != 1) CU_TRY_S(cudaMemcpy(ds.p, state, sizeof(float) * (size_t)kst * s, cudaMemcpyHostToDevice));
    if (mode != 0) CU_TRY_S(cudaMemcpy(da.p, action, sizeof(float) * (size_t)k * a, cudaMemcpyHostToDevice));
    CU_TRY_S(launch_model_step(mass, dt, s, a, kst, k, ds.as<float>(), da.as<float>(), dout.as<float>(), mode, 0));
    CU_TRY_S(cudaMemcpy(out, dout.p, sizeof(float) * (size_t)n_out * s, cudaMemcpyDeviceToHost));
    return MPPI_OK;
}
int mppi_model_free_step(int device, float mass, float dt, int s, int a, int kst, const float *state, float *out)
{
    return model_stage(device, mass, dt, s, a, kst, kst, state, nullptr, out, 0);
}
int mppi_model_action_step(int device, float mass, float dt, int s, int a, int k, const float *action, float *out)
{
    return model_stage(device, mass, dt, s, a, 1, k, nullptr, action, out, 1);
}
int mppi_model_step(int device, float mass, float dt, int s, int a, int kst, int k, const float *state,
                    const float *action, float *out)
{
    return model_stage(device, mass, dt, s, a, kst, k, state, action, out, 2);
}

static int cost_stage(int device, int k, int s, int a, float lambda, const float *sigma, const float *goal,
                      const float *q, const float *state, const float *action, const float *noise, float *out, int mode)
{
    if (!out || k <= 0) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad cost stage argument");
    if (mode != 1 && (!state || !goal || !q || s <= 0 || s > 64)) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad state-cost argument");
    if (mode != 0 && (!sigma || !action || !noise || a <= 0 || a > MPPI_MAX_A)) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad action-cost argument");
    int rc = stage_device(device);
    if (rc) return rc;
    float inv[kMaxA * kMaxA] = {0};
    if (mode != 0 && !invert_matrix(sigma, a, inv)) return fail(nullptr, MPPI_ERR_BAD_ARG, "sigma is singular");
    DevBuf dinv, dg, dq, dst, dac, dn, dout;
    const int ss = s > 0 ? s : 1, aa = a > 0 ? a : 1;
    CU_TRY_S(dinv.alloc(sizeof(inv)));
    CU_TRY_S(dg.alloc(sizeof(float) * ss)); CU_TRY_S(dq.alloc(sizeof(float) * ss));
    CU_TRY_S(dst.alloc(sizeof(float) * (size_t)k * ss));
    CU_TRY_S(dac.alloc(sizeof(float) * aa)); CU_TRY_S(dn.alloc(sizeof(float) * (size_t)k * aa));
    CU_TRY_S(dout.alloc(sizeof(float) * (size_t)k));
    CU_TRY_S(cudaMemcpy(dinv.p, inv, sizeof(inv), cudaMemcpyHostToDevice));
    if (mode != 1) {
        CU_TRY_S(cudaMemcpy(dg.p, goal, sizeof(float) * s, cudaMemcpyHostToDevice));
        CU_TRY_S(cudaMemcpy(dq.p, q, sizeof(float) * s, cudaMemcpyHostToDevice));
        CU_TRY_S(cudaMemcpy(dst.p, state, sizeof(float) * (size_t)k * s, cudaMemcpyHostToDevice));
    }
    if (mode != 0) {
        CU_TRY_S(cudaMemcpy(dac.p, action, sizeof(float) * a, cudaMemcpyHostToDevice));
        CU_TRY_S(cudaMemcpy(dn.p, noise, sizeof(float) * (size_t)k * a, cudaMemcpyHostToDevice));
    }
    CU_TRY_S(launch_cost(k, s, a, lambda, dinv.as<float>(), dg.as<float>(), dq.as<float>(), dst.as<float>(),
                         dac.as<float>(), dn.as<float>(), dout.as<float>(), mode, 0));
    CU_TRY_S(cudaMemcpy(out, dout.p, sizeof(float) * (size_t)k, cudaMemcpyDeviceToHost));
    return MPPI_OK;
}
int mppi_cost_state(int device, int k, int s, const float *state, const float *goal, const float *q, float *out)
{
    return cost_stage(device, k, s, 0, 1.f, nullptr, goal, q, state, nullptr, nullptr, out, 0);
}
int mppi_cost_action_py(int device, int k, int a, float lambda, float gamma, float upsilon, const float *sigma,
                        const float *action, const float *noise, float *out)
{
    if (!sigma || !action || !noise || !out || k <= 0 || a <= 0 || a > MPPI_MAX_A || !(upsilon > 0.f))
        return fail(nullptr, MPPI_ERR_BAD_ARG, "bad action-cost argument");
    int rc = stage_device(device);
    if (rc) return rc;
    float inv[kMaxA * kMaxA] = {0};
    if (!invert_matrix(sigma, a, inv)) return fail(nullptr, MPPI_ERR_BAD_ARG, "sigma is singular");
    DevBuf dinv, dac, dn, dout;
    CU_TRY_S(dinv.alloc(sizeof(inv)));
    CU_TRY_S(dac.alloc(sizeof(float) * a));
    CU_TRY_S(dn.alloc(sizeof(float) * (size_t)k * a));
    CU_TRY_S(dout.alloc(sizeof(float) * (size_t)k));
    CU_TRY_S(cudaMemcpy(dinv.p, inv, sizeof(inv), cudaMemcpyHostToDevice));
    CU_TRY_S(cudaMemcpy(dac.p, action, sizeof(float) * a, cudaMemcpyHostToDevice));
    CU_TRY_S(cudaMemcpy(dn.p, noise, sizeof(float) * (size_t)k * a, cudaMemcpyHostToDevice));
    CU_TRY_S(launch_action_cost_py(k, a, lambda, gamma, upsilon, dinv.as<float>(), dac.as<float>(), dn.as<float>(),
                                   dout.as<float>(), 0));
    CU_TRY_S(cudaMemcpy(out, dout.p, sizeof(float) * (size_t)k, cudaMemcpyDeviceToHost));
    return MPPI_OK;
}
int mppi_cost_state_ellipse(int device, int k, const float *state, float a, float b, float center_x, float center_y,
                            float speed, float m_state, float m_vel, float *out)
{
    if (!state || !out || k <= 0) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad ellipse-cost argument");
    float ell[8];
    if (!ellipse_params(a, b, center_x, center_y, speed, m_state, m_vel, ell)) return fail(nullptr, MPPI_ERR_BAD_ARG, "ellipse axes must be non-zero");
    int rc = stage_device(device);
    if (rc) return rc;
    DevBuf dst, dout;
    CU_TRY_S(dst.alloc(sizeof(float) * (size_t)k * 4));
    CU_TRY_S(dout.alloc(sizeof(float) * (size_t)k));
    CU_TRY_S(cudaMemcpy(dst.p, state, sizeof(float) * (size_t)k * 4, cudaMemcpyHostToDevice));
    CU_TRY_S(launch_ellipse_cost(k, dst.as<float>(), ell, dout.as<float>(), 0));
    CU_TRY_S(cudaMemcpy(out, dout.p, sizeof(float) * (size_t)k, cudaMemcpyDeviceToHost));
    return MPPI_OK;
}
int mppi_cost_action(int device, int k, int a, float lambda, const float *sigma, const float *action,
                     const float *noise, float *out)
{
    return cost_stage(device, k, 0, a, lambda, sigma, nullptr, nullptr, nullptr, action, noise, out, 1);
}
int mppi_cost_step(int device, int k, int s, int a, float lambda, const float *sigma, const float *goal,
                   const float *q, const float *state, const float *action, const float *noise, float *out)
{
    return cost_stage(device, k, s, a, lambda, sigma, goal, q, state, action, noise, out, 2);
}

int mppi_prepare_action(int T, int a, const float *U, int t, float *out)
{
    if (!U || !out || t < 0 || t >= T || a <= 0) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad prepare_action argument");
    memcpy(out, U + (size_t)t * a, sizeof(float) * a);
    return MPPI_OK;
}
int mppi_prepare_noise(int device, int k, int T, int a, const float *noise, int t, float *out)
{
    if (!noise || !out || k <= 0 || a <= 0 || t < 0 || t >= T) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad prepare_noise argument");
    int rc = stage_device(device);
    if (rc) return rc;
    DevBuf dn, dout;
    CU_TRY_S(dn.alloc(sizeof(float) * (size_t)k * T * a));
    CU_TRY_S(dout.alloc(sizeof(float) * (size_t)k * a));
    CU_TRY_S(cudaMemcpy(dn.p, noise, sizeof(float) * (size_t)k * T * a, cudaMemcpyHostToDevice));
    CU_TRY_S(launch_prepare_noise(k, T, a, dn.as<float>(), t, dout.as<float>(), 0));
    CU_TRY_S(cudaMemcpy(out, dout.p, sizeof(float) * (size_t)k * a, cudaMemcpyDeviceToHost));
    return MPPI_OK;
}

int mppi_update_stages(int device, int k, int T, int a, float lambda, const float *cost, const float *noise,
                       float *beta, float *exp_arg, float *exp_out, float *nabla, float *weights,
                       float *weighted_noise)
{
    if (!cost || !noise || k <= 0 || T <= 0 || a <= 0 || !(lambda > 0.f)) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad update_stages argument");
    int rc = stage_device(device);
    if (rc) return rc;
    DevBuf dc, dn, dscal, darg, dexp, dw, dwn;
    CU_TRY_S(dc.alloc(sizeof(float) * (size_t)k));
    CU_TRY_S(dn.alloc(sizeof(float) * (size_t)k * T * a));
    CU_TRY_S(dscal.alloc(sizeof(float) * 2));
    CU_TRY_S(darg.alloc(sizeof(float) * (size_t)k));
    CU_TRY_S(dexp.alloc(sizeof(float) * (size_t)k));
    CU_TRY_S(dw.alloc(sizeof(float) * (size_t)k));
    CU_TRY_S(dwn.alloc(sizeof(float) * (size_t)T * a));
    CU_TRY_S(cudaMemcpy(dc.p, cost, sizeof(float) * (size_t)k, cudaMemcpyHostToDevice));
    CU_TRY_S(cudaMemcpy(dn.p, noise, sizeof(float) * (size_t)k * T * a, cudaMemcpyHostToDevice));
    CU_TRY_S(launch_update_stages(k, T, a, lambda, dc.as<float>(), dn.as<float>(), dscal.as<float>(), darg.as<float>(),
                                  dexp.as<float>(), dw.as<float>(), dwn.as<float>(), 0));
    float scal[2];
    CU_TRY_S(cudaMemcpy(scal, dscal.p, sizeof(scal), cudaMemcpyDeviceToHost));
    if (beta) *beta = scal[0];
    if (nabla) *nabla = scal[1];
    if (exp_arg) CU_TRY_S(cudaMemcpy(exp_arg, darg.p, sizeof(float) * (size_t)k, cudaMemcpyDeviceToHost));
    if (exp_out) CU_TRY_S(cudaMemcpy(exp_out, dexp.p, sizeof(float) * (size_t)k, cudaMemcpyDeviceToHost));
    if (weights) CU_TRY_S(cudaMemcpy(weights, dw.p, sizeof(float) * (size_t)k, cudaMemcpyDeviceToHost));
    if (weighted_noise) CU_TRY_S(cudaMemcpy(weighted_noise, dwn.p, sizeof(float) * (size_t)T * a, cudaMemcpyDeviceToHost));
    return MPPI_OK;
}

int mppi_stage_vector_op(int device, int op, int k, const float *in, float s0, float s1, float *out)
{
    if (!in || !out || k <= 0 || op < MPPI_OP_MIN || op > MPPI_OP_DIV) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad stage_vector_op argument");
    if (op == MPPI_OP_EXP_ARG && !(s1 != 0.f)) return fail(nullptr, MPPI_ERR_BAD_ARG, "lambda must be non-zero");
    int rc = stage_device(device);
    if (rc) return rc;
    const size_t n_out = (op == MPPI_OP_MIN || op == MPPI_OP_SUM) ? 1 : (size_t)k;
    DevBuf din, dout;
    CU_TRY_S(din.alloc(sizeof(float) * (size_t)k));
    CU_TRY_S(dout.alloc(sizeof(float) * n_out));
    CU_TRY_S(cudaMemcpy(din.p, in, sizeof(float) * (size_t)k, cudaMemcpyHostToDevice));
    CU_TRY_S(launch_vector_op(op, k, din.as<float>(), s0, s1, dout.as<float>(), 0));
    CU_TRY_S(cudaMemcpy(out, dout.p, sizeof(float) * n_out, cudaMemcpyDeviceToHost));
    return MPPI_OK;
}

int mppi_weighted_noise(int device, int k, int TA, const float *weights, const float *noise, float *out)
{
    if (!weights || !noise || !out || k <= 0 || TA <= 0) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad weighted_noise argument");
    int rc = stage_device(device);
    if (rc) return rc;
    DevBuf dw, dn, dout;
    CU_TRY_S(dw.alloc(sizeof(float) * (size_t)k));
    CU_TRY_S(dn.alloc(sizeof(float) * (size_t)k * TA));
    CU_TRY_S(dout.alloc(sizeof(float) * (size_t)TA));
    CU_TRY_S(cudaMemcpy(dw.p, weights, sizeof(float) * (size_t)k, cudaMemcpyHostToDevice));
    CU_TRY_S(cudaMemcpy(dn.p, noise, sizeof(float) * (size_t)k * TA, cudaMemcpyHostToDevice));
    CU_TRY_S(launch_weighted_noise(k, TA, dw.as<float>(), dn.as<float>(), dout.as<float>(), 0));
    CU_TRY_S(cudaMemcpy(out, dout.p, sizeof(float) * (size_t)TA, cudaMemcpyDeviceToHost));
    return MPPI_OK;
}

int mppi_get_new(int T, int a, const float *cur, int nb, float *out)
{
    if (!cur || nb < 0 || nb > T || a <= 0 || (nb > 0 && !out)) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad get_new argument");
    if (nb > 0) memcpy(out, cur, sizeof(float) * (size_t)nb * a);
    return MPPI_OK;
}
int mppi_shift(int T, int a, const float *cur, const float *init, int nb, float *out)
{
    if (!cur || !out || nb < 0 || nb > T || a <= 0 || (nb > 0 && !init)) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad shift argument");
    memmove(out, cur + (size_t)nb * a, sizeof(float) * (size_t)(T - nb) * a);
    if (nb > 0) memcpy(out + (size_t)(T - nb) * a, init, sizeof(float) * (size_t)nb * a);
    return MPPI_OK;
}

int mppi_philox_raw(int device, uint64_t seed, uint32_t call0, uint32_t sample, uint32_t update, uint32_t stream,
                    int n_calls, uint32_t *out)
{
    return mppi_philox_raw_rounds(device, seed, call0, sample, update, stream, n_calls, 10, out);
}

int mppi_philox_raw_rounds(int device, uint64_t seed, uint32_t call0, uint32_t sample, uint32_t update, uint32_t stream,
                           int n_calls, int rounds, uint32_t *out)
{
    if (!out || n_calls <= 0 || (rounds != 7 && rounds != 10)) return fail(nullptr, MPPI_ERR_BAD_ARG, "bad philox_raw argument");
    int rc = stage_device(device);
    if (rc) return rc;
    DevBuf d;
    CU_TRY_S(d.alloc(sizeof(uint32_t) * 4 * (size_t)n_calls));
    CU_TRY_S(launch_philox_raw(seed, call0, sample, update, stream, n_calls, rounds, d.as<uint32_t>(), 0));
    CU_TRY_S(cudaMemcpy(out, d.p, sizeof(uint32_t) * 4 * (size_t)n_calls, cudaMemcpyDeviceToHost));
    return MPPI_OK;
}

}  // extern "C"
