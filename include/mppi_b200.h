/*
 * mppi_b200.h — C-ABI of the B200-native MPPI update step.
 *
 * This is the drop-in boundary for the hot path of NicolayP/mppi-tf: everything the reference
 * executes inside `ClientSession::Run` for one `ControllerBase::next` call
 * (/root/reference/src/controller_base.cpp:135-153) plus the sub-graph builders its unit tests
 * exercise.  The reference has no FFI of its own (its boundary is the public C++ API of
 * ControllerBase / ModelBase / CostBase, linked into one executable, CMakeLists.txt:56-64), so
 * each entry point below names the reference C++ method it replaces.  The C++ classes in
 * include/controller_base.hpp, model_base.hpp, cost_base.hpp, data_base.hpp keep the reference's
 * class names and numerical method signatures and are thin callers of this ABI.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types; never throws.
 *   - every function returns an mppi_status (0 = ok); mppi_last_error() gives the text.
 *   - all tensors are fp32, row-major, the reference's trailing singleton dimension dropped:
 *       state [k][s]   action [k][a]   noise [k][T][a]   sequence U [T][a]
 *     state layout per axis is interleaved (pos0, vel0, pos1, vel1, ...), s = 2a
 *     (/root/reference/src/model_base.cpp:61-64).
 *   - "host" pointers are ordinary host memory; the library copies in/out.  "dev" pointers are
 *     CUDA device memory on the handle's device.
 *   - a handle is not thread-safe; distinct handles are independent (one CUDA stream each).
 *   - there is NO CPU fallback: every compute entry point fails with MPPI_ERR_CUDA if no
 *     sm_100 device is usable.
 */
#ifndef MPPI_B200_H
#define MPPI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    MPPI_OK = 0,
    MPPI_ERR_BAD_ARG = 1,     /* null pointer, non-positive size, s != 2a, ... */
    MPPI_ERR_CUDA = 2,        /* CUDA runtime/driver failure (text in mppi_last_error) */
    MPPI_ERR_COMM = 3,        /* multi-rank exchange failure */
    MPPI_ERR_UNSUPPORTED = 4, /* size outside the compiled limits (see MPPI_MAX_*) */
    MPPI_ERR_STATE = 5        /* call sequence error (e.g. dump_noise before any next) */
} mppi_status;

#define MPPI_MAX_A 8          /* a_dim <= 8, s_dim = 2 a_dim <= 16 */
#define MPPI_MAX_S 16
#define MPPI_MAX_TA 4096      /* tau * a_dim accepted by mppi_create; the update kernels keep per-warp rows of tau * a_dim
                               * partial sums in shared memory, so an update with tau * a_dim above about 2400 (Philox mode) or
                               * 1700 (injected noise) returns MPPI_ERR_UNSUPPORTED on a 227 KB part */
#define MPPI_MAX_PEERS 8      /* ranks of one NVLink domain the fused exchange can address */

typedef struct mppi_handle mppi_handle;

/* Dynamics model selector (ModelBase vs the learning_base MLP). */
typedef enum { MPPI_MODEL_POINT_MASS = 0, MPPI_MODEL_MLP = 1, MPPI_MODEL_AUV = 2 } mppi_model_kind;

/*
 * Construction parameters.  Mirrors ControllerBase(k, tau, dt, mass, s_dim, a_dim)
 * (/root/reference/include/controller_base.hpp:60-65) with the constants the reference
 * hard-codes in that constructor made explicit (src/controller_base.cpp:37-69):
 *   lambda = 1, sigma = I, goal = (1,0,1,0,...), q = ones, model mass = 1.
 * NULL sigma/goal/q select exactly those defaults.
 */
typedef struct {
    int k;                 /* samples per controller, summed over all ranks */
    int tau;               /* horizon T */
    int s_dim;             /* must equal 2 * a_dim for the point-mass model; 13 for MPPI_MODEL_AUV */
    int a_dim;
    float dt;
    float mass;            /* model mass used in B = [dt^2/2; dt] / mass */
    float lambda;
    const float *sigma;    /* [a][a]; scale in eps = sigma z AND inverted in the action cost */
    const float *goal;     /* [s] (shared) or [n_controllers][s] if goal_per_controller != 0 */
    const float *q;        /* [s] diagonal of Q */
    uint64_t seed;         /* Philox key; the reference uses RandomNormal::Seed(1) (:199) */
    int device;            /* CUDA device ordinal, -1 = current device */
    int rank;              /* this process's shard: samples [rank*k/world, (rank+1)*k/world) */
    int world;             /* number of sample shards (1 = single GPU) */
    int n_controllers;     /* independent controllers batched in one handle (>= 1) */
    int goal_per_controller;
    void *stream;          /* cudaStream_t to launch on; NULL = library-owned stream */
    int model;             /* mppi_model_kind at construction: MPPI_MODEL_POINT_MASS (0, default) or MPPI_MODEL_AUV
                            * (s_dim = 13, a_dim = 6; mppi_set_auv_model must follow).  MPPI_MODEL_MLP is selected
                            * later, by mppi_set_mlp on a point-mass handle. */
    int philox_rounds;     /* rounds of the Philox4x32 noise generator: 10 (Random123's default and the generator behind
                            * TensorFlow's RandomNormal; 0 means 10) or 7 (the shortest variant of Salmon et al. that passes
                            * BigCrush; SURVEY.md section 7 allows a cheaper generator: parity is checked on injected noise
                            * and on store-then-replay of whatever stream was drawn).  Both are pinned on Random123's
                            * known-answer vectors (tests/golden/kats.py). */
} mppi_config;

/* Fill *cfg with the reference constructor's defaults for the given sizes. */
void mppi_config_default(mppi_config *cfg, int k, int tau, float dt, float mass, int s_dim, int a_dim);

/* ---- lifetime --------------------------------------------------------------------------- */
/* replaces ControllerBase::ControllerBase + mBuildGraph (src/controller_base.cpp:23-71,275-308) */
int mppi_create(const mppi_config *cfg, mppi_handle **out);
int mppi_destroy(mppi_handle *h);
/* Text of the last error on this handle (h may be NULL for the last create/stateless error). */
const char *mppi_last_error(const mppi_handle *h);

/* ---- the update step ---------------------------------------------------------------------- */
/*
 * replaces ControllerBase::next (src/controller_base.cpp:135-153).
 * x_host [n_controllers][s] -> action_host [n_controllers][a].  Draws fresh noise
 * (counter-based Philox, regenerated in registers), rolls out, reweights, updates and shifts the
 * stored sequence.  Synchronous: returns after the action is on the host.
 * With world > 1 the exchange must have been wired with mppi_exchange_* first.
 */
int mppi_next(mppi_handle *h, const float *x_host, float *action_host);

/*
 * Same update with the noise tensor INJECTED instead of generated — the parity/debug mode
 * required by BASELINE.json ("a debug mode reads eps from a buffer").
 * eps [n_controllers][k_local][T][a] are the already-scaled noises eps = sigma z
 * (the output of mNoiseGenGraph, src/controller_base.cpp:194-203) for this rank's samples.
 */
int mppi_next_with_noise(mppi_handle *h, const float *x_host, const float *eps_host, float *action_host);
int mppi_next_with_noise_dev(mppi_handle *h, const float *x_host, const float *eps_dev, float *action_host);

/*
 * Asynchronous halves used by the benchmark and by multi-rank callers.
 * mppi_enqueue_update launches the fused rollout (+ merge when world == 1) on the handle's
 * stream and returns without synchronising; eps_dev == NULL selects Philox mode.  The state must
 * have been staged with mppi_set_state.  With world > 1 the launch leaves this rank's payload
 * (beta_r, eta_r, N_r) in the exchange send buffer; after the caller's all-gather,
 * mppi_enqueue_finish merges the gathered payloads and applies the update.
 * mppi_fetch_action copies the [n_controllers][a] action to the host and synchronises.
 */
int mppi_set_state(mppi_handle *h, const float *x_host);
int mppi_enqueue_update(mppi_handle *h, const float *eps_dev);
int mppi_enqueue_exchange(mppi_handle *h);   /* in-library ncclAllGather of the payloads (after mppi_comm_init) */
int mppi_enqueue_finish(mppi_handle *h);
int mppi_fetch_action(mppi_handle *h, float *action_host);
int mppi_synchronize(mppi_handle *h);

/* ---- controller state ---------------------------------------------------------------------- */
/* replaces ControllerBase::setGoal (src/controller_base.cpp:126-133) — and takes effect, which
 * the reference's does not (the goal is baked into the graph as a Const, src/cost_base.cpp:57). */
int mppi_set_goal(mppi_handle *h, const float *goal_host);       /* [s] (shared goal) or [n][s] (goal_per_controller): the
                                                                  * handle's mode decides how many floats are read */
/* The same with the row count stated: n_rows = 1 sets the goal of every controller (broadcast on goal_per_controller
 * handles); n_rows = n_controllers needs a goal_per_controller handle (MPPI_ERR_BAD_ARG otherwise).  The C++ and Python
 * ControllerBase::setGoal go through this entry, so a size that does not fit the handle's mode is never read past. */
int mppi_set_goal_n(mppi_handle *h, const float *goal_host, int n_rows);
int mppi_set_lambda(mppi_handle *h, float lambda);
int mppi_set_sigma(mppi_handle *h, const float *sigma_host);     /* [a][a], must be invertible */
/* Python-twin extras (/root/reference/scripts/src):
 *   upsilon scales the sampling, eps = (upsilon sigma) z  (controllers/controller_base.py:348-369);
 *   MPPI_ACTION_COST_CPP     lambda u^T sigma^-1 eps                        (src/cost_base.cpp:63-68, the default)
 *   MPPI_ACTION_COST_PYTHON  0.5 [gamma (u^T S^-1 u + 2 u^T S^-1 eps) + lambda (1 - 1/upsilon) eps^T S^-1 eps]
 *                            (costs/cost_base.py:114-170); u is the un-noised U[t] in both forms.
 * With injected noise the caller passes the already scaled eps, as the reference's build_noise returns it. */
typedef enum { MPPI_ACTION_COST_CPP = 0, MPPI_ACTION_COST_PYTHON = 1 } mppi_action_cost_form;
int mppi_set_action_cost(mppi_handle *h, int form, float gamma, float upsilon);
/* State-cost functor.  ElipseCost.state_cost
 * (/root/reference/scripts/src/costs/elipse_cost.py:9-79) on the point_mass2d state (x, vx, y, vy):
 *   m_state |((x-cx)/a)^2 + ((y-cy)/b)^2 - 1| + m_vel (sqrt(vx^2 + vy^2) - speed)^2
 * replaces (x-g)^T Q (x-g) in the rollout and as terminal cost; needs s_dim = 4, a_dim = 2.
 * mppi_set_static_cost switches back to the quadratic StaticCost (goal / q are kept). */
int mppi_set_ellipse_cost(mppi_handle *h, float a, float b, float center_x, float center_y, float speed,
                          float m_state, float m_vel);
int mppi_set_static_cost(mppi_handle *h);
/* norm_arg (controllers/controller_base.py:468-474): the exponent becomes -(S - beta) / (lambda max_k(S_k - beta)).
 * Costs two launches per update (the range must be known before any weight).  With world > 1 the range is
 * exchanged through the fused peer-memory mailboxes: call mppi_peer_attach first (MPPI_ERR_UNSUPPORTED otherwise). */
int mppi_set_normalize_cost(mppi_handle *h, int on);
int mppi_set_q(mppi_handle *h, const float *q_host);             /* [s] */
/* clip_act of the Python twin (controllers/controller_base.py:500-504, models/model_base.py:121-126): the updated sequence is
 * clipped to [act_min, act_max] per action axis before the next action is taken and the sequence shifted,
 * U' = clip(U + sum_k w_k eps_k).  n = 1 (one limit for every axis, the reference's default shape) or n = a_dim; on = 0 turns
 * the clipping off again (the state at the reference's HEAD, where the call is commented out, :455-456). */
int mppi_set_action_limits(mppi_handle *h, int on, int n, const float *act_min, const float *act_max);
int mppi_set_mass(mppi_handle *h, float mass);                   /* model mass in B = [dt^2/2; dt] / mass */
int mppi_set_sequence(mppi_handle *h, const float *U_host);      /* m_U, [n][T][a] */
int mppi_get_sequence(mppi_handle *h, float *U_host);            /* shifted sequence kept for the next call */
int mppi_get_update(mppi_handle *h, float *U_new_host);          /* U + sum_k w_k eps_k of the last call, pre-shift */
int mppi_get_costs(mppi_handle *h, float *costs_host);           /* [n][k_local] per-sample costs of the last call */
int mppi_get_weight_stats(mppi_handle *h, float *beta, float *eta); /* [n] each: min cost, sum of exp */
int mppi_set_update_counter(mppi_handle *h, uint32_t counter);   /* Philox "update" counter word */
/* Regenerates the eps tensor of the LAST Philox-mode update for this rank's samples
 * ([n][k_local][T][a], host) so it can be replayed through the injected path / the oracle. */
int mppi_dump_noise(mppi_handle *h, float *eps_host);
int mppi_k_local(const mppi_handle *h);
int mppi_k_offset(const mppi_handle *h);

/* ---- multi-rank exchange (sample sharding) ------------------------------ */
/* Payload per rank: n_controllers * mppi_exchange_stride(h) floats =
 * {beta_r, eta_r, 0, 0, N_r[T*a] (padded to a multiple of 4)} per controller. */
int mppi_exchange_stride(const mppi_handle *h);
/* Pure host helpers (no GPU needed): the shard of rank `rank` of `world` over k samples, and the
 * payload stride for a given tau * a_dim — the contract multi-rank callers and tests rely on. */
int mppi_shard_range(int k, int rank, int world, int *k_offset, int *k_local);
int mppi_payload_stride(int tau_times_a);
/* Device buffers the caller all-gathers between enqueue_update and enqueue_finish:
 * send = this rank's payload, recv = [world] payloads in rank order.  The library owns them
 * unless external ones are supplied (e.g. tensors registered with torch.distributed). */
int mppi_exchange_buffers(mppi_handle *h, void **send_dev, void **recv_dev);
int mppi_exchange_set_buffers(mppi_handle *h, void *send_dev, void *recv_dev);
/* In-library exchange over NCCL (dlopen'ed libnccl.so.2): every rank passes the same 128-byte
 * ncclUniqueId (mppi_comm_unique_id fills one on the caller that broadcasts it). */
int mppi_comm_unique_id(void *id128);
int mppi_comm_init(mppi_handle *h, const void *id128);

/* Fused exchange over peer memory (one box, NVLink / NVSwitch): instead of all-gather + finish, the last CTA
 * of the update kernel stores its payload into every rank's mailbox (every word together with the update's epoch in
 * one 8-byte store: no fence, no flag), polls its own mailbox for the other ranks' payloads and finishes the update
 * in the same launch.  Every rank exports the 64-byte CUDA IPC handle of
 * its mailbox (mppi_peer_handle), the caller all-gathers them (any host channel) and every rank attaches the
 * world handles in rank order (mppi_peer_attach; world <= MPPI_MAX_PEERS, one process per GPU).  After a
 * successful attach mppi_next needs no exchange call, and mppi_enqueue_exchange / _finish are no-ops.
 * A payload that does not arrive within about a second makes mppi_fetch_action return MPPI_ERR_COMM. */
#define MPPI_PEER_HANDLE_BYTES 64
int mppi_peer_handle(mppi_handle *h, void *handle64);
int mppi_peer_attach(mppi_handle *h, const void *handles /* [world][64] */);

/* ---- learned MLP dynamics (learning_base, row A13) ------------------------------------------- */
/* x' = x + (W3^T relu(W2^T relu(W1^T X + b1) + b2) + b3) * Ystd + Ymean,
 * X = (concat(x, u) - Xmean) / Xstd; weights in Keras layout [in][out]
 * (behaviour of /root/reference/scripts/src/models/nn_model.py:54-60,215-239,289-304).
 * Switches the handle to MPPI_MODEL_MLP; bf16 tensor-core rollout, fp32 state.  hidden = 1..128 (the tensor-core tiles are
 * 128 wide: narrower networks - the reference's use 32 units - are zero padded, which is exact: a padded unit outputs
 * relu(0) = 0 and its gradients vanish), two hidden layers, point-mass state layout, s + a <= 15 (a <= 5: the input tile has
 * 16 columns, one of which carries the bias).  This is the BASELINE config-4 network (2 x 128 on the point_mass3d state), a
 * new specification derived from the reference's NNModel family: NNModel.build_step_graph itself raises NotImplementedError
 * (nn_model.py:101-117).  The reference's one concrete learned model, NNAUVModel (three hidden layers of 32 on the AUV state
 * without its position, nn_model.py:181-304), is a separate model kind: mppi_set_nn_auv_model below. */
int mppi_set_mlp(mppi_handle *h, int hidden, const float *W1, const float *b1, const float *W2,
                 const float *b2, const float *W3, const float *b3, const float *Xmean,
                 const float *Xstd, const float *Ymean, const float *Ystd);

/* One batched MLP step on the tensor cores (the learned model's `predict`): state [kst][s] with kst in
 * {1, k}, action [k][a] -> next [k][s].  Requires mppi_set_mlp. */
int mppi_mlp_predict(mppi_handle *h, int kst, int k, const float *state, const float *action, float *out);

/* Learner side of the MLP model: ONE full-batch Adam step on the mean squared
 * error of the normalised prediction, as LearnerBase._train_step does
 * (/root/reference/scripts/src/learners/learner_base.py:469-496 with tf.optimizers.Adam, :325):
 *   Xn = (concat(x,u) - Xmean)/Xstd, Yn = ((x' - x) - Ymean)/Ystd, loss = mean((net(Xn) - Yn)^2)
 * over n transitions (state [n][s], action [n][a], next_state [n][s], host).  fp32 master weights and Adam
 * moments live in the handle; the bf16 copy the rollout uses is refreshed after the step, so the next
 * mppi_next already plans with the updated model.  *loss_out (may be NULL) is the loss BEFORE the step.
 * mppi_mlp_set_adam changes (beta1, beta2, epsilon) from the Keras defaults (0.9, 0.999, 1e-7) and resets the
 * moments and the step count; mppi_set_mlp resets them too. */
int mppi_mlp_train_step(mppi_handle *h, int n, const float *state, const float *action, const float *next_state,
                        float learning_rate, float *loss_out);
int mppi_mlp_set_adam(mppi_handle *h, float beta1, float beta2, float epsilon);
/* The learner's loop, LearnerBase.train (learners/learner_base.py:324-358): `epochs` times { augment_data (:455-467) when
 * augment_samples > 0: every transition repeated augment_samples times with Gaussian noise of standard deviation
 * augment_sigma on the normalised inputs (Philox, key = seed), the normalised target kept; then Adam on the normalised MSE }.
 * batch_size <= 0: one full-batch step per epoch, as the reference runs it (its batchSize argument is never used,
 * :146-153,470); batch_size > 0: consecutive minibatches of the epoch's data (an extension).  The data are uploaded once and
 * the steps run back to back on the handle's stream.  losses_out (may be NULL, room for epochs * ceil(n_epoch / batch)
 * floats) receives the loss before every step; *n_steps_out (may be NULL) their number.  Deviation: the reference hands the
 * augmented data to the step DE-normalised (:463-467) although the step expects normalised data; both agree for the
 * default normalisation (mean 0, std 1, nn_model.py:65-69). */
int mppi_mlp_train(mppi_handle *h, int n, const float *state, const float *action, const float *next_state, int epochs,
                   int batch_size, float learning_rate, int augment_samples, float augment_sigma, uint64_t seed, float *losses_out,
                   int *n_steps_out);
int mppi_mlp_get_weights(mppi_handle *h, float *W1, float *b1, float *W2, float *b2, float *W3, float *b3);

/* ---- AUV (Fossen) dynamics and the quaternion goal cost -------------- */
/* Parameters of AUVModel (/root/reference/scripts/src/models/auv_model.py:146-255); matrices row-major [6][6].
 * State x = (position[3], quaternion (qx, qy, qz, qw), body velocity nu[6]); action = generalised force [6]. */
typedef struct {
    float mass, volume, density;
    float cog[3], cob[3];
    float added_mass[36];                     /* "Ma" */
    float inertia[6];                         /* ixx, iyy, izz, ixy, ixz, iyz */
    float linear_damping[36];                 /* a diagonal parameter list is expanded by the caller */
    float quad_damping[6];
    float linear_damping_forward_speed[36];
    int rk;                                   /* integrator of AUVModel.step (:285-306): 1, 2 or 4 */
} mppi_auv_params;

/* Installs the Fossen model on a handle created with cfg.model = MPPI_MODEL_AUV: x' = normalize_quat(x + rk(x, u)),
 * acc = M^-1 (u - C(nu) nu - D(nu) nu - g(q)) (auv_model.py:285-333,450-559).  The rigid-body mass matrix uses the
 * reference's transposed skew of cog (tf_skew_op, :23-40) and its rk = 4 branch is reproduced as written. */
int mppi_set_auv_model(mppi_handle *h, const mppi_auv_params *prm);
/* The reference's own LEARNED model of the AUV in place of the Fossen equations: NNAUVModel
 * (/root/reference/scripts/src/models/nn_model.py:181-304): X = (concat(state[3:13], action) - Xmean) / Xstd (prepare_data,
 * :289-293: the position is dropped, 16 inputs), n_hidden Dense(hidden, relu) layers and a linear Dense(13) (the reference
 * builds 3 x 32, :54-60), delta = out * Ystd + Ymean, next = state + delta (a plain add: the quaternion is not renormalised,
 * :303-304).  W[l] / b[l], l = 0 .. n_hidden, in Keras layout [in][out]: W[0] [16][hidden], W[l] [hidden][hidden],
 * W[n_hidden] [hidden][13]; Xmean / Xstd [16], Ymean / Ystd [13] (NULL = 0 / 1).  n_hidden = 1..4, hidden = 1..32 (zero padded
 * to 32: exact).  Replaces the model of an MPPI_MODEL_AUV handle; rollout, costs, exchange and mppi_auv_predict work as
 * with the Fossen model.  fp32 CUDA-core kernel (3 kFLOP per sample-step in 32-wide layers is too thin for the 128-row
 * tcgen05 tile of the config-4 MLP kernel). */
int mppi_set_nn_auv_model(mppi_handle *h, int n_hidden, int hidden, const float *const *W, const float *const *b, const float *Xmean,
                          const float *Xstd, const float *Ymean, const float *Ystd);
/* AUVModel.build_step_graph on a batch (the model's predict): state [kst][13], kst in {1, k}; action [k][6]. */
int mppi_auv_predict(mppi_handle *h, int kst, int k, const float *state, const float *action, float *out);
/* StaticQuatCost (scripts/src/costs/static_cost.py:73-159) as the state cost of an AUV handle:
 * d = (p - g_p, 2 acos(q . g_q), nu - g_nu), cost = sum_i q10_i d_i^2 (diagonal Q; the dot product is clamped to
 * [-1, 1], where the reference would return NaN).  mppi_set_static_cost returns to StaticCost (diag q [13]). */
int mppi_set_quat_cost(mppi_handle *h, const float *q10);
int mppi_cost_state_quat(int device, int k, const float *state, const float *goal, const float *q10, float *out);
/* ElipseCost3D (scripts/src/costs/elipse_cost.py:99-246) as the state cost of an AUV handle: the ellipse with half-axes
 * axis = (a, b) lies in the plane with unit `normal` whose first axis is `a_vec` (all [3]); cost of one sample =
 * m_state (|(p'x/a)^2 + (p'y/b)^2 + p'z^2 - 1| + angle between the attitude and the ellipse tangent) + m_vel | |v|^2 -
 * speed^2 |, p' and the attitude expressed in the plane frame.  `center` is accepted and ignored, as the reference's
 * state_cost ignores it; the per-sample sum is the reference's k = 1 result (for k > 1 its own sum mis-broadcasts). */
int mppi_set_ellipse3d_cost(mppi_handle *h, const float *normal, const float *a_vec, const float *axis, const float *center,
                            float speed, float m_state, float m_vel);
int mppi_cost_state_ellipse3d(int device, int k, const float *state, const float *normal, const float *a_vec, const float *axis,
                              const float *center, float speed, float m_state, float m_vel, float *out);

/* ---- stateless stage entry points (the reference's graph-builder methods on plain buffers) ----
 * All pointers are host memory; each call runs the corresponding CUDA kernel on `device`. */
/* utile::blockDiag (src/utile.cpp:10-43): in [rows][cols] -> out [nb*rows][nb*cols] */
int mppi_block_diag(const float *in, int rows, int cols, int nb, float *out);
/* ModelBase::mBuildFreeStepGraph (src/model_base.cpp:59-68): state [kst][s] -> out [kst][s] */
int mppi_model_free_step(int device, float mass, float dt, int s, int a, int kst, const float *state, float *out);
/* ModelBase::mBuildActionStepGraph (src/model_base.cpp:70-82): action [k][a] -> out [k][s] */
int mppi_model_action_step(int device, float mass, float dt, int s, int a, int k, const float *action, float *out);
/* ModelBase::mBuildModelStepGraph (src/model_base.cpp:53-57): state [kst][s], kst in {1,k} */
int mppi_model_step(int device, float mass, float dt, int s, int a, int kst, int k, const float *state,
                    const float *action, float *out);
/* CostBase::mStateCost / mBuildFinalStepCostGraph (src/cost_base.cpp:52-61) */
int mppi_cost_state(int device, int k, int s, const float *state, const float *goal, const float *q, float *out);
/* CostBase.action_cost of the Python twin (scripts/src/costs/cost_base.py:114-170), a <= MPPI_MAX_A:
 * 0.5 [gamma (u^T S^-1 u + 2 u^T S^-1 eps) + lambda (1 - 1/upsilon) eps^T S^-1 eps] */
int mppi_cost_action_py(int device, int k, int a, float lambda, float gamma, float upsilon, const float *sigma,
                        const float *action, const float *noise, float *out);
/* ElipseCost.state_cost (scripts/src/costs/elipse_cost.py:46-79): state [k][4] = (x, vx, y, vy) -> out [k] */
int mppi_cost_state_ellipse(int device, int k, const float *state, float a, float b, float center_x, float center_y,
                            float speed, float m_state, float m_vel, float *out);
/* CostBase::mActionCost (src/cost_base.cpp:63-68): lambda * action^T sigma^-1 noise */
int mppi_cost_action(int device, int k, int a, float lambda, const float *sigma, const float *action,
                     const float *noise, float *out);
/* CostBase::mBuildStepCostGraph (src/cost_base.cpp:43-50) */
int mppi_cost_step(int device, int k, int s, int a, float lambda, const float *sigma, const float *goal,
                   const float *q, const float *state, const float *action, const float *noise, float *out);
/* ControllerBase::mPrepareAction / mPrepareNoise (src/controller_base.cpp:205-213) */
int mppi_prepare_action(int T, int a, const float *U, int t, float *out);
int mppi_prepare_noise(int device, int k, int T, int a, const float *noise, int t, float *out);
/* ControllerBase::mBeta, mExpArg, mExp, mNabla, mWeights, mWeightedNoise
 * (src/controller_base.cpp:166-192); any output pointer may be NULL. */
int mppi_update_stages(int device, int k, int T, int a, float lambda, const float *cost, const float *noise,
                       float *beta, float *exp_arg, float *exp_out, float *nabla, float *weights,
                       float *weighted_noise);
/* The same stages one at a time, as the reference's test chains them (test/test_controller.cpp:146-151):
 *   MPPI_OP_MIN      mBeta     out[0] = min_k in[k]
 *   MPPI_OP_EXP_ARG  mExpArg   out[k] = (-1/s1) * (in[k] - s0)        s0 = beta, s1 = lambda
 *   MPPI_OP_EXP      mExp      out[k] = exp(in[k])
 *   MPPI_OP_SUM      mNabla    out[0] = sum_k in[k]
 *   MPPI_OP_DIV      mWeights  out[k] = in[k] / s0                    s0 = nabla */
typedef enum { MPPI_OP_MIN = 0, MPPI_OP_EXP_ARG = 1, MPPI_OP_EXP = 2, MPPI_OP_SUM = 3, MPPI_OP_DIV = 4 } mppi_stage_op;
int mppi_stage_vector_op(int device, int op, int k, const float *in, float s0, float s1, float *out);
/* ControllerBase::mWeightedNoise alone (src/controller_base.cpp:188-192): weights [k], noise [k][TA] */
int mppi_weighted_noise(int device, int k, int TA, const float *weights, const float *noise, float *out);
/* ControllerBase::mGetNew / mShift (src/controller_base.cpp:310-329) */
int mppi_get_new(int T, int a, const float *cur, int nb, float *out);
int mppi_shift(int T, int a, const float *cur, const float *init, int nb, float *out);
/* Savitzky-Golay smoothing of a sequence [T][a] along T (host code, double precision) - the pass the Python twin runs on its
 * action sequence when `filterSeq` is set: scipy.signal.savgol_filter(seq, window, polyorder, deriv=0, axis=0), mode
 * "interp" (/root/reference/scripts/src/controllers/controller_base.py:281-291 uses window 10, polyorder 9).  Like the
 * reference, the result is a filtered COPY: the controller's own sequence is not touched.  window <= T, polyorder < window. */
int mppi_savgol_filter(int T, int a, const float *U, int window, int polyorder, float *out);
/* Raw Philox4x32-10 words of the noise stream (integer contract, bit-exact vs the oracle):
 * out [n_calls][4] for counter (call0 + i, sample, update, stream). */
int mppi_philox_raw(int device, uint64_t seed, uint32_t call0, uint32_t sample, uint32_t update,
                    uint32_t stream, int n_calls, uint32_t *out);
/* The same with the round count (7 or 10) of mppi_config.philox_rounds. */
int mppi_philox_raw_rounds(int device, uint64_t seed, uint32_t call0, uint32_t sample, uint32_t update,
                           uint32_t stream, int n_calls, int rounds, uint32_t *out);

/* Developer knobs: per-CTA %globaltimer stamps of the update kernel's phases (12 per CTA, nanoseconds: start, tables built,
 * rollout done, weighted sum done, partial published, partials merged, peer payloads in, update applied, then internal stamps)
 * of the last update, [n_controllers][grid.x][12]; mppi_last_grid_x = grid.x of the last update launch. */
int mppi_debug_trace(mppi_handle *h, int on);
int mppi_debug_get_trace(mppi_handle *h, unsigned long long *out, int n_ctas);
int mppi_last_grid_x(const mppi_handle *h);

/* Library/version info: returns a static string such as "mppi_b200 0.1 sm_100a". */
const char *mppi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MPPI_B200_H */
