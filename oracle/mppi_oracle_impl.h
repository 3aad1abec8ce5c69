/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement of the NicolayP/mppi-tf update step, one C function per
 * reference sub-graph, written op-for-op so that every known-answer test of
 * the reference (test/test_{model,cost,controller,utile}.cpp) can be replayed
 * against it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may use it.
 *
 * This file is included twice by mppi_oracle.c: once with REAL=float
 * (the reference's DT_FLOAT graph) and once with REAL=double (the Python
 * twin's fp64 graph, used as the "exact" value in tolerance tests).
 *
 * All file:line citations are relative to /root/reference.
 *
 * Layouts (row-major, the reference's trailing singleton dim dropped):
 *   state  [k][s]      reference [k, s, 1]
 *   action [k][a]      reference [k, a, 1]
 *   noise  [k][T][a]   reference [k, T, a, 1]
 *   U      [T][a]      reference [T, a, 1]
 */

#ifndef REAL
#error "include from mppi_oracle.c"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* utile::blockDiag — src/utile.cpp:10-43.  `in` is [rows][cols]; the result is
 * the nb-fold block diagonal [nb*rows][nb*cols] (zero padding elsewhere). */
void FN(orc_block_diag)(const REAL *in, int rows, int cols, int nb, REAL *out)
{
    int R = rows * nb, C = cols * nb;
    for (int i = 0; i < R * C; i++) out[i] = (REAL)0;
    for (int b = 0; b < nb; b++)
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++)
                out[(b * rows + r) * C + (b * cols + c)] = in[r * cols + c];
}

/* A = blockDiag([[1,dt],[0,1]], s/2) — src/model_base.cpp:59-64.
 * B = blockDiag([[dt*dt/2],[dt]] / mass, a) — src/model_base.cpp:70-78; the
 * mass is the constructor argument as pinned by test/test_model.cpp:123-125
 * (the "mass" Variable of :27-32 is never assigned at HEAD). */
void FN(orc_model_matrices)(REAL mass, REAL dt, int s, int a, REAL *A, REAL *B)
{
    REAL a_blk[4] = {(REAL)1, dt, (REAL)0, (REAL)1};
    REAL b_blk[2] = {(dt * dt) / (REAL)2 / mass, dt / mass};
    FN(orc_block_diag)(a_blk, 2, 2, s / 2, A);
    FN(orc_block_diag)(b_blk, 2, 1, a, B);
}

/* BatchMatMul(M[r][c], v[k|1][c][1]) with the matrix broadcast over the batch
 * (model_base.cpp:65-67, :79-81).  Dense on purpose: the reference multiplies
 * by the full block-diagonal matrix, zeros included. */
static void FN(orc_bmm_bcast)(const REAL *M, int r, int c, const REAL *v, int k, REAL *out)
{
    for (int i = 0; i < k; i++)
        for (int row = 0; row < r; row++) {
            REAL acc = (REAL)0;
            for (int col = 0; col < c; col++) acc += M[row * c + col] * v[i * c + col];
            out[i * r + row] = acc;
        }
}

/* ModelBase::mBuildFreeStepGraph — src/model_base.cpp:59-68.  A x, state [kst][s]. */
void FN(orc_model_free_step)(REAL mass, REAL dt, int s, int a, int kst, const REAL *state, REAL *out)
{
    REAL A[ORC_MAX_S * ORC_MAX_S], B[ORC_MAX_S * ORC_MAX_A];
    FN(orc_model_matrices)(mass, dt, s, a, A, B);
    FN(orc_bmm_bcast)(A, s, s, state, kst, out);
}

/* ModelBase::mBuildActionStepGraph — src/model_base.cpp:70-82.  (B/m) u, action [k][a]. */
void FN(orc_model_action_step)(REAL mass, REAL dt, int s, int a, int k, const REAL *action, REAL *out)
{
    REAL A[ORC_MAX_S * ORC_MAX_S], B[ORC_MAX_S * ORC_MAX_A];
    FN(orc_model_matrices)(mass, dt, s, a, A, B);
    FN(orc_bmm_bcast)(B, s, a, action, k, out);
}

/* ModelBase::mBuildModelStepGraph — src/model_base.cpp:53-57.  state is
 * [kst][s] with kst in {1, k}: kst == 1 broadcasts against [k][a] actions
 * (test/test_model.cpp:216-255, src/controller_base.cpp:247). out [k][s]. */
void FN(orc_model_step)(REAL mass, REAL dt, int s, int a, int kst, int k,
                        const REAL *state, const REAL *action, REAL *out)
{
    REAL A[ORC_MAX_S * ORC_MAX_S], B[ORC_MAX_S * ORC_MAX_A];
    REAL fr[ORC_MAX_S], ac[ORC_MAX_S];
    FN(orc_model_matrices)(mass, dt, s, a, A, B);
    for (int i = 0; i < k; i++) {
        const REAL *st = state + (kst == 1 ? 0 : i * s);
        FN(orc_bmm_bcast)(A, s, s, st, 1, fr);
        FN(orc_bmm_bcast)(B, s, a, action + i * a, 1, ac);
        for (int j = 0; j < s; j++) out[i * s + j] = fr[j] + ac[j];
    }
}

/* MatrixInverse — src/cost_base.cpp:39.  Gauss-Jordan with partial pivoting,
 * carried out in double so both instantiations share one inverse. */
int FN(orc_mat_inverse)(const REAL *M, int n, REAL *inv)
{
    double w[ORC_MAX_A][2 * ORC_MAX_A];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            w[i][j] = (double)M[i * n + j];
            w[i][n + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int r = c + 1; r < n; r++)
            if (fabs(w[r][c]) > fabs(w[p][c])) p = r;
        if (w[p][c] == 0.0) return 1;
        if (p != c)
            for (int j = 0; j < 2 * n; j++) { double t = w[c][j]; w[c][j] = w[p][j]; w[p][j] = t; }
        double d = w[c][c];
        for (int j = 0; j < 2 * n; j++) w[c][j] /= d;
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            double f = w[r][c];
            if (f == 0.0) continue;
            for (int j = 0; j < 2 * n; j++) w[r][j] -= f * w[c][j];
        }
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) inv[i * n + j] = (REAL)w[i][n + j];
    return 0;
}

/* CostBase::mStateCost — src/cost_base.cpp:56-61 with Q = Diag(q) (:40):
 * diff = x - g; left = Q diff (dense); cost = diff^T left. */
void FN(orc_cost_state)(int k, int s, const REAL *state, const REAL *goal, const REAL *q, REAL *out)
{
    for (int i = 0; i < k; i++) {
        REAL diff[ORC_MAX_S], left[ORC_MAX_S];
        for (int j = 0; j < s; j++) diff[j] = state[i * s + j] - goal[j];
        for (int r = 0; r < s; r++) {
            REAL acc = (REAL)0;
            for (int c = 0; c < s; c++) acc += ((r == c) ? q[r] : (REAL)0) * diff[c];
            left[r] = acc;
        }
        REAL acc = (REAL)0;
        for (int j = 0; j < s; j++) acc += diff[j] * left[j];
        out[i] = acc;
    }
}

/* CostBase::mActionCost — src/cost_base.cpp:63-68:
 * lambda * action^T (Sigma^-1 noise); action [a] is the UN-perturbed U[t]. */
void FN(orc_cost_action)(int k, int a, REAL lambda, const REAL *sigma, const REAL *action,
                         const REAL *noise, REAL *out)
{
    REAL inv[ORC_MAX_A * ORC_MAX_A];
    FN(orc_mat_inverse)(sigma, a, inv);
    for (int i = 0; i < k; i++) {
        REAL nc[ORC_MAX_A];
        FN(orc_bmm_bcast)(inv, a, a, noise + i * a, 1, nc);
        REAL acc = (REAL)0;
        for (int j = 0; j < a; j++) acc += action[j] * nc[j];
        out[i] = lambda * acc;
    }
}

/* CostBase::mBuildStepCostGraph — src/cost_base.cpp:43-50. */
void FN(orc_cost_step)(int k, int s, int a, REAL lambda, const REAL *sigma, const REAL *goal,
                       const REAL *q, const REAL *state, const REAL *action, const REAL *noise,
                       REAL *out)
{
    REAL *ac = (REAL *)malloc(sizeof(REAL) * (size_t)k);
    FN(orc_cost_state)(k, s, state, goal, q, out);
    FN(orc_cost_action)(k, a, lambda, sigma, action, noise, ac);
    for (int i = 0; i < k; i++) out[i] = out[i] + ac[i];
    free(ac);
}

/* mNoiseGenGraph scaling — src/controller_base.cpp:201: eps = Sigma z. */
void FN(orc_scale_noise)(int n, int a, const REAL *sigma, const REAL *z, REAL *eps)
{
    FN(orc_bmm_bcast)(sigma, a, a, z, n, eps);
}

/* mPrepareAction / mPrepareNoise — src/controller_base.cpp:205-213. */
void FN(orc_prepare_action)(int T, int a, const REAL *U, int t, REAL *out)
{
    (void)T;
    for (int j = 0; j < a; j++) out[j] = U[t * a + j];
}
void FN(orc_prepare_noise)(int k, int T, int a, const REAL *noise, int t, REAL *out)
{
    for (int i = 0; i < k; i++)
        for (int j = 0; j < a; j++) out[i * a + j] = noise[((size_t)i * T + t) * a + j];
}

/* mBeta / mExpArg / mExp / mNabla / mWeights / mWeightedNoise —
 * src/controller_base.cpp:166-192.  Any output pointer may be NULL. */
void FN(orc_update_stages)(int k, int T, int a, REAL lambda, const REAL *cost, const REAL *noise,
                           REAL *beta_o, REAL *arg_o, REAL *exp_o, REAL *nabla_o, REAL *w_o,
                           REAL *wn_o)
{
    REAL beta = cost[0];
    for (int i = 1; i < k; i++) beta = cost[i] < beta ? cost[i] : beta;     /* Min :167 */
    REAL *e = (REAL *)malloc(sizeof(REAL) * (size_t)k);
    REAL nabla = (REAL)0;
    REAL neg_inv_lambda = (REAL)(-1) / lambda;                               /* :171 */
    for (int i = 0; i < k; i++) {
        REAL arg = neg_inv_lambda * (cost[i] - beta);                        /* :171-173 */
        if (arg_o) arg_o[i] = arg;
        e[i] = REAL_EXP(arg);                                                /* :177 */
        if (exp_o) exp_o[i] = e[i];
        nabla += e[i];                                                       /* :181 */
    }
    if (beta_o) *beta_o = beta;
    if (nabla_o) *nabla_o = nabla;
    if (wn_o) for (int j = 0; j < T * a; j++) wn_o[j] = (REAL)0;
    for (int i = 0; i < k; i++) {
        REAL w = e[i] / nabla;                                               /* :185 */
        if (w_o) w_o[i] = w;
        if (wn_o)
            for (int j = 0; j < T * a; j++) wn_o[j] += w * noise[(size_t)i * T * a + j]; /* :189-191 */
    }
    free(e);
}

/* mGetNew — src/controller_base.cpp:327-329: first nb rows. */
void FN(orc_get_new)(int T, int a, const REAL *cur, int nb, REAL *out)
{
    (void)T;
    for (int j = 0; j < nb * a; j++) out[j] = cur[j];
}

/* mShift — src/controller_base.cpp:310-324: concat(cur[nb:], init[nb]). */
void FN(orc_shift)(int T, int a, const REAL *cur, const REAL *init, int nb, REAL *out)
{
    for (int j = 0; j < (T - nb) * a; j++) out[j] = cur[nb * a + j];
    for (int j = 0; j < nb * a; j++) out[(T - nb) * a + j] = init[j];
}

/* Rollout + cost for samples [k0, k1) — mBuildModelGraph, src/controller_base.cpp:226-273:
 *   x <- x0 (broadcast, :247); S <- 0 (:248)
 *   for t: u = U[t] + eps[:,t] (:256-258); x <- A x + (B/m) u (:260);
 *          S += q(x) + lambda U[t]^T Sigma^-1 eps[:,t] (:264-268, un-noised U[t])
 *   S += q(x_T) (:271-272, on top of step T-1's q(x_T)).
 * eps_sample_stride lets the caller pass a slice of a larger [K][T][a] tensor. */
void FN(orc_rollout_costs)(int k0, int k1, int T, int s, int a, REAL dt, REAL mass, REAL lambda,
                           const REAL *sigma, const REAL *goal, const REAL *q, const REAL *x0,
                           const REAL *U, const REAL *eps, REAL *costs)
{
    REAL A[ORC_MAX_S * ORC_MAX_S], B[ORC_MAX_S * ORC_MAX_A], inv[ORC_MAX_A * ORC_MAX_A];
    FN(orc_model_matrices)(mass, dt, s, a, A, B);
    FN(orc_mat_inverse)(sigma, a, inv);
    for (int i = k0; i < k1; i++) {
        REAL x[ORC_MAX_S], xn[ORC_MAX_S], u[ORC_MAX_A], nc[ORC_MAX_A], fr[ORC_MAX_S], ac[ORC_MAX_S];
        REAL S = (REAL)0, c;
        for (int j = 0; j < s; j++) x[j] = x0[j];
        for (int t = 0; t < T; t++) {
            const REAL *e = eps + ((size_t)i * T + t) * a;
            const REAL *ut = U + t * a;
            for (int j = 0; j < a; j++) u[j] = ut[j] + e[j];
            FN(orc_bmm_bcast)(A, s, s, x, 1, fr);
            FN(orc_bmm_bcast)(B, s, a, u, 1, ac);
            for (int j = 0; j < s; j++) xn[j] = fr[j] + ac[j];
            FN(orc_cost_state)(1, s, xn, goal, q, &c);
            FN(orc_bmm_bcast)(inv, a, a, e, 1, nc);
            REAL acc = (REAL)0;
            for (int j = 0; j < a; j++) acc += ut[j] * nc[j];
            S = S + (c + lambda * acc);
            for (int j = 0; j < s; j++) x[j] = xn[j];
        }
        FN(orc_cost_state)(1, s, x, goal, q, &c);
        costs[i] = S + c;
    }
}

/* One full update — ControllerBase::next's graph, src/controller_base.cpp:135-153,
 * 275-308: rollout costs -> update stages -> U' = U + sum_k w_k eps_k (:223) ->
 * next = U'[0] (:303) -> shifted = concat(U'[1:], 0) (:305-307).
 * Outputs: costs [k], U_new [T][a] (pre-shift), next [a], U_shift [T][a]. */
void FN(orc_mppi_update)(int k, int T, int s, int a, REAL dt, REAL mass, REAL lambda,
                         const REAL *sigma, const REAL *goal, const REAL *q, const REAL *x0,
                         const REAL *U, const REAL *eps, REAL *costs, REAL *U_new, REAL *next,
                         REAL *U_shift)
{
    REAL *wn = (REAL *)malloc(sizeof(REAL) * (size_t)T * a);
    REAL *zero = (REAL *)calloc((size_t)a, sizeof(REAL));
#pragma omp parallel
    {
        int nt = 1, id = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads();
        id = omp_get_thread_num();
#endif
        int lo = (int)((long long)k * id / nt), hi = (int)((long long)k * (id + 1) / nt);
        FN(orc_rollout_costs)(lo, hi, T, s, a, dt, mass, lambda, sigma, goal, q, x0, U, eps, costs);
    }
    FN(orc_update_stages)(k, T, a, lambda, costs, eps, NULL, NULL, NULL, NULL, NULL, wn);
    for (int j = 0; j < T * a; j++) U_new[j] = U[j] + wn[j];
    FN(orc_get_new)(T, a, U_new, 1, next);
    FN(orc_shift)(T, a, U_new, zero, 1, U_shift);
    free(wn);
    free(zero);
}

/* Rank-partial of the update over samples [k0,k1): (beta_r, eta_r, N_r[T*a]) with
 * N_r = sum_k exp(-(S_k-beta_r)/lambda) eps_k — the payload exchanged between ranks
 * (SURVEY.md section 8e).  New design, no reference line; checked against
 * orc_update_stages by tests/test_oracle_kats.py. */
void FN(orc_partial)(int k0, int k1, int T, int a, REAL lambda, const REAL *costs, const REAL *eps,
                     REAL *beta_r, REAL *eta_r, REAL *N_r)
{
    REAL beta = costs[k0];
    for (int i = k0 + 1; i < k1; i++) beta = costs[i] < beta ? costs[i] : beta;
    REAL eta = (REAL)0;
    for (int j = 0; j < T * a; j++) N_r[j] = (REAL)0;
    for (int i = k0; i < k1; i++) {
        REAL e = REAL_EXP(-(costs[i] - beta) / lambda);
        eta += e;
        for (int j = 0; j < T * a; j++) N_r[j] += e * eps[(size_t)i * T * a + j];
    }
    *beta_r = beta;
    *eta_r = eta;
}

/* Learned-MLP dynamics step (row A13): behaviour of NNAUVModel.build_step_graph,
 * scripts/src/models/nn_model.py:215-239,289-304 with the Keras Sequential of :54-60
 * re-shaped to BASELINE config 4 (input [x(s), u(a)] -> H -> H -> s, ReLU, linear head):
 *   X = (concat(x,u) - Xmean)/Xstd; h1 = relu(W1^T X + b1); h2 = relu(W2^T h1 + b2);
 *   d = W3^T h2 + b3; x' = x + d*Ystd + Ymean.
 * Weights are Keras-layout [in][out] row-major.  PARITY UNPINNED in the reference (no
 * forward-value KAT exists for this shape; scripts/test.py:592-682 only covers data prep). */
void FN(orc_mlp_step)(int s, int a, int H, const REAL *W1, const REAL *b1, const REAL *W2,
                      const REAL *b2, const REAL *W3, const REAL *b3, const REAL *Xmean,
                      const REAL *Xstd, const REAL *Ymean, const REAL *Ystd, const REAL *x,
                      const REAL *u, REAL *xn)
{
    REAL X[ORC_MAX_S + ORC_MAX_A], h1[ORC_MAX_H], h2[ORC_MAX_H];
    int in = s + a;
    for (int j = 0; j < s; j++) X[j] = (x[j] - Xmean[j]) / Xstd[j];
    for (int j = 0; j < a; j++) X[s + j] = (u[j] - Xmean[s + j]) / Xstd[s + j];
    for (int o = 0; o < H; o++) {
        REAL acc = b1[o];
        for (int i = 0; i < in; i++) acc += X[i] * W1[i * H + o];
        h1[o] = acc > (REAL)0 ? acc : (REAL)0;
    }
    for (int o = 0; o < H; o++) {
        REAL acc = b2[o];
        for (int i = 0; i < H; i++) acc += h1[i] * W2[i * H + o];
        h2[o] = acc > (REAL)0 ? acc : (REAL)0;
    }
    for (int o = 0; o < s; o++) {
        REAL acc = b3[o];
        for (int i = 0; i < H; i++) acc += h2[i] * W3[i * s + o];
        xn[o] = x[o] + (acc * Ystd[o] + Ymean[o]);
    }
}

/* Rollout costs with the MLP dynamics in place of the point mass (same cost and
 * loop structure as orc_rollout_costs). */
void FN(orc_rollout_costs_mlp)(int k0, int k1, int T, int s, int a, int H, REAL lambda,
                               const REAL *sigma, const REAL *goal, const REAL *q, const REAL *W1,
                               const REAL *b1, const REAL *W2, const REAL *b2, const REAL *W3,
                               const REAL *b3, const REAL *Xmean, const REAL *Xstd,
                               const REAL *Ymean, const REAL *Ystd, const REAL *x0, const REAL *U,
                               const REAL *eps, REAL *costs)
{
    REAL inv[ORC_MAX_A * ORC_MAX_A];
    FN(orc_mat_inverse)(sigma, a, inv);
    for (int i = k0; i < k1; i++) {
        REAL x[ORC_MAX_S], xn[ORC_MAX_S], u[ORC_MAX_A], nc[ORC_MAX_A];
        REAL S = (REAL)0, c;
        for (int j = 0; j < s; j++) x[j] = x0[j];
        for (int t = 0; t < T; t++) {
            const REAL *e = eps + ((size_t)i * T + t) * a;
            const REAL *ut = U + t * a;
            for (int j = 0; j < a; j++) u[j] = ut[j] + e[j];
            FN(orc_mlp_step)(s, a, H, W1, b1, W2, b2, W3, b3, Xmean, Xstd, Ymean, Ystd, x, u, xn);
            FN(orc_cost_state)(1, s, xn, goal, q, &c);
            FN(orc_bmm_bcast)(inv, a, a, e, 1, nc);
            REAL acc = (REAL)0;
            for (int j = 0; j < a; j++) acc += ut[j] * nc[j];
            S = S + (c + lambda * acc);
            for (int j = 0; j < s; j++) x[j] = xn[j];
        }
        FN(orc_cost_state)(1, s, x, goal, q, &c);
        costs[i] = S + c;
    }
}

void FN(orc_mppi_update_mlp)(int k, int T, int s, int a, int H, REAL lambda, const REAL *sigma,
                             const REAL *goal, const REAL *q, const REAL *W1, const REAL *b1,
                             const REAL *W2, const REAL *b2, const REAL *W3, const REAL *b3,
                             const REAL *Xmean, const REAL *Xstd, const REAL *Ymean,
                             const REAL *Ystd, const REAL *x0, const REAL *U, const REAL *eps,
                             REAL *costs, REAL *U_new, REAL *next, REAL *U_shift)
{
    REAL *wn = (REAL *)malloc(sizeof(REAL) * (size_t)T * a);
    REAL *zero = (REAL *)calloc((size_t)a, sizeof(REAL));
#pragma omp parallel
    {
        int nt = 1, id = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads();
        id = omp_get_thread_num();
#endif
        int lo = (int)((long long)k * id / nt), hi = (int)((long long)k * (id + 1) / nt);
        FN(orc_rollout_costs_mlp)(lo, hi, T, s, a, H, lambda, sigma, goal, q, W1, b1, W2, b2, W3, b3,
                                  Xmean, Xstd, Ymean, Ystd, x0, U, eps, costs);
    }
    FN(orc_update_stages)(k, T, a, lambda, costs, eps, NULL, NULL, NULL, NULL, NULL, wn);
    for (int j = 0; j < T * a; j++) U_new[j] = U[j] + wn[j];
    FN(orc_get_new)(T, a, U_new, 1, next);
    FN(orc_shift)(T, a, U_new, zero, 1, U_shift);
    free(wn);
    free(zero);
}

/* ---- Python-twin variants ---------------------------------------
 * CostBase.action_cost of the Python controller — scripts/src/costs/cost_base.py:114-170:
 *   0.5 * [ gamma * (u^T S^-1 u + 2 u^T S^-1 eps) + lambda * (1 - 1/upsilon) * eps^T S^-1 eps ]
 * with u the UN-perturbed action and S = sigma (the unscaled covariance factor). */
void FN(orc_cost_action_py)(int k, int a, REAL lambda, REAL gamma, REAL upsilon, const REAL *sigma,
                            const REAL *action, const REAL *noise, REAL *out)
{
    REAL inv[ORC_MAX_A * ORC_MAX_A], rhsA[ORC_MAX_A];
    FN(orc_mat_inverse)(sigma, a, inv);
    FN(orc_bmm_bcast)(inv, a, a, action, 1, rhsA);                      /* rhsAcost :137-138 */
    REAL aCost = (REAL)0;
    for (int j = 0; j < a; j++) aCost += action[j] * rhsA[j];           /* :151-152 */
    aCost = gamma * aCost;                                              /* :154-155 */
    for (int i = 0; i < k; i++) {
        REAL rhsN[ORC_MAX_A];
        FN(orc_bmm_bcast)(inv, a, a, noise + i * a, 1, rhsN);           /* rhsNcost :134-135 */
        REAL mix = (REAL)0, nC = (REAL)0;
        for (int j = 0; j < a; j++) {
            mix += action[j] * rhsN[j];                                 /* :141-142 */
            nC += noise[i * a + j] * rhsN[j];                           /* :147-148 */
        }
        mix = gamma * ((REAL)2 * mix);                                  /* :144,157-158 */
        nC = (lambda * ((REAL)1 - (REAL)1 / upsilon)) * nC;             /* :160-162 */
        out[i] = (REAL)0.5 * ((aCost + mix) + nC);                      /* :165-167 */
    }
}

/* ElipseCost.state_cost — scripts/src/costs/elipse_cost.py:46-79; state [k][4] = (x, vx, y, vy);
 * ell = {a, b, cx, cy, speed, m_state, m_vel}:
 *   v = sqrt(vx^2 + vy^2); d = |((x-cx)/a)^2 + ((y-cy)/b)^2 - 1|; cost = m_state d + m_vel (v - speed)^2 */
void FN(orc_cost_state_ellipse)(int k, const REAL *state, const REAL *ell, REAL *out)
{
    for (int i = 0; i < k; i++) {
        REAL x = state[4 * i], vx = state[4 * i + 1], y = state[4 * i + 2], vy = state[4 * i + 3];
        REAL v = (REAL)sqrt((double)(vx * vx + vy * vy));                       /* :68 */
        REAL diffx = (x - ell[2]) / ell[0], diffy = (y - ell[3]) / ell[1];      /* :69-70 */
        REAL d = diffx * diffx + diffy * diffy - (REAL)1;                       /* :71 */
        d = ell[5] * (d < (REAL)0 ? -d : d);                                    /* :71-72 */
        REAL dv = ell[6] * ((v - ell[4]) * (v - ell[4]));                       /* :73-74 */
        out[i] = d + dv;                                                        /* :75 */
    }
}

/* Python-twin rollout costs — scripts/src/controllers/controller_base.py:371-434 with
 * PointMassModel (point_mass_model.py:66-151) and StaticCost.state_cost (static_cost.py:40-63):
 * same loop as orc_rollout_costs with the action cost above; eps is build_noise's output
 * (upsilon * sigma) z (:348-369). */
void FN(orc_rollout_costs_py)(int k0, int k1, int T, int s, int a, REAL dt, REAL mass, REAL lambda,
                              REAL gamma, REAL upsilon, const REAL *sigma, const REAL *goal,
                              const REAL *q, const REAL *ell /* NULL: StaticCost; else ElipseCost, s == 4 */,
                              const REAL *x0, const REAL *U, const REAL *eps, REAL *costs)
{
    REAL A[ORC_MAX_S * ORC_MAX_S], B[ORC_MAX_S * ORC_MAX_A];
    FN(orc_model_matrices)(mass, dt, s, a, A, B);
    for (int i = k0; i < k1; i++) {
        REAL x[ORC_MAX_S], xn[ORC_MAX_S], u[ORC_MAX_A], fr[ORC_MAX_S], ac[ORC_MAX_S];
        REAL S = (REAL)0, c, acst;
        for (int j = 0; j < s; j++) x[j] = x0[j];
        for (int t = 0; t < T; t++) {
            const REAL *e = eps + ((size_t)i * T + t) * a;
            const REAL *ut = U + t * a;
            for (int j = 0; j < a; j++) u[j] = ut[j] + e[j];
            FN(orc_bmm_bcast)(A, s, s, x, 1, fr);
            FN(orc_bmm_bcast)(B, s, a, u, 1, ac);
            for (int j = 0; j < s; j++) xn[j] = fr[j] + ac[j];
            if (ell) FN(orc_cost_state_ellipse)(1, xn, ell, &c);
            else FN(orc_cost_state)(1, s, xn, goal, q, &c);
            FN(orc_cost_action_py)(1, a, lambda, gamma, upsilon, sigma, ut, e, &acst);
            S = S + (c + acst);
            for (int j = 0; j < s; j++) x[j] = xn[j];
        }
        if (ell) FN(orc_cost_state_ellipse)(1, x, ell, &c);
        else FN(orc_cost_state)(1, s, x, goal, q, &c);
        costs[i] = c + S;                                               /* add_cost(fCost, cost) :428 */
    }
}

/* Python-twin update — controller_base.py:436-474: beta = min S; arg = S - beta, divided by its
 * maximum when `normalize` (norm_arg :468-474); e = exp(-arg/lambda); w = e / sum e;
 * U' = U + sum_k w_k eps_k; next = U'[0]; shifted = concat(U'[1:], 0) (:547-560). */
static void FN(orc_update_tail_py)(int k, int T, int a, REAL lambda, int normalize, const REAL *costs, const REAL *U,
                                   const REAL *eps, REAL *U_new, REAL *next, REAL *U_shift)
{
    REAL *zero = (REAL *)calloc((size_t)a, sizeof(REAL));
    REAL *e = (REAL *)malloc(sizeof(REAL) * (size_t)k);
    REAL beta = costs[0];
    for (int i = 1; i < k; i++) beta = costs[i] < beta ? costs[i] : beta;
    REAL mx = (REAL)0;
    for (int i = 0; i < k; i++) mx = (costs[i] - beta) > mx ? (costs[i] - beta) : mx;
    REAL nabla = (REAL)0;
    for (int i = 0; i < k; i++) {
        REAL arg = costs[i] - beta;
        if (normalize) arg = arg / mx;
        e[i] = REAL_EXP(((REAL)(-1) / lambda) * arg);
        nabla += e[i];
    }
    for (int j = 0; j < T * a; j++) U_new[j] = (REAL)0;
    for (int i = 0; i < k; i++) {
        REAL w = e[i] / nabla;
        for (int j = 0; j < T * a; j++) U_new[j] += w * eps[(size_t)i * T * a + j];
    }
    for (int j = 0; j < T * a; j++) U_new[j] = U[j] + U_new[j];
    FN(orc_get_new)(T, a, U_new, 1, next);
    FN(orc_shift)(T, a, U_new, zero, 1, U_shift);
    free(e);
    free(zero);
}

void FN(orc_mppi_update_py)(int k, int T, int s, int a, REAL dt, REAL mass, REAL lambda, REAL gamma,
                            REAL upsilon, int normalize, const REAL *sigma, const REAL *goal,
                            const REAL *q, const REAL *ell, const REAL *x0, const REAL *U, const REAL *eps,
                            REAL *costs, REAL *U_new, REAL *next, REAL *U_shift)
{
    FN(orc_rollout_costs_py)(0, k, T, s, a, dt, mass, lambda, gamma, upsilon, sigma, goal, q, ell, x0, U, eps, costs);
    FN(orc_update_tail_py)(k, T, a, lambda, normalize, costs, U, eps, U_new, next, U_shift);
}

/* ---- AUV (Fossen) dynamics — scripts/src/models/auv_model.py ------------
 * prm (primitive parameters, 127 values): mass, volume, density, cog[3], cob[3], Ma[36], inertia {ixx, iyy,
 * izz, ixy, ixz, iyz}, linear_damping[36], quad_damping[6], linear_damping_forward_speed[36].
 * State x = (p[3], q = (qx, qy, qz, qw), nu[6]); action u[6] = generalised force. */
#define ORC_AUV_NPRM 127
static void FN(orc_auv_mass)(const REAL *prm, REAL *Mtot /*36*/, REAL *invM /*36*/)
{
    const REAL m = prm[0];
    const REAL *cog = prm + 3, *Ma = prm + 9, *in = prm + 45;
    /* tf_skew_op (:23-40) concatenates its three rows along axis 1, i.e. as COLUMNS: it returns the transposed
     * skew matrix.  Restated as written (it only matters for an off-centre centre of gravity). */
    REAL S[9] = {0, cog[2], -cog[1], -cog[2], 0, cog[0], cog[1], -cog[0], 0};
    REAL I[9] = {in[0], in[3], in[4], in[3], in[1], in[5], in[4], in[5], in[2]};      /* get_inertial :265-280 */
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {                                                 /* rigid_body_mass :257-260 */
            Mtot[r * 6 + c] = (r == c) ? m : (REAL)0;
            Mtot[r * 6 + 3 + c] = -(m * S[r * 3 + c]);
            Mtot[(3 + r) * 6 + c] = m * S[r * 3 + c];
            Mtot[(3 + r) * 6 + 3 + c] = I[r * 3 + c];
        }
    for (int i = 0; i < 36; i++) Mtot[i] = Mtot[i] + Ma[i];                           /* total_mass :262-263 */
    FN(orc_mat_inverse)(Mtot, 6, invM);                                               /* :238 */
}

/* state_dot — auv_model.py:308-333 with body2inertial_transform :353-398, get_jacobian :335-351,
 * damping_matrix :478-506, coriolis_matrix :508-542, restoring_forces :450-476, acc :544-559 */
static void FN(orc_auv_state_dot)(const REAL *prm, const REAL *Mtot, const REAL *invM, const REAL *x, const REAL *u,
                                  REAL *xd /*13*/)
{
    const REAL m = prm[0], vol = prm[1], rho = prm[2], g = (REAL)9.81;
    const REAL *cog = prm + 3, *cob = prm + 6, *Dl = prm + 51, *dq = prm + 87, *Dlf = prm + 93;
    const REAL qx = x[3], qy = x[4], qz = x[5], qw = x[6];
    const REAL *nu = x + 7;
    REAL R[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw),
                 2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw),
                 2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)};
    REAL T[12] = {qw, -qz, qy, qz, qw, -qx, -qy, qx, qw, -qx, -qy, -qz};             /* rows x, y, z, w; times 0.5 */
    for (int r = 0; r < 3; r++) xd[r] = R[r * 3] * nu[0] + R[r * 3 + 1] * nu[1] + R[r * 3 + 2] * nu[2];
    for (int r = 0; r < 4; r++)
        xd[3 + r] = (REAL)0.5 * T[r * 3] * nu[3] + (REAL)0.5 * T[r * 3 + 1] * nu[4] + (REAL)0.5 * T[r * 3 + 2] * nu[5];
    /* D nu */
    REAL Dv[6], Cv[6], gv[6], rhs[6];
    for (int r = 0; r < 6; r++) {
        REAL acc = (REAL)0;
        for (int c = 0; c < 6; c++) {
            REAL d = -Dl[r * 6 + c] - nu[0] * Dlf[r * 6 + c];
            if (r == c) d = d + (-(dq[r] * (nu[r] < 0 ? -nu[r] : nu[r])));
            acc += d * nu[c];
        }
        Dv[r] = acc;
    }
    /* C nu: C = [[0, S12], [S12, S22]], S12 = -skew(M11 nu1 + M12 nu2), S22 = -skew(M21 nu1 + M22 nu2) */
    REAL a1[3], a2[3];
    for (int r = 0; r < 3; r++) {
        a1[r] = (Mtot[r * 6] * nu[0] + Mtot[r * 6 + 1] * nu[1] + Mtot[r * 6 + 2] * nu[2]) +
                (Mtot[r * 6 + 3] * nu[3] + Mtot[r * 6 + 4] * nu[4] + Mtot[r * 6 + 5] * nu[5]);
        a2[r] = (Mtot[(3 + r) * 6] * nu[0] + Mtot[(3 + r) * 6 + 1] * nu[1] + Mtot[(3 + r) * 6 + 2] * nu[2]) +
                (Mtot[(3 + r) * 6 + 3] * nu[3] + Mtot[(3 + r) * 6 + 4] * nu[4] + Mtot[(3 + r) * 6 + 5] * nu[5]);
    }
    REAL S12[9] = {0, a1[2], -a1[1], -a1[2], 0, a1[0], a1[1], -a1[0], 0};             /* -skew(a1) */
    REAL S22[9] = {0, a2[2], -a2[1], -a2[2], 0, a2[0], a2[1], -a2[0], 0};
    for (int r = 0; r < 3; r++) {
        Cv[r] = S12[r * 3] * nu[3] + S12[r * 3 + 1] * nu[4] + S12[r * 3 + 2] * nu[5];
        Cv[3 + r] = (S12[r * 3] * nu[0] + S12[r * 3 + 1] * nu[1] + S12[r * 3 + 2] * nu[2]) +
                    (S22[r * 3] * nu[3] + S22[r * 3 + 1] * nu[4] + S22[r * 3 + 2] * nu[5]);
    }
    /* restoring: fbg = R^T (0,0,-m g), fbb = R^T (0,0,V rho g); g = -[fbg + fbb; cog x fbg + cob x fbb] */
    REAL fng = -(m * g), fnb = vol * rho * g, fbg[3], fbb[3];
    for (int r = 0; r < 3; r++) { fbg[r] = R[6 + r] * fng; fbb[r] = R[6 + r] * fnb; }
    REAL mbg[3] = {cog[1] * fbg[2] - cog[2] * fbg[1], cog[2] * fbg[0] - cog[0] * fbg[2], cog[0] * fbg[1] - cog[1] * fbg[0]};
    REAL mbb[3] = {cob[1] * fbb[2] - cob[2] * fbb[1], cob[2] * fbb[0] - cob[0] * fbb[2], cob[0] * fbb[1] - cob[1] * fbb[0]};
    for (int r = 0; r < 3; r++) { gv[r] = -(fbg[r] + fbb[r]); gv[3 + r] = -(mbg[r] + mbb[r]); }
    for (int r = 0; r < 6; r++) rhs[r] = ((u[r] - Cv[r]) - Dv[r]) - gv[r];           /* :555 */
    for (int r = 0; r < 6; r++) {
        REAL acc = (REAL)0;
        for (int c = 0; c < 6; c++) acc += invM[r * 6 + c] * rhs[c];
        xd[7 + r] = acc;
    }
}

/* AUVModel.step — :285-306 (rk = 1, 2; the rk = 4 branch of the reference multiplies k4 by dt twice and is
 * restated as written) followed by normalize_quat :426-448 (tf.math.l2_normalize, epsilon 1e-12). */
/* NNAUVModel.build_step_graph — /root/reference/scripts/src/models/nn_model.py:215-239 with prepare_data :289-293,
 * denormalizeY :295-297 and next_state :303-304:
 *   X = (concat(state[3:13], action) - Xmean) / Xstd ; h = relu(h W_l + b_l) for every hidden layer ; d = (h W_o + b_o) Ystd + Ymean
 *   next = state + d                      (the quaternion is NOT renormalised)
 * nn = [n_hidden, H, Xmean[16], Xstd[16], Ymean[13], Ystd[13], then per layer W [in][out] (Keras layout), b [out]]:
 * in = 16 for the first layer, out = H for the hidden layers and 13 for the last one. */
void FN(orc_nn_auv_step)(int k, const REAL *nn, const REAL *state, const REAL *action, REAL *out)
{
    const int nh = (int)nn[0], H = (int)nn[1];
    const REAL *Xmean = nn + 2, *Xstd = Xmean + 16, *Ymean = Xstd + 16, *Ystd = Ymean + 13, *W0 = Ystd + 13;
    for (int i = 0; i < k; i++) {
        const REAL *x = state + 13 * i, *u = action + 6 * i;
        REAL a0[ORC_MAX_H], a1[ORC_MAX_H];
        for (int j = 0; j < 10; j++) a0[j] = (x[3 + j] - Xmean[j]) / Xstd[j];
        for (int j = 0; j < 6; j++) a0[10 + j] = (u[j] - Xmean[10 + j]) / Xstd[10 + j];
        const REAL *W = W0;
        int n_in = 16;
        for (int l = 0; l <= nh; l++) {
            const int n_out = (l == nh) ? 13 : H;
            const REAL *b = W + n_in * n_out;
            for (int o = 0; o < n_out; o++) {
                REAL acc = b[o];
                for (int q = 0; q < n_in; q++) acc += a0[q] * W[q * n_out + o];
                a1[o] = (l == nh) ? acc : (acc > (REAL)0 ? acc : (REAL)0);
            }
            for (int o = 0; o < n_out; o++) a0[o] = a1[o];
            W = b + n_out;
            n_in = n_out;
        }
        for (int j = 0; j < 13; j++) out[13 * i + j] = x[j] + (a0[j] * Ystd[j] + Ymean[j]);
    }
}

/* rk = 0 selects the learned model: `prm` is then the nn vector of orc_nn_auv_step. */
void FN(orc_auv_step)(int k, const REAL *prm, REAL dt, int rk, const REAL *state, const REAL *action, REAL *out)
{
    if (rk == 0) {
        FN(orc_nn_auv_step)(k, prm, state, action, out);
        return;
    }
    REAL Mtot[36], invM[36];
    FN(orc_auv_mass)(prm, Mtot, invM);
    for (int i = 0; i < k; i++) {
        const REAL *x = state + 13 * i, *u = action + 6 * i;
        REAL k1[13], k2[13], k3[13], k4[13], tmp[13], xs[13];
        FN(orc_auv_state_dot)(prm, Mtot, invM, x, u, k1);
        if (rk == 2) {
            for (int j = 0; j < 13; j++) xs[j] = x[j] + dt * k1[j];
            FN(orc_auv_state_dot)(prm, Mtot, invM, xs, u, k2);
            for (int j = 0; j < 13; j++) tmp[j] = dt / (REAL)2 * (k1[j] + k2[j]);
        } else if (rk == 4) {
            for (int j = 0; j < 13; j++) xs[j] = x[j] + dt * k1[j] / (REAL)2;
            FN(orc_auv_state_dot)(prm, Mtot, invM, xs, u, k2);
            for (int j = 0; j < 13; j++) xs[j] = x[j] + dt * k2[j] / (REAL)2;
            FN(orc_auv_state_dot)(prm, Mtot, invM, xs, u, k3);
            for (int j = 0; j < 13; j++) xs[j] = x[j] + dt * k3[j];
            FN(orc_auv_state_dot)(prm, Mtot, invM, xs, u, k4);
            for (int j = 0; j < 13; j++)
                tmp[j] = (REAL)(1. / 6.) * ((k1[j] + (REAL)2 * k2[j]) + ((REAL)2 * k3[j] + k4[j] * dt)) * dt;
        } else {
            for (int j = 0; j < 13; j++) tmp[j] = k1[j] * dt;
        }
        REAL *o = out + 13 * i;
        for (int j = 0; j < 13; j++) o[j] = x[j] + tmp[j];
        REAL n2 = o[3] * o[3] + o[4] * o[4] + o[5] * o[5] + o[6] * o[6];
        REAL inv = (REAL)1 / (REAL)sqrt((double)(n2 > (REAL)1e-12 ? n2 : (REAL)1e-12));
        for (int j = 3; j < 7; j++) o[j] = o[j] * inv;
    }
}

/* the state derivative alone, for the component fixtures (rotation / damping / Coriolis / restoring enter it) */
void FN(orc_auv_state_dot_k)(int k, const REAL *prm, const REAL *state, const REAL *action, REAL *out)
{
    REAL Mtot[36], invM[36];
    FN(orc_auv_mass)(prm, Mtot, invM);
    for (int i = 0; i < k; i++) FN(orc_auv_state_dot)(prm, Mtot, invM, state + 13 * i, action + 6 * i, out + 13 * i);
}

/* StaticQuatCost.state_cost — scripts/src/costs/static_cost.py:116-159: d = (p - g_p, 2 acos(q . g_q), nu - g_nu),
 * cost = d^T Q d with Q [10][10] (full, as the reference multiplies it). */
void FN(orc_cost_state_quat)(int k, const REAL *state, const REAL *goal, const REAL *Q, REAL *out)
{
    for (int i = 0; i < k; i++) {
        const REAL *x = state + 13 * i;
        REAL d[10], r[10];
        REAL dot = ((x[3] * goal[3] + x[4] * goal[4]) + x[5] * goal[5]) + x[6] * goal[6];     /* tensordot :148 */
        for (int j = 0; j < 3; j++) d[j] = x[j] - goal[j];
        d[3] = (REAL)2 * (REAL)acos((double)dot);
        for (int j = 0; j < 6; j++) d[4 + j] = x[7 + j] - goal[7 + j];
        for (int a = 0; a < 10; a++) {
            r[a] = (REAL)0;
            for (int b = 0; b < 10; b++) r[a] += Q[a * 10 + b] * d[b];
        }
        REAL c = (REAL)0;
        for (int a = 0; a < 10; a++) c += d[a] * r[a];
        out[i] = c;
    }
}

/* ElipseCost3D — scripts/src/costs/elipse_cost.py:99-246.  tensorflow_graphics (un-vendored, unpinned; imported at
 * elipse_cost.py:3) supplies the quaternion algebra; its published algorithms are restated here, quaternions (x, y, z, w):
 * multiply = Hamilton product, rotate(p, q) = q (p, 0) q*, from_rotation_matrix = the four-branch trace method,
 * between_two_vectors_3d(v1, v2) = normalise(cross(v1, v2), 1 + v1.v2) with the antiparallel fallback,
 * relative_angle = 2 acos(|q1.q2|).  Pinned by the reference's known-answer tests (scripts/test.py:1183-1359) through
 * tests/golden/gen_ellipse3d_fixtures.py.
 * prepare_consts (:150-154): N = [aVec, normal x aVec, normal] (columns), R = inv(N)^T, q = from_rotation_matrix(R). */
static void FN(orc_quat_mul)(const REAL *a, const REAL *b, REAL *o)
{
    o[0] = ((a[0] * b[3] + a[1] * b[2]) - a[2] * b[1]) + a[3] * b[0];
    o[1] = ((-a[0] * b[2] + a[1] * b[3]) + a[2] * b[0]) + a[3] * b[1];
    o[2] = ((a[0] * b[1] - a[1] * b[0]) + a[2] * b[3]) + a[3] * b[2];
    o[3] = ((-a[0] * b[0] - a[1] * b[1]) - a[2] * b[2]) + a[3] * b[3];
}

void FN(orc_ellipse3d_prep)(const REAL *normal, const REAL *aVec, REAL *R /*9*/, REAL *q /*4*/)
{
    REAL b[3] = {normal[1] * aVec[2] - normal[2] * aVec[1], normal[2] * aVec[0] - normal[0] * aVec[2],
                 normal[0] * aVec[1] - normal[1] * aVec[0]};                                /* bVec = normal x aVec :137-142 */
    REAL N[9] = {aVec[0], b[0], normal[0], aVec[1], b[1], normal[1], aVec[2], b[2], normal[2]}, inv[9];
    FN(orc_mat_inverse)(N, 3, inv);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) R[r * 3 + c] = inv[c * 3 + r];                          /* transpose :152 */
    REAL tr = R[0] + R[4] + R[8], sq;
    if (tr > (REAL)0) {
        sq = (REAL)sqrt((double)(tr + (REAL)1)) * (REAL)2;
        q[0] = (R[7] - R[5]) / sq; q[1] = (R[2] - R[6]) / sq; q[2] = (R[3] - R[1]) / sq; q[3] = (REAL)0.25 * sq;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        sq = (REAL)sqrt((double)((REAL)1 + R[0] - R[4] - R[8])) * (REAL)2;
        q[0] = (REAL)0.25 * sq; q[1] = (R[1] + R[3]) / sq; q[2] = (R[2] + R[6]) / sq; q[3] = (R[7] - R[5]) / sq;
    } else if (R[4] > R[8]) {
        sq = (REAL)sqrt((double)((REAL)1 + R[4] - R[0] - R[8])) * (REAL)2;
        q[0] = (R[1] + R[3]) / sq; q[1] = (REAL)0.25 * sq; q[2] = (R[5] + R[7]) / sq; q[3] = (R[2] - R[6]) / sq;
    } else {
        sq = (REAL)sqrt((double)((REAL)1 + R[8] - R[0] - R[4])) * (REAL)2;
        q[0] = (R[2] + R[6]) / sq; q[1] = (R[5] + R[7]) / sq; q[2] = (REAL)0.25 * sq; q[3] = (R[3] - R[1]) / sq;
    }
}

/* state_cost (:156-170) of ONE state at a time (for k > 1 the reference's sum broadcasts [k,1,1] + [k] to [k,1,k]; the
 * per-sample meaning is the k = 1 call).  e3 = {q[4], a, b, speed, mS, mV}.  `center` never enters state_cost in the
 * reference (self.t is stored and not used) and does not here. */
void FN(orc_cost_state_ellipse3d)(int k, const REAL *state, const REAL *e3, REAL *out)
{
    const REAL *q = e3, a = e3[4], b = e3[5], gv = e3[6], mS = e3[7], mV = e3[8];
    for (int i = 0; i < k; i++) {
        const REAL *x = state + 13 * i;
        REAL p4[4] = {x[0], x[1], x[2], (REAL)0}, qc[4] = {-q[0], -q[1], -q[2], q[3]}, t[4], pf[4], qpf[4];
        FN(orc_quat_mul)(q, p4, t);                                                         /* rotate :160 */
        FN(orc_quat_mul)(t, qc, pf);
        FN(orc_quat_mul)(q, x + 3, qpf);                                                    /* :162 */
        REAL d = (pf[0] / a) * (pf[0] / a) + (pf[1] / b) * (pf[1] / b) + (pf[2] / (REAL)1) * (pf[2] / (REAL)1);   /* :188-190 */
        d = d - (REAL)1;
        REAL pos = d < (REAL)0 ? -d : d;
        /* orientation_error :193-219: tangent (-a/b y, b/a x, 0), normalised; q_t = between((1,0,0), tangent) */
        REAL tg[3] = {pf[1] * (-a / b), pf[0] * (b / a), (REAL)0};
        REAL nrm = (REAL)sqrt((double)(tg[0] * tg[0] + tg[1] * tg[1] + tg[2] * tg[2]));
        for (int j = 0; j < 3; j++) tg[j] = tg[j] / nrm;
        REAL n2 = tg[0] * tg[0] + tg[1] * tg[1] + tg[2] * tg[2];
        REAL in = (REAL)1 / (REAL)sqrt((double)(n2 > (REAL)1e-12 ? n2 : (REAL)1e-12));      /* l2_normalize of both inputs */
        for (int j = 0; j < 3; j++) tg[j] = tg[j] * in;
        REAL real = (REAL)1 + tg[0];                                                        /* 1 + x . tg */
        REAL rot[4] = {(REAL)0, -tg[2], tg[1], real};                                       /* cross((1,0,0), tg) */
        if (real < (REAL)1e-6) { rot[0] = (REAL)0; rot[1] = (REAL)0; rot[2] = (REAL)1; rot[3] = (REAL)0; }   /* |x| > |y|: (-z, 0, x) of (1,0,0) */
        REAL r2 = rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2] + rot[3] * rot[3];
        REAL ir = (REAL)1 / (REAL)sqrt((double)(r2 > (REAL)1e-12 ? r2 : (REAL)1e-12));
        REAL dot = (REAL)0;
        for (int j = 0; j < 4; j++) dot += (rot[j] * ir) * qpf[j];
        dot = dot < (REAL)0 ? -dot : dot;
        if (dot > (REAL)1) dot = (REAL)1;
        REAL ori = (REAL)2 * (REAL)acos((double)dot);                                       /* relative_angle */
        REAL v2 = x[7] * x[7] + x[8] * x[8] + x[9] * x[9];                                  /* :236-238: | |v|^2 - gv^2 | */
        REAL vn = (REAL)sqrt((double)v2);
        REAL dv = vn * vn - gv * gv;
        dv = dv < (REAL)0 ? -dv : dv;
        out[i] = (mS * pos + mS * ori) + mV * dv;                                           /* :168 */
    }
}

/* Python controller rollout with the AUV model — controller_base.py:371-434 with AUVModel.build_step_graph
 * (auv_model.py:285-306) and StaticCost (quat_cost = 0: Q = diag(q[13])), StaticQuatCost (quat_cost = 1:
 * Q [10][10]) or ElipseCost3D (quat_cost = 2: Q = e3[9]); action cost cost_base.py:114-170; eps = (upsilon * sigma) z. */
void FN(orc_rollout_costs_auv)(int k0, int k1, int T, const REAL *prm, REAL dt, int rk, REAL lambda, REAL gamma,
                               REAL upsilon, const REAL *sigma, const REAL *goal, const REAL *Q, int quat_cost,
                               const REAL *x0, const REAL *U, const REAL *eps, REAL *costs)
{
    const int s = 13, a = 6;
    for (int i = k0; i < k1; i++) {
        REAL x[13], xn[13], u[6];
        REAL S = (REAL)0, c, acst;
        for (int j = 0; j < s; j++) x[j] = x0[j];
        for (int t = 0; t < T; t++) {
            const REAL *e = eps + ((size_t)i * T + t) * a;
            const REAL *ut = U + t * a;
            for (int j = 0; j < a; j++) u[j] = ut[j] + e[j];
            FN(orc_auv_step)(1, prm, dt, rk, x, u, xn);
            if (quat_cost == 2) FN(orc_cost_state_ellipse3d)(1, xn, Q, &c);
            else if (quat_cost) FN(orc_cost_state_quat)(1, xn, goal, Q, &c);
            else FN(orc_cost_state)(1, s, xn, goal, Q, &c);
            FN(orc_cost_action_py)(1, a, lambda, gamma, upsilon, sigma, ut, e, &acst);
            S = S + (c + acst);
            for (int j = 0; j < s; j++) x[j] = xn[j];
        }
        if (quat_cost == 2) FN(orc_cost_state_ellipse3d)(1, x, Q, &c);
        else if (quat_cost) FN(orc_cost_state_quat)(1, x, goal, Q, &c);
        else FN(orc_cost_state)(1, s, x, goal, Q, &c);
        costs[i] = c + S;
    }
}

void FN(orc_mppi_update_auv)(int k, int T, const REAL *prm, REAL dt, int rk, REAL lambda, REAL gamma, REAL upsilon,
                             int normalize, const REAL *sigma, const REAL *goal, const REAL *Q, int quat_cost,
                             const REAL *x0, const REAL *U, const REAL *eps, REAL *costs, REAL *U_new, REAL *next,
                             REAL *U_shift)
{
#pragma omp parallel
    {
        int nt = 1, id = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads();
        id = omp_get_thread_num();
#endif
        int lo = (int)((long long)k * id / nt), hi = (int)((long long)k * (id + 1) / nt);
        FN(orc_rollout_costs_auv)(lo, hi, T, prm, dt, rk, lambda, gamma, upsilon, sigma, goal, Q, quat_cost, x0, U, eps, costs);
    }
    FN(orc_update_tail_py)(k, T, 6, lambda, normalize, costs, U, eps, U_new, next, U_shift);
}


#undef FN
#undef CAT
#undef CAT_
