"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Graph-faithful restatement of the reference's TensorFlow-C++ MPPI graph in torch-CPU:
the same op sequence, the same [K, ., 1] batched-matmul-per-step structure, the same
materialised tensors (eps [K,T,a,1], the w*eps product), T-unrolled, nothing fused.  It
stands in for "TF C++ r2.1 CPU" (which cannot be installed offline) as bench.py's
`--impl reference` arm and as a second, independently written statement of the maths
that tests/ cross-check against the C oracle.

Every op cites the reference line (relative to /root/reference) it restates.
"""
import torch


def block_diag(blk, nb):
    """utile::blockDiag, src/utile.cpp:10-43 (Concat of the block and ZerosLike pads)."""
    pad = torch.zeros_like(blk)
    rows = []
    for i in range(nb):
        rows.append(torch.cat([blk if i == j else pad for j in range(nb)], dim=1))
    return torch.cat(rows, dim=0)


class GraphOracle:
    def __init__(self, k, tau, dt, mass, s_dim, a_dim, lam, sigma, goal, q, dtype=torch.float32):
        self.k, self.tau, self.s, self.a = k, tau, s_dim, a_dim
        self.dtype = dtype
        t = lambda v: torch.as_tensor(v, dtype=dtype)
        self.lam = t(lam)
        self.sigma = t(sigma).reshape(a_dim, a_dim)
        self.goal = t(goal).reshape(s_dim, 1)
        # CostBase::setConsts, src/cost_base.cpp:37-41
        self.inv_sigma = torch.linalg.inv(self.sigma.double()).to(dtype)
        self.Q = torch.diag(t(q).reshape(s_dim))
        # ModelBase::mBuildFreeStepGraph / mBuildActionStepGraph, src/model_base.cpp:59-64,70-78
        self.A = block_diag(t([[1.0, dt], [0.0, 1.0]]), s_dim // 2)
        self.B = block_diag(t([[dt * dt / 2.0], [dt]]) / t(mass), a_dim)
        self.neg_inv_lam = t(-1.0) / self.lam

    def noise(self, z):
        """mNoiseGenGraph, src/controller_base.cpp:194-203: eps = BatchMatMulV2(sigma, rng)."""
        return torch.matmul(self.sigma, z.reshape(self.k, self.tau, self.a, 1))

    def state_cost(self, state):
        """CostBase::mStateCost, src/cost_base.cpp:56-61."""
        diff = state - self.goal
        return torch.matmul(diff.transpose(-1, -2), torch.matmul(self.Q, diff))

    def action_cost(self, action, noise):
        """CostBase::mActionCost, src/cost_base.cpp:63-68."""
        noise_cost = torch.matmul(self.inv_sigma, noise)
        return self.lam * torch.matmul(action.transpose(-1, -2), noise_cost)

    def model_step(self, state, action):
        """ModelBase::mBuildModelStepGraph, src/model_base.cpp:53-82."""
        return torch.matmul(self.A, state) + torch.matmul(self.B, action)

    def rollout(self, x, U, eps):
        """mBuildModelGraph, src/controller_base.cpp:226-273."""
        next_state = x.reshape(1, self.s, 1)                              # :247
        cost = torch.zeros(self.k, 1, 1, dtype=self.dtype)                # :248
        for i in range(self.tau):                                         # :251
            action = U[i:i + 1].squeeze(0)                                # :205-208
            noise = eps[:, i:i + 1].squeeze(1)                            # :210-213 (strided slice)
            to_apply = action + noise                                     # :258
            next_state = self.model_step(next_state, to_apply)            # :260
            tmp = self.state_cost(next_state) + self.action_cost(action, noise)  # :264
            cost = cost + tmp                                             # :268
        return cost + self.state_cost(next_state)                         # :271-272

    def update(self, U, cost, eps):
        """mBuildUpdateGraph, src/controller_base.cpp:166-192,215-224."""
        beta = torch.min(cost, dim=0).values                              # :167
        arg = self.neg_inv_lam * (cost - beta)                            # :171-173
        e = torch.exp(arg)                                                # :177
        nabla = torch.sum(e, dim=0)                                       # :181
        w = e / nabla                                                     # :185
        weighted = torch.sum(w.unsqueeze(-1) * eps, dim=0)                # :189-191 (materialised)
        return U + weighted, dict(beta=beta, exp_arg=arg, exp=e, nabla=nabla, weights=w,
                                  weighted_noise=weighted)

    def next(self, x, U, eps):
        """ControllerBase::next / mBuildGraph, src/controller_base.cpp:135-153,275-308.
        x [s], U [T,a], eps [K,T,a] -> dict(costs [K], U_new, next [a], U_shift)."""
        x = torch.as_tensor(x, dtype=self.dtype)
        U = torch.as_tensor(U, dtype=self.dtype).reshape(self.tau, self.a, 1)
        eps = torch.as_tensor(eps, dtype=self.dtype).reshape(self.k, self.tau, self.a, 1)
        cost = self.rollout(x, U, eps)
        upd, _ = self.update(U, cost, eps)
        nxt = upd[0:1]                                                    # mGetNew :327-329
        init = torch.zeros(1, self.a, 1, dtype=self.dtype)                # mInit0 :310-312
        shifted = torch.cat([upd[1:], init], dim=0)                       # mShift :314-324
        return dict(costs=cost.reshape(-1).numpy(), U_new=upd.reshape(self.tau, self.a).numpy(),
                    next=nxt.reshape(-1).numpy(), U_shift=shifted.reshape(self.tau, self.a).numpy())

    def next_generating(self, x, U, gen):
        """Same as next() but with RandomNormal inside the step (the reference regenerates
        noise on every Run, :196-199); used only for CPU timing."""
        z = torch.randn(self.k, self.tau, self.a, 1, dtype=self.dtype, generator=gen)
        eps = self.noise(z)
        return self.next(x, U, eps)
