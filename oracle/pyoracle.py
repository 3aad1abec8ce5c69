"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes wrapper over oracle/_build/libmppi_oracle.so (see mppi_oracle.c /
mppi_oracle_impl.h for the reference file:line each function restates).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmppi_oracle.so")
_lib = None


def build(force=False):
    """Compile the C oracle (gcc, OpenMP).  Building the checker is not using it."""
    src = [os.path.join(_HERE, f) for f in ("mppi_oracle.c", "mppi_oracle_impl.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    return _SO


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _c(arr, dt):
    return np.ascontiguousarray(arr, dtype=dt)


def _p(arr):
    return arr.ctypes.data_as(C.c_void_p)


class Oracle:
    """One instance per precision ("f32" = the reference's DT_FLOAT graph, "f64" = the
    Python twin's precision, used as the exact value in tolerance tests)."""

    def __init__(self, precision="f32"):
        assert precision in ("f32", "f64")
        self.sfx = "_" + precision
        self.dt = np.float32 if precision == "f32" else np.float64
        self.creal = C.c_float if precision == "f32" else C.c_double
        self.lib = load()

    def _fn(self, name):
        return getattr(self.lib, name + self.sfx)

    # --- utile::blockDiag -------------------------------------------------------------
    def block_diag(self, block, nb):
        block = _c(block, self.dt)
        r, c = block.shape
        out = np.empty((r * nb, c * nb), self.dt)
        self._fn("orc_block_diag")(_p(block), r, c, nb, _p(out))
        return out

    def model_matrices(self, mass, dt, s, a):
        A = np.empty((s, s), self.dt)
        B = np.empty((s, a), self.dt)
        self._fn("orc_model_matrices")(self.creal(mass), self.creal(dt), s, a, _p(A), _p(B))
        return A, B

    # --- ModelBase ----------------------------------------------------------------------
    def model_free_step(self, mass, dt, s, a, state):
        state = _c(state, self.dt).reshape(-1, s)
        out = np.empty_like(state)
        self._fn("orc_model_free_step")(self.creal(mass), self.creal(dt), s, a, state.shape[0],
                                        _p(state), _p(out))
        return out

    def model_action_step(self, mass, dt, s, a, action):
        action = _c(action, self.dt).reshape(-1, a)
        out = np.empty((action.shape[0], s), self.dt)
        self._fn("orc_model_action_step")(self.creal(mass), self.creal(dt), s, a, action.shape[0],
                                          _p(action), _p(out))
        return out

    def model_step(self, mass, dt, s, a, state, action):
        state = _c(state, self.dt).reshape(-1, s)
        action = _c(action, self.dt).reshape(-1, a)
        k = action.shape[0]
        out = np.empty((k, s), self.dt)
        self._fn("orc_model_step")(self.creal(mass), self.creal(dt), s, a, state.shape[0], k,
                                   _p(state), _p(action), _p(out))
        return out

    # --- CostBase -----------------------------------------------------------------------
    def mat_inverse(self, M):
        M = _c(M, self.dt)
        out = np.empty_like(M)
        rc = self._fn("orc_mat_inverse")(_p(M), M.shape[0], _p(out))
        if rc:
            raise np.linalg.LinAlgError("singular")
        return out

    def cost_state(self, state, goal, q):
        goal = _c(goal, self.dt).ravel()
        s = goal.size
        state = _c(state, self.dt).reshape(-1, s)
        q = _c(q, self.dt).ravel()
        out = np.empty(state.shape[0], self.dt)
        self._fn("orc_cost_state")(state.shape[0], s, _p(state), _p(goal), _p(q), _p(out))
        return out

    def cost_action(self, lam, sigma, action, noise):
        action = _c(action, self.dt).ravel()
        a = action.size
        noise = _c(noise, self.dt).reshape(-1, a)
        sigma = _c(sigma, self.dt)
        out = np.empty(noise.shape[0], self.dt)
        self._fn("orc_cost_action")(noise.shape[0], a, self.creal(lam), _p(sigma), _p(action),
                                    _p(noise), _p(out))
        return out

    def cost_step(self, lam, sigma, goal, q, state, action, noise):
        goal = _c(goal, self.dt).ravel()
        s = goal.size
        action = _c(action, self.dt).ravel()
        a = action.size
        state = _c(state, self.dt).reshape(-1, s)
        noise = _c(noise, self.dt).reshape(-1, a)
        out = np.empty(state.shape[0], self.dt)
        self._fn("orc_cost_step")(state.shape[0], s, a, self.creal(lam), _p(_c(sigma, self.dt)),
                                  _p(goal), _p(_c(q, self.dt).ravel()), _p(state), _p(action),
                                  _p(noise), _p(out))
        return out

    # --- ControllerBase stages ------------------------------------------------------------
    def scale_noise(self, sigma, z):
        sigma = _c(sigma, self.dt)
        a = sigma.shape[0]
        z = _c(z, self.dt)
        out = np.empty_like(z)
        self._fn("orc_scale_noise")(z.size // a, a, _p(sigma), _p(z), _p(out))
        return out

    def prepare_action(self, U, t):
        U = _c(U, self.dt)
        T, a = U.shape
        out = np.empty(a, self.dt)
        self._fn("orc_prepare_action")(T, a, _p(U), t, _p(out))
        return out

    def prepare_noise(self, noise, t):
        noise = _c(noise, self.dt)
        k, T, a = noise.shape
        out = np.empty((k, a), self.dt)
        self._fn("orc_prepare_noise")(k, T, a, _p(noise), t, _p(out))
        return out

    def update_stages(self, lam, cost, noise):
        cost = _c(cost, self.dt).ravel()
        noise = _c(noise, self.dt)
        k, T, a = noise.shape
        beta = self.creal()
        nabla = self.creal()
        arg = np.empty(k, self.dt)
        e = np.empty(k, self.dt)
        w = np.empty(k, self.dt)
        wn = np.empty((T, a), self.dt)
        self._fn("orc_update_stages")(k, T, a, self.creal(lam), _p(cost), _p(noise), C.byref(beta),
                                      _p(arg), _p(e), C.byref(nabla), _p(w), _p(wn))
        return dict(beta=beta.value, exp_arg=arg, exp=e, nabla=nabla.value, weights=w,
                    weighted_noise=wn)

    def get_new(self, cur, nb):
        cur = _c(cur, self.dt)
        T, a = cur.shape
        out = np.empty((nb, a), self.dt)
        self._fn("orc_get_new")(T, a, _p(cur), nb, _p(out))
        return out

    def shift(self, cur, init, nb):
        cur = _c(cur, self.dt)
        init = _c(init, self.dt)
        T, a = cur.shape
        out = np.empty((T, a), self.dt)
        self._fn("orc_shift")(T, a, _p(cur), _p(init), nb, _p(out))
        return out

    def rollout_costs(self, cfg, x0, U, eps):
        k, T, s, a = cfg["k"], cfg["tau"], cfg["s_dim"], cfg["a_dim"]
        eps = _c(eps, self.dt)
        assert eps.shape == (k, T, a)
        costs = np.empty(k, self.dt)
        self._fn("orc_rollout_costs")(0, k, T, s, a, self.creal(cfg["dt"]), self.creal(cfg["mass"]),
                                      self.creal(cfg["lambda"]), _p(_c(cfg["sigma"], self.dt)),
                                      _p(_c(cfg["goal"], self.dt)), _p(_c(cfg["q"], self.dt)),
                                      _p(_c(x0, self.dt)), _p(_c(U, self.dt)), _p(eps), _p(costs))
        return costs

    def mppi_update(self, cfg, x0, U, eps):
        """Full update (ControllerBase::next graph).  Returns dict(costs, U_new, next, U_shift)."""
        k, T, s, a = cfg["k"], cfg["tau"], cfg["s_dim"], cfg["a_dim"]
        eps = _c(eps, self.dt)
        assert eps.shape == (k, T, a)
        costs = np.empty(k, self.dt)
        U_new = np.empty((T, a), self.dt)
        nxt = np.empty(a, self.dt)
        U_shift = np.empty((T, a), self.dt)
        self._fn("orc_mppi_update")(k, T, s, a, self.creal(cfg["dt"]), self.creal(cfg["mass"]),
                                    self.creal(cfg["lambda"]), _p(_c(cfg["sigma"], self.dt)),
                                    _p(_c(cfg["goal"], self.dt)), _p(_c(cfg["q"], self.dt)),
                                    _p(_c(x0, self.dt)), _p(_c(U, self.dt)), _p(eps), _p(costs),
                                    _p(U_new), _p(nxt), _p(U_shift))
        return dict(costs=costs, U_new=U_new, next=nxt, U_shift=U_shift)

    def cost_action_py(self, lam, gamma, upsilon, sigma, action, noise):
        """Python-twin action cost (scripts/src/costs/cost_base.py:114-170)."""
        noise = _c(noise, self.dt)
        k, a = noise.shape
        out = np.empty(k, self.dt)
        self._fn("orc_cost_action_py")(k, a, self.creal(lam), self.creal(gamma), self.creal(upsilon),
                                       _p(_c(sigma, self.dt)), _p(_c(action, self.dt)), _p(noise), _p(out))
        return out

    def cost_state_ellipse(self, state, ell):
        """ElipseCost.state_cost (scripts/src/costs/elipse_cost.py:46-79); ell = (a, b, cx, cy, speed, m_state, m_vel)."""
        state = _c(np.asarray(state).reshape(-1, 4), self.dt)
        out = np.empty(state.shape[0], self.dt)
        self._fn("orc_cost_state_ellipse")(state.shape[0], _p(state), _p(_c(ell, self.dt)), _p(out))
        return out

    def mppi_update_py(self, cfg, x0, U, eps, gamma=None, upsilon=1.0, normalize=False, ellipse=None):
        """Python-twin update (controller_base.py:371-474): gamma / upsilon action cost, optional cost
        normalisation, optional ElipseCost state cost (ellipse = (a, b, cx, cy, speed, m_state, m_vel)).
        eps is the already scaled noise (upsilon * sigma) z."""
        k, T, s, a = cfg["k"], cfg["tau"], cfg["s_dim"], cfg["a_dim"]
        eps = _c(eps, self.dt)
        assert eps.shape == (k, T, a)
        gamma = cfg["lambda"] if gamma is None else gamma
        costs = np.empty(k, self.dt)
        U_new = np.empty((T, a), self.dt)
        nxt = np.empty(a, self.dt)
        U_shift = np.empty((T, a), self.dt)
        self._fn("orc_mppi_update_py")(k, T, s, a, self.creal(cfg["dt"]), self.creal(cfg["mass"]),
                                       self.creal(cfg["lambda"]), self.creal(gamma), self.creal(upsilon),
                                       int(bool(normalize)), _p(_c(cfg["sigma"], self.dt)),
                                       _p(_c(cfg["goal"], self.dt)), _p(_c(cfg["q"], self.dt)),
                                       _p(_c(ellipse, self.dt)) if ellipse is not None else None,
                                       _p(_c(x0, self.dt)), _p(_c(U, self.dt)), _p(eps), _p(costs),
                                       _p(U_new), _p(nxt), _p(U_shift))
        return dict(costs=costs, U_new=U_new, next=nxt, U_shift=U_shift)

    @staticmethod
    def auv_pack(prm):
        """Flat primitive-parameter vector of the AUV model (oracle/mppi_oracle_impl.h, ORC_AUV_NPRM = 127)."""
        sq = lambda v: (np.diag(v) if np.ndim(v) == 1 else np.asarray(v, np.float64)).ravel()
        i = prm["inertial"]
        return np.concatenate([[prm["mass"], prm["volume"], prm["density"]], prm["cog"], prm["cob"],
                               np.asarray(prm["Ma"], np.float64).ravel(),
                               [i["ixx"], i["iyy"], i["izz"], i["ixy"], i["ixz"], i["iyz"]],
                               sq(np.asarray(prm["linear_damping"], np.float64)), np.asarray(prm["quad_damping"], np.float64),
                               sq(np.asarray(prm["linear_damping_forward_speed"], np.float64))]).astype(np.float64)

    @staticmethod
    def nn_auv_pack(nn):
        """Parameter vector of the learned AUV model (orc_nn_auv_step): nn = dict(W=[W0, ..], b=[b0, ..], Xmean, Xstd, Ymean,
        Ystd) with Keras-layout weights, W0 [16, H], .., W_last [H, 13]."""
        W, b = nn["W"], nn["b"]
        parts = [[len(W) - 1, np.asarray(W[0]).shape[1]], nn["Xmean"], nn["Xstd"], nn["Ymean"], nn["Ystd"]]
        for Wl, bl in zip(W, b):
            parts += [np.asarray(Wl, np.float64).ravel(), np.asarray(bl, np.float64).ravel()]
        return np.concatenate([np.asarray(v, np.float64).ravel() for v in parts])

    def nn_auv_step(self, nn, state, action):
        """NNAUVModel.build_step_graph (scripts/src/models/nn_model.py:215-239): state [k,13], action [k,6] -> [k,13]."""
        st, ac = _c(np.asarray(state).reshape(-1, 13), self.dt), _c(np.asarray(action).reshape(-1, 6), self.dt)
        out = np.empty_like(st)
        self._fn("orc_nn_auv_step")(st.shape[0], _p(_c(self.nn_auv_pack(nn), self.dt)), _p(st), _p(ac), _p(out))
        return out

    def mppi_update_nn_auv(self, nn, lam, sigma, goal, q, x0, U, eps, gamma=None, upsilon=1.0, normalize=False, quat_cost=False):
        """Python-controller update with the learned AUV model in place of AUVModel (rk = 0 in the C restatement)."""
        eps = _c(eps, self.dt)
        k, T, a = eps.shape
        gamma = lam if gamma is None else gamma
        Q = _c(np.diag(np.asarray(q, np.float64)) if quat_cost else np.asarray(q, np.float64), self.dt)
        costs, U_new, nxt, U_shift = np.empty(k, self.dt), np.empty((T, a), self.dt), np.empty(a, self.dt), np.empty((T, a), self.dt)
        self._fn("orc_mppi_update_auv")(k, T, _p(_c(self.nn_auv_pack(nn), self.dt)), self.creal(0.1), 0, self.creal(lam),
                                        self.creal(gamma), self.creal(upsilon), int(bool(normalize)),
                                        _p(_c(sigma, self.dt)), _p(_c(np.asarray(goal).ravel(), self.dt)), _p(Q),
                                        int(bool(quat_cost)), _p(_c(x0, self.dt)), _p(_c(U, self.dt)), _p(eps), _p(costs),
                                        _p(U_new), _p(nxt), _p(U_shift))
        return dict(costs=costs, U_new=U_new, next=nxt, U_shift=U_shift)

    def auv_step(self, prm, dt, rk, state, action):
        """AUVModel.step (scripts/src/models/auv_model.py:285-306): state [k,13], action [k,6] -> [k,13]."""
        st, ac = _c(np.asarray(state).reshape(-1, 13), self.dt), _c(np.asarray(action).reshape(-1, 6), self.dt)
        out = np.empty_like(st)
        self._fn("orc_auv_step")(st.shape[0], _p(_c(self.auv_pack(prm), self.dt)), self.creal(dt), int(rk), _p(st), _p(ac), _p(out))
        return out

    def auv_state_dot(self, prm, state, action):
        st, ac = _c(np.asarray(state).reshape(-1, 13), self.dt), _c(np.asarray(action).reshape(-1, 6), self.dt)
        out = np.empty_like(st)
        self._fn("orc_auv_state_dot_k")(st.shape[0], _p(_c(self.auv_pack(prm), self.dt)), _p(st), _p(ac), _p(out))
        return out

    def cost_state_quat(self, state, goal, Q):
        """StaticQuatCost.state_cost (scripts/src/costs/static_cost.py:116-159); Q [10][10] or its diagonal [10]."""
        st = _c(np.asarray(state).reshape(-1, 13), self.dt)
        Q = np.asarray(Q, np.float64)
        Q = np.diag(Q) if Q.ndim == 1 else Q
        out = np.empty(st.shape[0], self.dt)
        self._fn("orc_cost_state_quat")(st.shape[0], _p(st), _p(_c(np.asarray(goal).ravel(), self.dt)), _p(_c(Q, self.dt)), _p(out))
        return out

    def ellipse3d_prep(self, normal, aVec):
        """ElipseCost3D.prepare_consts (elipse_cost.py:150-154): returns (R [3,3], q [4] as (x, y, z, w))."""
        R, q = np.empty((3, 3), self.dt), np.empty(4, self.dt)
        self._fn("orc_ellipse3d_prep")(_p(_c(np.asarray(normal).ravel(), self.dt)), _p(_c(np.asarray(aVec).ravel(), self.dt)), _p(R), _p(q))
        return R, q

    def ellipse3d_pack(self, normal, aVec, axis, speed, m_state, m_vel):
        _, q = self.ellipse3d_prep(normal, aVec)
        return np.concatenate([q, np.asarray(axis, np.float64).ravel(), [speed, m_state, m_vel]])

    def cost_state_ellipse3d(self, state, e3):
        """ElipseCost3D.state_cost, one state at a time (elipse_cost.py:156-170); e3 from ellipse3d_pack."""
        st = _c(np.asarray(state).reshape(-1, 13), self.dt)
        out = np.empty(st.shape[0], self.dt)
        self._fn("orc_cost_state_ellipse3d")(st.shape[0], _p(st), _p(_c(e3, self.dt)), _p(out))
        return out

    def mppi_update_auv(self, prm, dt, rk, lam, sigma, goal, q, x0, U, eps, gamma=None, upsilon=1.0, normalize=False,
                        quat_cost=False, ellipse3d=None):
        """Python-controller update with the AUV model (controller_base.py:371-474, auv_model.py:285-306) and
        StaticCost (q [13]) or StaticQuatCost (q [10], quat_cost=True).  eps = (upsilon * sigma) z, [k][T][6]."""
        eps = _c(eps, self.dt)
        k, T, a = eps.shape
        assert a == 6
        gamma = lam if gamma is None else gamma
        Q = _c(np.diag(np.asarray(q, np.float64)) if quat_cost else np.asarray(q, np.float64), self.dt)
        kind = int(bool(quat_cost))
        if ellipse3d is not None:                       # ElipseCost3D parameters from ellipse3d_pack replace goal / q
            Q, kind = _c(ellipse3d, self.dt), 2
        costs = np.empty(k, self.dt)
        U_new = np.empty((T, a), self.dt)
        nxt = np.empty(a, self.dt)
        U_shift = np.empty((T, a), self.dt)
        self._fn("orc_mppi_update_auv")(k, T, _p(_c(self.auv_pack(prm), self.dt)), self.creal(dt), int(rk), self.creal(lam),
                                        self.creal(gamma), self.creal(upsilon), int(bool(normalize)),
                                        _p(_c(sigma, self.dt)), _p(_c(np.asarray(goal).ravel(), self.dt)), _p(Q),
                                        kind, _p(_c(x0, self.dt)), _p(_c(U, self.dt)), _p(eps), _p(costs),
                                        _p(U_new), _p(nxt), _p(U_shift))
        return dict(costs=costs, U_new=U_new, next=nxt, U_shift=U_shift)

    def partial(self, lam, costs, eps, k0, k1):
        costs = _c(costs, self.dt)
        eps = _c(eps, self.dt)
        _, T, a = eps.shape
        beta = self.creal()
        eta = self.creal()
        N = np.empty((T, a), self.dt)
        self._fn("orc_partial")(k0, k1, T, a, self.creal(lam), _p(costs), _p(eps), C.byref(beta),
                                C.byref(eta), _p(N))
        return beta.value, eta.value, N

    # --- MLP dynamics (row A13) -----------------------------------------------------------
    def _mlp_args(self, mlp):
        keys = ("W1", "b1", "W2", "b2", "W3", "b3", "Xmean", "Xstd", "Ymean", "Ystd")
        arrs = [_c(mlp[k], self.dt) for k in keys]
        return arrs

    def mlp_step(self, mlp, x, u):
        x = _c(x, self.dt).ravel()
        u = _c(u, self.dt).ravel()
        s, a, H = x.size, u.size, np.asarray(mlp["b1"]).size
        arrs = self._mlp_args(mlp)
        out = np.empty(s, self.dt)
        self._fn("orc_mlp_step")(s, a, H, *[_p(v) for v in arrs], _p(x), _p(u), _p(out))
        return out

    def mppi_update_mlp(self, cfg, mlp, x0, U, eps):
        k, T, s, a = cfg["k"], cfg["tau"], cfg["s_dim"], cfg["a_dim"]
        H = np.asarray(mlp["b1"]).size
        eps = _c(eps, self.dt)
        arrs = self._mlp_args(mlp)
        costs = np.empty(k, self.dt)
        U_new = np.empty((T, a), self.dt)
        nxt = np.empty(a, self.dt)
        U_shift = np.empty((T, a), self.dt)
        self._fn("orc_mppi_update_mlp")(k, T, s, a, H, self.creal(cfg["lambda"]),
                                        _p(_c(cfg["sigma"], self.dt)), _p(_c(cfg["goal"], self.dt)),
                                        _p(_c(cfg["q"], self.dt)), *[_p(v) for v in arrs],
                                        _p(_c(x0, self.dt)), _p(_c(U, self.dt)), _p(eps), _p(costs),
                                        _p(U_new), _p(nxt), _p(U_shift))
        return dict(costs=costs, U_new=U_new, next=nxt, U_shift=U_shift)


# --- noise stream specification (integer part is a bit-exact contract) ---------------------
def philox4x32_10(ctr, key, rounds=10):
    lib = load()
    c = (C.c_uint32 * 4)(*[int(v) & 0xFFFFFFFF for v in ctr])
    k = (C.c_uint32 * 2)(*[int(v) & 0xFFFFFFFF for v in key])
    o = (C.c_uint32 * 4)()
    lib.orc_philox4x32_r(c, k, C.c_int(rounds), o)
    return [int(v) for v in o]


def philox_normals(seed, update, stream, k0, k1, n_per_sample, rounds=10):
    lib = load()
    z = np.empty((k1 - k0, n_per_sample), np.float32)
    lib.orc_philox_normals_r(C.c_uint64(seed), C.c_uint32(update), C.c_uint32(stream),
                             C.c_uint32(k0), C.c_uint32(k1), C.c_int(n_per_sample), C.c_int(rounds), _p(z))
    return z


def num_threads():
    return load().orc_num_threads()
