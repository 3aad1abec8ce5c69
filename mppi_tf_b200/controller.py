"""Python mirror of the reference's C++ classes over the C-ABI (used by tests and bench.py).

Same names and argument meaning as /root/reference/include/{controller,model,cost}_base.hpp;
the TensorFlow graph-builder methods become calls on plain numpy buffers.  All arithmetic runs in
libmppi_b200.so on the GPU — nothing here computes.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import MppiConfig, check

_fp = C.POINTER(C.c_float)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(_fp)


class ModelBase:
    """ModelBase(mass, dt, s_dim, a_dim) — include/model_base.hpp:58-61."""

    def __init__(self, mass=1.0, dt=0.01, s_dim=2, a_dim=1, device=-1):
        self.mass, self.dt, self.s_dim, self.a_dim, self.device = float(mass), float(dt), s_dim, a_dim, device
        self._lib = _capi.load()

    def freeStep(self, state):
        """mBuildFreeStepGraph: state [k|1, s] -> [k|1, s]."""
        st = _f32(state).reshape(-1, self.s_dim)
        out = np.empty_like(st)
        check(self._lib.mppi_model_free_step(self.device, self.mass, self.dt, self.s_dim, self.a_dim,
                                             st.shape[0], _ptr(st), _ptr(out)))
        return out

    def actionStep(self, action):
        """mBuildActionStepGraph: action [k, a] -> [k, s]."""
        ac = _f32(action).reshape(-1, self.a_dim)
        out = np.empty((ac.shape[0], self.s_dim), np.float32)
        check(self._lib.mppi_model_action_step(self.device, self.mass, self.dt, self.s_dim, self.a_dim,
                                               ac.shape[0], _ptr(ac), _ptr(out)))
        return out

    def predict(self, state, action):
        """mBuildModelStepGraph: state [k|1, s], action [k, a] -> next state [k, s]."""
        st = _f32(state).reshape(-1, self.s_dim)
        ac = _f32(action).reshape(-1, self.a_dim)
        out = np.empty((ac.shape[0], self.s_dim), np.float32)
        check(self._lib.mppi_model_step(self.device, self.mass, self.dt, self.s_dim, self.a_dim, st.shape[0],
                                        ac.shape[0], _ptr(st), _ptr(ac), _ptr(out)))
        return out


class CostBase:
    """CostBase(lambda, sigma[a,a], goal[s], Q[s]) — include/cost_base.hpp:77-80."""

    def __init__(self, lam, sigma, goal, Q, device=-1):
        self.lam = float(lam)
        self.sigma = _f32(sigma)
        self.goal = _f32(goal).ravel()
        self.Q = _f32(Q).ravel()
        self.a_dim = self.sigma.shape[0]
        self.s_dim = self.goal.size
        self.device = device
        self._lib = _capi.load()

    def setGoal(self, goal):
        goal = _f32(goal).ravel()
        if goal.size != self.s_dim:
            return False
        self.goal = goal
        return True

    def stateCost(self, state):
        st = _f32(state).reshape(-1, self.s_dim)
        out = np.empty(st.shape[0], np.float32)
        check(self._lib.mppi_cost_state(self.device, st.shape[0], self.s_dim, _ptr(st), _ptr(self.goal),
                                        _ptr(self.Q), _ptr(out)))
        return out

    finalCost = stateCost   # mBuildFinalStepCostGraph, src/cost_base.cpp:52-54

    def actionCost(self, action, noise):
        ac = _f32(action).ravel()
        nz = _f32(noise).reshape(-1, self.a_dim)
        out = np.empty(nz.shape[0], np.float32)
        check(self._lib.mppi_cost_action(self.device, nz.shape[0], self.a_dim, self.lam, _ptr(self.sigma),
                                         _ptr(ac), _ptr(nz), _ptr(out)))
        return out

    def stepCost(self, state, action, noise):
        st = _f32(state).reshape(-1, self.s_dim)
        ac = _f32(action).ravel()
        nz = _f32(noise).reshape(-1, self.a_dim)
        out = np.empty(st.shape[0], np.float32)
        check(self._lib.mppi_cost_step(self.device, st.shape[0], self.s_dim, self.a_dim, self.lam,
                                       _ptr(self.sigma), _ptr(self.goal), _ptr(self.Q), _ptr(st), _ptr(ac),
                                       _ptr(nz), _ptr(out)))
        return out


def actionCostPython(lam, gamma, upsilon, sigma, action, noise, device=-1):
    """CostBase.action_cost of the Python twin (scripts/src/costs/cost_base.py:114-170) on the GPU: noise [k, a] -> [k]."""
    lib = _capi.load()
    ac = _f32(action).ravel()
    nz = _f32(np.asarray(noise).reshape(-1, ac.size))
    out = np.empty(nz.shape[0], np.float32)
    check(lib.mppi_cost_action_py(device, nz.shape[0], ac.size, float(lam), float(gamma), float(upsilon), _ptr(_f32(sigma)),
                                  _ptr(ac), _ptr(nz), _ptr(out)))
    return out


def ellipseStateCost(state, a, b, center_x, center_y, speed, m_state, m_vel, device=-1):
    """ElipseCost.state_cost (scripts/src/costs/elipse_cost.py:46-79) on the GPU: state [k, 4] -> [k]."""
    lib = _capi.load()
    st = _f32(np.asarray(state).reshape(-1, 4))
    out = np.empty(st.shape[0], np.float32)
    check(lib.mppi_cost_state_ellipse(device, st.shape[0], _ptr(st), *[float(v) for v in (a, b, center_x, center_y, speed, m_state, m_vel)],
                                      _ptr(out)))
    return out


def quatStateCost(state, goal, q10, device=-1):
    """StaticQuatCost.state_cost on a batch: state [k, 13], goal [13], q10 [10] -> [k]."""
    st = _f32(state).reshape(-1, 13)
    out = np.empty(st.shape[0], np.float32)
    check(_capi.load().mppi_cost_state_quat(device, st.shape[0], _ptr(st), _ptr(_f32(goal).ravel()), _ptr(_f32(q10).ravel()),
                                            _ptr(out)))
    return out


def ellipse3dStateCost(state, normal, a_vec, axis, center, speed, m_state, m_vel, device=-1):
    """ElipseCost3D.state_cost on a batch: state [k, 13] -> [k]."""
    st = _f32(state).reshape(-1, 13)
    out = np.empty(st.shape[0], np.float32)
    check(_capi.load().mppi_cost_state_ellipse3d(device, st.shape[0], _ptr(st), _ptr(_f32(normal).ravel()), _ptr(_f32(a_vec).ravel()),
                                                 _ptr(_f32(axis).ravel()), _ptr(_f32(center).ravel()), float(speed), float(m_state),
                                                 float(m_vel), _ptr(out)))
    return out


def blockDiag(block, nb):
    """utile::blockDiag — src/utile.cpp:10-43."""
    lib = _capi.load()
    b = _f32(block)
    out = np.empty((b.shape[0] * nb, b.shape[1] * nb), np.float32)
    check(lib.mppi_block_diag(_ptr(b), b.shape[0], b.shape[1], nb, _ptr(out)))
    return out


class ControllerBase:
    """ControllerBase(k, tau, dt, mass, s_dim, a_dim) — include/controller_base.hpp:60-65.

    Extra keyword arguments expose what the reference hard-codes in its constructor
    (lambda=1, sigma=I, goal=(1,0,..), Q=1; src/controller_base.cpp:37-69) and the B200 additions:
    rank/world sample sharding, n_controllers batching, an external CUDA stream.

    `mass` behaves as in the reference's C++ class and in include/controller_base.hpp: it is stored and NOT forwarded
    to the model, which is built with mass 1 (src/controller_base.cpp:32,68).  `model_mass` (or setModelMass) sets the
    mass the dynamics use.
    """

    def __init__(self, k, tau, dt, mass, s_dim, a_dim, lam=1.0, sigma=None, goal=None, Q=None, seed=1,
                 device=-1, rank=0, world=1, n_controllers=1, goal_per_controller=False, stream=None, model="point_mass",
                 philox_rounds=10, model_mass=1.0):
        self._lib = _capi.load()
        self.k, self.tau, self.dt, self.mass, self.s_dim, self.a_dim = k, tau, dt, mass, s_dim, a_dim
        self.n = n_controllers
        self.device = device
        self._lam = float(lam)
        self._goal_per_controller = bool(goal_per_controller)
        cfg = MppiConfig()
        self._lib.mppi_config_default(C.byref(cfg), k, tau, dt, float(model_mass), s_dim, a_dim)
        cfg.lambda_ = lam
        keep = []
        for name, val in (("sigma", sigma), ("goal", goal), ("q", Q)):
            if val is not None:
                arr = _f32(val)
                keep.append(arr)
                setattr(cfg, name, _ptr(arr))
        cfg.seed = seed
        cfg.device = device
        cfg.rank, cfg.world = rank, world
        cfg.n_controllers = n_controllers
        cfg.goal_per_controller = 1 if goal_per_controller else 0
        cfg.stream = stream
        cfg.model = {"point_mass": _capi.MODEL_POINT_MASS, "auv": _capi.MODEL_AUV}[model]
        cfg.philox_rounds = int(philox_rounds)
        self._h = C.c_void_p()
        check(self._lib.mppi_create(C.byref(cfg), C.byref(self._h)))
        self.k_local = self._lib.mppi_k_local(self._h)
        self.k_offset = self._lib.mppi_k_offset(self._h)
        self._action = np.empty((n_controllers, a_dim), np.float32)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mppi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _shape_out(self, arr):
        return arr[0].copy() if self.n == 1 else arr.copy()

    # ---- reference API -------------------------------------------------------------------------
    def next(self, x):
        """ControllerBase::next — src/controller_base.cpp:135-153."""
        x = _f32(x).reshape(self.n, self.s_dim)
        check(self._lib.mppi_next(self._h, _ptr(x), _ptr(self._action)), self._h)
        return self._shape_out(self._action)

    def setGoal(self, goal):
        """ControllerBase::setGoal — src/controller_base.cpp:126-133 (size mismatch -> False).  [s_dim] sets the goal of
        every controller of the handle; [n, s_dim] one goal each, on handles created with goal_per_controller."""
        g = _f32(goal).ravel()
        if g.size == self.s_dim:
            rows = 1
        elif g.size == self.n * self.s_dim and self._goal_per_controller:
            rows = self.n
        else:
            return False
        check(self._lib.mppi_set_goal_n(self._h, _ptr(g), rows), self._h)
        return True

    def setModelMass(self, mass):
        check(self._lib.mppi_set_mass(self._h, float(mass)), self._h)

    # ---- parity / debug ---------------------------------------------------------------------------
    def nextWithNoise(self, x, eps):
        """Same update with eps [n, k_local, tau, a] injected (host buffer)."""
        x = _f32(x).reshape(self.n, self.s_dim)
        eps = _f32(eps)
        assert eps.size == self.n * self.k_local * self.tau * self.a_dim, "eps has the wrong size"
        check(self._lib.mppi_next_with_noise(self._h, _ptr(x), _ptr(eps), _ptr(self._action)), self._h)
        return self._shape_out(self._action)

    def nextWithNoiseDev(self, x, eps_dev_ptr):
        x = _f32(x).reshape(self.n, self.s_dim)
        check(self._lib.mppi_next_with_noise_dev(self._h, _ptr(x), C.c_void_p(eps_dev_ptr), _ptr(self._action)),
              self._h)
        return self._shape_out(self._action)

    def getCosts(self):
        out = np.empty((self.n, self.k_local), np.float32)
        check(self._lib.mppi_get_costs(self._h, _ptr(out)), self._h)
        return out[0] if self.n == 1 else out

    def getSequence(self):
        out = np.empty((self.n, self.tau, self.a_dim), np.float32)
        check(self._lib.mppi_get_sequence(self._h, _ptr(out)), self._h)
        return out[0] if self.n == 1 else out

    def setSequence(self, U):
        U = _f32(U)
        assert U.size == self.n * self.tau * self.a_dim
        check(self._lib.mppi_set_sequence(self._h, _ptr(U)), self._h)

    def getUpdate(self):
        out = np.empty((self.n, self.tau, self.a_dim), np.float32)
        check(self._lib.mppi_get_update(self._h, _ptr(out)), self._h)
        return out[0] if self.n == 1 else out

    def getWeightStats(self):
        beta = np.empty(self.n, np.float32)
        eta = np.empty(self.n, np.float32)
        check(self._lib.mppi_get_weight_stats(self._h, _ptr(beta), _ptr(eta)), self._h)
        return beta, eta

    def dumpNoise(self):
        out = np.empty((self.n, self.k_local, self.tau, self.a_dim), np.float32)
        check(self._lib.mppi_dump_noise(self._h, _ptr(out)), self._h)
        return out[0] if self.n == 1 else out

    def setLambda(self, lam):
        check(self._lib.mppi_set_lambda(self._h, float(lam)), self._h)
        self._lam = float(lam)

    def setSigma(self, sigma):
        check(self._lib.mppi_set_sigma(self._h, _ptr(_f32(sigma))), self._h)

    def setActionCost(self, form="cpp", gamma=None, upsilon=1.0):
        """Python-twin extras (scripts/src/costs/cost_base.py:114-170, controller_base.py:348-369): `form`
        "cpp" = lambda u^T S^-1 eps (src/cost_base.cpp:63-68), "python" = the gamma / upsilon form;
        upsilon also scales the sampling, eps = (upsilon sigma) z.  gamma defaults to lambda."""
        forms = {"cpp": _capi.MPPI_ACTION_COST_CPP, "python": _capi.MPPI_ACTION_COST_PYTHON}
        g = self._lam if gamma is None else float(gamma)
        check(self._lib.mppi_set_action_cost(self._h, forms[form], g, float(upsilon)), self._h)

    def setEllipseCost(self, a, b, center_x, center_y, speed, m_state, m_vel):
        """ElipseCost (scripts/src/costs/elipse_cost.py:9-79) as the state cost; point_mass2d only."""
        check(self._lib.mppi_set_ellipse_cost(self._h, *[float(v) for v in (a, b, center_x, center_y, speed, m_state, m_vel)]),
              self._h)

    def setStaticCost(self):
        check(self._lib.mppi_set_static_cost(self._h), self._h)

    def setNormalizeCost(self, on=True):
        """norm_arg of the Python twin (controller_base.py:468-474): exponent -(S - beta)/(lambda max(S - beta))."""
        check(self._lib.mppi_set_normalize_cost(self._h, int(bool(on))), self._h)

    def setActionLimits(self, act_min=None, act_max=None):
        """clip_act of the Python twin (controller_base.py:500-504): the updated sequence is clipped to [act_min, act_max]
        (scalars or [a_dim]) before the next action is taken; None, None turns the clipping off."""
        if act_min is None and act_max is None:
            check(self._lib.mppi_set_action_limits(self._h, 0, 1, None, None), self._h)
            return
        lo, hi = _f32(np.atleast_1d(act_min)).ravel(), _f32(np.atleast_1d(act_max)).ravel()
        assert lo.size == hi.size
        check(self._lib.mppi_set_action_limits(self._h, 1, lo.size, _ptr(lo), _ptr(hi)), self._h)

    def filterSequence(self, window=10, polyorder=9):
        """The Savitzky-Golay pass of the Python twin (controller_base.py:281-291): a filtered COPY of the current sequence
        (the reference stores it aside too); scipy.signal.savgol_filter semantics, mode "interp"."""
        U = _f32(self.getSequence()).reshape(self.n, self.tau, self.a_dim)
        out = np.empty_like(U)
        for c in range(self.n):
            check(self._lib.mppi_savgol_filter(self.tau, self.a_dim, _ptr(np.ascontiguousarray(U[c])), int(window), int(polyorder),
                                               _ptr(out[c])))
        return out[0] if self.n == 1 else out

    def setQ(self, q):
        check(self._lib.mppi_set_q(self._h, _ptr(_f32(q))), self._h)

    def setUpdateCounter(self, c):
        check(self._lib.mppi_set_update_counter(self._h, int(c)), self._h)

    # ---- learned MLP dynamics (learning_base, row A13) ----------------------------------------------
    def setMlp(self, mlp):
        """mlp: dict W1 [s+a,H], b1 [H], W2 [H,H], b2 [H], W3 [H,s], b3 [s] (Keras layout) and optional
        Xmean/Xstd [s+a], Ymean/Ystd [s].  Switches the rollout to the tensor-core MLP model."""
        keep = {k: _f32(v) for k, v in mlp.items()}
        H = keep["b1"].size
        self._mlp_hidden = H
        opt = lambda k: _ptr(keep[k]) if k in keep else None
        check(self._lib.mppi_set_mlp(self._h, H, _ptr(keep["W1"]), _ptr(keep["b1"]), _ptr(keep["W2"]), _ptr(keep["b2"]),
                                     _ptr(keep["W3"]), _ptr(keep["b3"]), opt("Xmean"), opt("Xstd"), opt("Ymean"),
                                     opt("Ystd")), self._h)

    def mlpTrainStep(self, state, action, next_state, learning_rate):
        """One full-batch Adam step on the normalised MSE (learner_base.py:469-496); returns the loss before the step."""
        st, ac, nx = _f32(state).reshape(-1, self.s_dim), _f32(action).reshape(-1, self.a_dim), _f32(next_state).reshape(-1, self.s_dim)
        assert st.shape[0] == ac.shape[0] == nx.shape[0]
        loss = C.c_float(0)
        check(self._lib.mppi_mlp_train_step(self._h, st.shape[0], _ptr(st), _ptr(ac), _ptr(nx), float(learning_rate),
                                            C.byref(loss)), self._h)
        return loss.value

    def mlpTrain(self, state, action, next_state, epochs=1, learning_rate=0.1, batch_size=-1, augment_samples=0, augment_sigma=0.001,
                 seed=1):
        """LearnerBase.train (learner_base.py:324-358): epochs x { augment_data, Adam step(s) on the normalised MSE }.
        batch_size = -1 is the reference's one full-batch step per epoch.  Returns the loss before every step."""
        st, ac, nx = _f32(state).reshape(-1, self.s_dim), _f32(action).reshape(-1, self.a_dim), _f32(next_state).reshape(-1, self.s_dim)
        assert st.shape[0] == ac.shape[0] == nx.shape[0]
        n_ep = st.shape[0] * (augment_samples if augment_samples > 0 else 1)
        bs = batch_size if 0 < batch_size < n_ep else n_ep
        losses = np.empty(epochs * ((n_ep + bs - 1) // bs), np.float32)
        nsteps = C.c_int(0)
        check(self._lib.mppi_mlp_train(self._h, st.shape[0], _ptr(st), _ptr(ac), _ptr(nx), int(epochs), int(batch_size), float(learning_rate),
                                       int(augment_samples), float(augment_sigma), int(seed), _ptr(losses), C.byref(nsteps)), self._h)
        return losses[:nsteps.value]

    def mlpSetAdam(self, beta1=0.9, beta2=0.999, epsilon=1e-7):
        check(self._lib.mppi_mlp_set_adam(self._h, float(beta1), float(beta2), float(epsilon)), self._h)

    def mlpGetWeights(self, hidden=None):
        s, a, H = self.s_dim, self.a_dim, (hidden or getattr(self, "_mlp_hidden", 128))
        out = dict(W1=np.empty((s + a, H), np.float32), b1=np.empty(H, np.float32), W2=np.empty((H, H), np.float32),
                   b2=np.empty(H, np.float32), W3=np.empty((H, s), np.float32), b3=np.empty(s, np.float32))
        check(self._lib.mppi_mlp_get_weights(self._h, *[_ptr(out[k]) for k in ("W1", "b1", "W2", "b2", "W3", "b3")]), self._h)
        return out

    def mlpPredict(self, state, action):
        st = _f32(state).reshape(-1, self.s_dim)
        ac = _f32(action).reshape(-1, self.a_dim)
        out = np.empty((ac.shape[0], self.s_dim), np.float32)
        check(self._lib.mppi_mlp_predict(self._h, st.shape[0], ac.shape[0], _ptr(st), _ptr(ac), _ptr(out)), self._h)
        return out

    # ---- AUV (Fossen) dynamics and StaticQuatCost (rows N3 / N4) -------------------------------------
    def setAuvModel(self, parameters, rk=None):
        """AUVModel parameters as the reference's dict (scripts/src/models/auv_model.py:146-255): mass, volume,
        density, cog, cob, Ma, inertial{ixx..iyz}, linear_damping, quad_damping, linear_damping_forward_speed
        (damping lists of 6 are diagonals), rk.  Needs a controller created with model="auv"."""
        sq = lambda v: (np.diag(np.asarray(v, np.float64)) if np.ndim(v) == 1 else np.asarray(v, np.float64)).ravel()
        p = _capi.MppiAuvParams()
        p.mass, p.volume, p.density = float(parameters["mass"]), float(parameters["volume"]), float(parameters["density"])
        p.cog[:] = [float(v) for v in parameters["cog"]]
        p.cob[:] = [float(v) for v in parameters["cob"]]
        p.added_mass[:] = [float(v) for v in np.asarray(parameters["Ma"], np.float64).ravel()]
        i = parameters["inertial"]
        p.inertia[:] = [float(i[k]) for k in ("ixx", "iyy", "izz", "ixy", "ixz", "iyz")]
        p.linear_damping[:] = [float(v) for v in sq(parameters["linear_damping"])]
        p.quad_damping[:] = [float(v) for v in parameters["quad_damping"]]
        p.linear_damping_forward_speed[:] = [float(v) for v in sq(parameters["linear_damping_forward_speed"])]
        p.rk = int(parameters.get("rk", 1) if rk is None else rk)
        check(self._lib.mppi_set_auv_model(self._h, C.byref(p)), self._h)

    def setNnAuvModel(self, nn):
        """The reference's learned AUV model (NNAUVModel, scripts/src/models/nn_model.py:181-304) as the dynamics of an "auv"
        controller: nn = dict(W=[W0, .., W_last], b=[b0, ..], Xmean, Xstd, Ymean, Ystd), Keras-layout weights, W0 [16, H],
        hidden layers [H, H], W_last [H, 13]."""
        Ws = [_f32(w) for w in nn["W"]]
        bs = [_f32(v).ravel() for v in nn["b"]]
        n_hidden, H = len(Ws) - 1, Ws[0].shape[1]
        assert Ws[0].shape[0] == 16 and Ws[-1].shape == (H, 13) and all(w.shape == (H, H) for w in Ws[1:-1])
        Wp = (_fp * len(Ws))(*[_ptr(w) for w in Ws])
        bp = (_fp * len(bs))(*[_ptr(v) for v in bs])
        opt = {k: _f32(nn[k]).ravel() for k in ("Xmean", "Xstd", "Ymean", "Ystd") if k in nn and nn[k] is not None}
        g = lambda k: _ptr(opt[k]) if k in opt else None
        check(self._lib.mppi_set_nn_auv_model(self._h, n_hidden, H, Wp, bp, g("Xmean"), g("Xstd"), g("Ymean"), g("Ystd")), self._h)

    def auvPredict(self, state, action):
        """AUVModel.build_step_graph (auv_model.py:285-306): state [k|1, 13], action [k, 6] -> [k, 13]."""
        st = _f32(state).reshape(-1, 13)
        ac = _f32(action).reshape(-1, 6)
        out = np.empty((ac.shape[0], 13), np.float32)
        check(self._lib.mppi_auv_predict(self._h, st.shape[0], ac.shape[0], _ptr(st), _ptr(ac), _ptr(out)), self._h)
        return out

    def setQuatCost(self, q10):
        """StaticQuatCost (scripts/src/costs/static_cost.py:73-159) with Q = diag(q10) as the state cost."""
        q = _f32(q10).ravel()
        assert q.size == 10
        check(self._lib.mppi_set_quat_cost(self._h, _ptr(q)), self._h)

    def setEllipse3dCost(self, normal, a_vec, axis, center, speed, m_state, m_vel):
        """ElipseCost3D (scripts/src/costs/elipse_cost.py:99-246) as the state cost of an AUV controller."""
        check(self._lib.mppi_set_ellipse3d_cost(self._h, _ptr(_f32(normal).ravel()), _ptr(_f32(a_vec).ravel()), _ptr(_f32(axis).ravel()),
                                                _ptr(_f32(center).ravel()), float(speed), float(m_state), float(m_vel)), self._h)

    # ---- developer knobs ------------------------------------------------------------------------------
    def debugTrace(self, on=True):
        check(self._lib.mppi_debug_trace(self._h, int(bool(on))), self._h)

    def getTrace(self):
        """[n, grid_x, 12] uint64 nanosecond stamps of the last update's phases (see include/mppi_b200.h)."""
        gx = self._lib.mppi_last_grid_x(self._h)
        out = np.zeros((self.n * gx, 12), np.uint64)
        check(self._lib.mppi_debug_get_trace(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), self.n * gx), self._h)
        return out.reshape(self.n, gx, 12)

    # ---- asynchronous halves (bench / multi-rank) ---------------------------------------------------
    def setState(self, x):
        x = _f32(x).reshape(self.n, self.s_dim)
        check(self._lib.mppi_set_state(self._h, _ptr(x)), self._h)

    def enqueueUpdate(self, eps_dev_ptr=None):
        check(self._lib.mppi_enqueue_update(self._h, C.c_void_p(eps_dev_ptr) if eps_dev_ptr else None), self._h)

    def enqueueExchange(self):
        check(self._lib.mppi_enqueue_exchange(self._h), self._h)

    def enqueueFinish(self):
        check(self._lib.mppi_enqueue_finish(self._h), self._h)

    def fetchAction(self):
        check(self._lib.mppi_fetch_action(self._h, _ptr(self._action)), self._h)
        return self._shape_out(self._action)

    def synchronize(self):
        check(self._lib.mppi_synchronize(self._h), self._h)

    def exchangeStride(self):
        return self._lib.mppi_exchange_stride(self._h)

    def exchangeBuffers(self):
        s, r = C.c_void_p(), C.c_void_p()
        check(self._lib.mppi_exchange_buffers(self._h, C.byref(s), C.byref(r)), self._h)
        return s.value, r.value

    def setExchangeBuffers(self, send_ptr, recv_ptr):
        check(self._lib.mppi_exchange_set_buffers(self._h, C.c_void_p(send_ptr), C.c_void_p(recv_ptr)), self._h)

    def peerHandle(self):
        """64-byte CUDA IPC handle of this rank's mailbox (fused exchange over peer memory)."""
        buf = C.create_string_buffer(64)
        check(self._lib.mppi_peer_handle(self._h, buf), self._h)
        return buf.raw

    def peerAttach(self, handles):
        """handles: the world 64-byte handles in rank order (all-gathered by the caller)."""
        blob = b"".join(handles)
        check(self._lib.mppi_peer_attach(self._h, C.create_string_buffer(blob, len(blob))), self._h)

    def commInit(self, unique_id_bytes):
        buf = C.create_string_buffer(bytes(unique_id_bytes), 128)
        check(self._lib.mppi_comm_init(self._h, C.cast(buf, C.c_void_p)), self._h)

    # ---- stage entry points (the reference's m* graph builders) ---------------------------------------
    def prepareAction(self, actions, t):
        U = _f32(actions).reshape(-1, self.a_dim)
        out = np.empty(self.a_dim, np.float32)
        check(self._lib.mppi_prepare_action(U.shape[0], self.a_dim, _ptr(U), t, _ptr(out)))
        return out

    def prepareNoise(self, noises, t):
        nz = _f32(noises)
        k, T, a = nz.shape
        out = np.empty((k, a), np.float32)
        check(self._lib.mppi_prepare_noise(self.device, k, T, a, _ptr(nz), t, _ptr(out)))
        return out

    def updateStages(self, cost, noises, lam=1.0):
        """mBeta/mExpArg/mExp/mNabla/mWeights/mWeightedNoise on given costs [k] and noises [k,T,a]."""
        c = _f32(cost).ravel()
        nz = _f32(noises)
        k, T, a = nz.shape
        beta, nabla = C.c_float(), C.c_float()
        arg, e, w = (np.empty(k, np.float32) for _ in range(3))
        wn = np.empty((T, a), np.float32)
        check(self._lib.mppi_update_stages(self.device, k, T, a, float(lam), _ptr(c), _ptr(nz), C.byref(beta),
                                           _ptr(arg), _ptr(e), C.byref(nabla), _ptr(w), _ptr(wn)))
        return dict(beta=beta.value, exp_arg=arg, exp=e, nabla=nabla.value, weights=w, weighted_noise=wn)

    def getNew(self, current, nb):
        cur = _f32(current).reshape(-1, self.a_dim)
        out = np.empty((nb, self.a_dim), np.float32)
        check(self._lib.mppi_get_new(cur.shape[0], self.a_dim, _ptr(cur), nb, _ptr(out) if nb else None))
        return out

    def shift(self, current, init, nb):
        cur = _f32(current).reshape(-1, self.a_dim)
        ini = _f32(init).reshape(-1, self.a_dim)
        out = np.empty_like(cur)
        check(self._lib.mppi_shift(cur.shape[0], self.a_dim, _ptr(cur), _ptr(ini), nb, _ptr(out)))
        return out


def comm_unique_id():
    lib = _capi.load()
    buf = C.create_string_buffer(128)
    check(lib.mppi_comm_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf.raw)


def philox_raw(seed, call0, sample, update, stream, n_calls, device=-1, rounds=10):
    lib = _capi.load()
    out = np.empty((n_calls, 4), np.uint32)
    check(lib.mppi_philox_raw_rounds(device, seed, call0, sample, update, stream, n_calls, int(rounds),
                                     out.ctypes.data_as(C.POINTER(C.c_uint32))))
    return out
