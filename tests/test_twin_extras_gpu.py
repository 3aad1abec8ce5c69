"""Python-twin extras on the fused kernels: upsilon noise scaling, the
gamma / upsilon action cost (scripts/src/costs/cost_base.py:114-170) and cost normalisation
(scripts/src/controllers/controller_base.py:468-474), against the fp64 C restatement that
tests/test_python_twin_fixtures.py pins on the reference's own Python code."""
import numpy as np
import pytest

from tests.util import assert_update_close, controller_from_cfg, make_cfg, rel_err

pytestmark = pytest.mark.gpu


def _full_sigma(a, seed):
    rng = np.random.default_rng(seed)
    L = 0.2 * rng.standard_normal((a, a))
    return (L @ L.T + 0.2 * np.eye(a)).astype(np.float32)


def _inputs(cfg, seed, upsilon):
    rng = np.random.default_rng(seed)
    k, tau, s, a = cfg["k"], cfg["tau"], cfg["s_dim"], cfg["a_dim"]
    x0 = rng.uniform(-1, 1, s).astype(np.float32)
    U0 = (0.2 * rng.standard_normal((tau, a))).astype(np.float32)
    z = rng.standard_normal((k, tau, a)).astype(np.float32)
    eps = np.einsum("ij,ktj->kti", upsilon * cfg["sigma"].astype(np.float64), z).astype(np.float32)
    return x0, U0, eps


CASES = [
    # k, tau, a, full sigma, gamma, upsilon, normalize
    (1000, 20, 1, False, 0.4, 1.7, False),
    (4096, 50, 2, True, 0.7, 1.0, True),
    (3000, 33, 3, True, 0.9, 2.5, True),        # T*a % 4 != 0: generic (non-TMA) injected path
    (8192, 100, 3, False, 1.5, 0.6, False),     # group-cooperative injected kernel (3 warps per tile)
    (2048, 12, 4, True, 2.0, 1.3, True),
]


@pytest.mark.parametrize("k,tau,a,full,gamma,upsilon,normalize", CASES)
def test_injected_noise_matches_twin_oracle(oracle64, oracle32, k, tau, a, full, gamma, upsilon, normalize):
    s = 2 * a
    cfg = make_cfg(k, tau, s, a, lam=1.3, sigma=_full_sigma(a, a) if full else None, mass=1.5,
                   goal=np.random.default_rng(a).uniform(-1, 1, s), q=1 + np.random.default_rng(a + 1).random(s))
    x0, U0, eps = _inputs(cfg, k, upsilon)
    c = controller_from_cfg(cfg)
    try:
        c.setActionCost("python", gamma=gamma, upsilon=upsilon)
        c.setNormalizeCost(normalize)
        c.setSequence(U0)
        act = c.nextWithNoise(x0, eps)
        U_new, U_shift, costs = c.getUpdate(), c.getSequence(), c.getCosts()
    finally:
        c.close()
    kw = dict(gamma=gamma, upsilon=upsilon, normalize=normalize)
    ref = oracle64.mppi_update_py(cfg, x0, U0, eps, **kw)
    ref32 = oracle32.mppi_update_py(cfg, x0, U0, eps, **kw)
    assert rel_err(costs, ref["costs"]) < 1e-5
    assert_update_close(U_new, ref["U_new"], ref32["U_new"], what="U_new")
    assert_update_close(U_shift, ref["U_shift"], ref32["U_shift"], what="U_shift")
    assert np.abs(act - ref["next"]).max() <= 1e-5 * np.abs(ref["U_new"]).max()
    np.testing.assert_array_equal(U_shift[:-1], U_new[1:])


@pytest.mark.parametrize("k,tau,a,full,gamma,upsilon,normalize", CASES)
def test_philox_store_then_replay_matches_twin_oracle(oracle64, oracle32, k, tau, a, full, gamma, upsilon, normalize):
    """Philox mode generates eps = (upsilon sigma) z itself; dump that tensor and replay it through the oracle."""
    s = 2 * a
    cfg = make_cfg(k, tau, s, a, lam=0.8, sigma=_full_sigma(a, 10 + a) if full else None)
    x0, U0, _ = _inputs(cfg, 7 * k, upsilon)
    c = controller_from_cfg(cfg, seed=11)
    try:
        c.setActionCost("python", gamma=gamma, upsilon=upsilon)
        c.setNormalizeCost(normalize)
        c.setSequence(U0)
        act = c.next(x0)
        U_new, costs = c.getUpdate(), c.getCosts()
        eps = c.dumpNoise()
    finally:
        c.close()
    # the dumped tensor carries the upsilon scaling: its per-axis spread follows upsilon * sigma
    want_std = upsilon * np.sqrt(np.diag(cfg["sigma"].astype(np.float64) @ cfg["sigma"].astype(np.float64).T))
    np.testing.assert_allclose(eps.reshape(-1, a).std(0), want_std, rtol=0.05)
    kw = dict(gamma=gamma, upsilon=upsilon, normalize=normalize)
    ref = oracle64.mppi_update_py(cfg, x0, U0, eps, **kw)
    ref32 = oracle32.mppi_update_py(cfg, x0, U0, eps, **kw)
    assert rel_err(costs, ref["costs"]) < 1e-5
    assert_update_close(U_new, ref["U_new"], ref32["U_new"], what="U_new")
    assert np.abs(act - ref["next"]).max() <= 1e-5 * np.abs(ref["U_new"]).max()


def test_upsilon_with_cpp_action_cost(oracle64):
    """upsilon only scales the sampling when the C++ action cost is kept: same as sigma' = upsilon sigma for the
    noise, with Sigma^-1 of the unscaled sigma in the cost."""
    k, tau, a, ups = 2048, 16, 2, 1.9
    cfg = make_cfg(k, tau, 2 * a, a, lam=1.1, sigma=_full_sigma(a, 3))
    x0, U0, eps = _inputs(cfg, 5, ups)
    c = controller_from_cfg(cfg)
    try:
        c.setActionCost("cpp", upsilon=ups)
        c.setSequence(U0)
        c.nextWithNoise(x0, eps)
        U_inj, costs_inj = c.getUpdate(), c.getCosts()
        c.setSequence(U0)
        c.next(x0)
        U_phx, costs_phx, eps_phx = c.getUpdate(), c.getCosts(), c.dumpNoise()
    finally:
        c.close()
    ref = oracle64.mppi_update(cfg, x0, U0, eps)
    assert rel_err(costs_inj, ref["costs"]) < 1e-5 and rel_err(U_inj, ref["U_new"]) < 1e-5
    ref = oracle64.mppi_update(cfg, x0, U0, eps_phx)
    assert rel_err(costs_phx, ref["costs"]) < 1e-5 and rel_err(U_phx, ref["U_new"]) < 2e-5


def test_normalised_update_batched_controllers(oracle64):
    """n_controllers > 1 with per-controller goals: every controller normalises with its own cost range."""
    n, k, tau, a = 5, 512, 10, 2
    s = 2 * a
    rng = np.random.default_rng(3)
    goals = rng.uniform(-1, 1, (n, s)).astype(np.float32)
    cfg = make_cfg(k, tau, s, a, lam=0.9)
    xs = rng.uniform(-1, 1, (n, s)).astype(np.float32)
    Us = (0.2 * rng.standard_normal((n, tau, a))).astype(np.float32)
    eps = (0.25 * rng.standard_normal((n, k, tau, a))).astype(np.float32)
    from mppi_tf_b200 import ControllerBase
    c = ControllerBase(k, tau, cfg["dt"], cfg["mass"], s, a, lam=cfg["lambda"], sigma=cfg["sigma"], goal=goals,
                       Q=cfg["q"], n_controllers=n, goal_per_controller=True)
    try:
        c.setActionCost("python", gamma=0.5, upsilon=1.0)
        c.setNormalizeCost(True)
        c.setSequence(Us)
        c.nextWithNoise(xs, eps)
        U_new, costs = c.getUpdate(), c.getCosts()
    finally:
        c.close()
    for i in range(n):
        ci = dict(cfg, goal=goals[i])
        ref = oracle64.mppi_update_py(ci, xs[i], Us[i], eps[i], gamma=0.5, upsilon=1.0, normalize=True)
        assert rel_err(costs[i], ref["costs"]) < 1e-5
        assert rel_err(U_new[i], ref["U_new"]) < 1e-5


def test_normalise_rejected_when_sharded_without_peer_exchange():
    """With world > 1 the cost range travels through the fused peer-memory mailboxes (tests/peer_check_worker.py
    covers that on two GPUs); without them it is refused."""
    from mppi_tf_b200 import ControllerBase, MppiError, _capi
    c = ControllerBase(1024, 8, 0.1, 1.0, 2, 1, rank=0, world=2)
    try:
        with pytest.raises(MppiError) as e:
            c.setNormalizeCost(True)
        assert e.value.code == _capi.MPPI_ERR_UNSUPPORTED
    finally:
        c.close()


def test_mlp_dynamics_with_twin_cost(oracle64):
    """The learned-MLP rollout takes the same cost options (bf16 path: 2e-2 bar).  The oracle here is the
    MLP rollout with the C++ cost; gamma = lambda, upsilon = 1 differs from it by the per-update constant
    0.5 lambda sum_t U_t^T S^-1 U_t, which is checked explicitly, and normalisation is checked through U'."""
    from tests.test_mlp_gpu import glorot_mlp
    k, tau, a = 2048, 20, 3
    s = 2 * a
    mlp = glorot_mlp(s, a, scale=0.5, bias=True)
    cfg = make_cfg(k, tau, s, a, lam=2.0)
    x0, U0, eps = _inputs(cfg, 9, 1.0)
    c = controller_from_cfg(cfg)
    try:
        c.setMlp(mlp)
        c.setSequence(U0)
        c.nextWithNoise(x0, eps)
        costs_cpp, U_cpp = c.getCosts(), c.getUpdate()
        c.setActionCost("python", gamma=cfg["lambda"], upsilon=1.0)
        c.setSequence(U0)
        c.nextWithNoise(x0, eps)
        costs_py, U_py = c.getCosts(), c.getUpdate()
    finally:
        c.close()
    const = 0.5 * cfg["lambda"] * np.einsum("ti,ij,tj->", U0, np.linalg.inv(cfg["sigma"]), U0)
    np.testing.assert_allclose(costs_py - costs_cpp, const, rtol=1e-3, atol=1e-3 * np.abs(costs_cpp).max())
    assert rel_err(U_py, U_cpp) < 1e-4
    ref = oracle64.mppi_update_mlp(cfg, mlp, x0, U0, eps)
    assert rel_err(U_cpp, ref["U_new"]) < 2e-2


@pytest.mark.parametrize("k,tau,normalize,upsilon", [(4096, 50, False, 1.0), (3000, 27, True, 1.6), (8192, 100, False, 0.8)])
def test_ellipse_cost_update(oracle64, oracle32, k, tau, normalize, upsilon):
    """ElipseCost (elipse_cost.py:46-79) as the rollout / terminal state cost, both noise modes."""
    a, s = 2, 4
    ell = (1.5, 0.8, 0.2, -0.1, 0.7, 2.0, 0.5)
    cfg = make_cfg(k, tau, s, a, lam=0.9, sigma=_full_sigma(a, 21))
    x0, U0, eps = _inputs(cfg, 31, upsilon)
    c = controller_from_cfg(cfg, seed=3)
    try:
        c.setActionCost("python", gamma=0.4, upsilon=upsilon)
        c.setNormalizeCost(normalize)
        c.setEllipseCost(*ell)
        c.setSequence(U0)
        act = c.nextWithNoise(x0, eps)
        U_inj, costs_inj = c.getUpdate(), c.getCosts()
        c.setSequence(U0)
        c.next(x0)
        U_phx, costs_phx, eps_phx = c.getUpdate(), c.getCosts(), c.dumpNoise()
        c.setStaticCost()
        c.setSequence(U0)
        c.nextWithNoise(x0, eps)
        costs_static = c.getCosts()
    finally:
        c.close()
    kw = dict(gamma=0.4, upsilon=upsilon, normalize=normalize, ellipse=ell)
    ref, ref32 = oracle64.mppi_update_py(cfg, x0, U0, eps, **kw), oracle32.mppi_update_py(cfg, x0, U0, eps, **kw)
    assert rel_err(costs_inj, ref["costs"]) < 1e-5
    assert_update_close(U_inj, ref["U_new"], ref32["U_new"], what="U_new injected")
    assert np.abs(act - ref["next"]).max() <= 1e-5 * np.abs(ref["U_new"]).max()
    ref, ref32 = oracle64.mppi_update_py(cfg, x0, U0, eps_phx, **kw), oracle32.mppi_update_py(cfg, x0, U0, eps_phx, **kw)
    assert rel_err(costs_phx, ref["costs"]) < 1e-5
    assert_update_close(U_phx, ref["U_new"], ref32["U_new"], what="U_new philox")
    kw["ellipse"] = None                       # and back to the quadratic cost
    assert rel_err(costs_static, oracle64.mppi_update_py(cfg, x0, U0, eps, **kw)["costs"]) < 1e-5


def test_ellipse_cost_needs_point_mass2d():
    from mppi_tf_b200 import ControllerBase, MppiError, _capi
    c = ControllerBase(256, 8, 0.1, 1.0, 6, 3)
    try:
        with pytest.raises(MppiError) as e:
            c.setEllipseCost(1, 1, 0, 0, 1, 1, 1)
        assert e.value.code == _capi.MPPI_ERR_UNSUPPORTED
    finally:
        c.close()
