"""Pins the CPU oracle against every known-answer vector the reference's tests hold for the
MPPI update path (tests/golden/kats.py cites each one), and cross-checks the C oracle against
the independently written graph-faithful torch restatement.  CPU only."""
import numpy as np
import pytest

from tests.golden import kats

FLOAT_EQ = dict(rtol=4 * np.finfo(np.float32).eps, atol=1e-30)   # EXPECT_FLOAT_EQ = 4 ULP


# ---- utile::blockDiag ---------------------------------------------------------------------
@pytest.mark.parametrize("nb", [1, 2, 3, 4])
def test_block_diag(oracle32, nb):
    dt, m = np.float32(kats.UTILE_DT), np.float32(kats.UTILE_M)
    A_blk = np.array([[1, dt], [0, 1]], np.float32)
    B_blk = np.array([[dt * dt / (np.float32(2) * m)], [dt / m]], np.float32)
    exp_a, exp_b = kats.blockdiag_expected(nb)
    assert oracle32.block_diag(A_blk, nb).shape == (2 * nb, 2 * nb)
    assert oracle32.block_diag(B_blk, nb).shape == (2 * nb, nb)
    np.testing.assert_allclose(oracle32.block_diag(A_blk, nb), exp_a, **FLOAT_EQ)
    np.testing.assert_allclose(oracle32.block_diag(B_blk, nb), exp_b, **FLOAT_EQ)
    if nb == 2:
        np.testing.assert_array_equal(exp_a, kats.BLOCKDIAG2_A_LITERAL(dt))


# ---- ModelBase -----------------------------------------------------------------------------
@pytest.mark.parametrize("case", kats.MODEL_CASES, ids=lambda c: c["name"])
def test_model_steps(oracle32, case):
    exp_s, exp_u, exp_res = kats.model_expected(case)
    s, a, m, dt = case["s"], case["a"], case["m"], case["dt"]
    free = oracle32.model_free_step(m, dt, s, a, case["state"])
    act = oracle32.model_action_step(m, dt, s, a, case["action"])
    full = oracle32.model_step(m, dt, s, a, case["state"], case["action"])
    assert free.shape == (len(case["state"]), s)          # InitTest: [1,s,1] stays [1,s,1]
    assert act.shape == (case["k"], s) and full.shape == (case["k"], s)
    np.testing.assert_allclose(free, exp_s, **FLOAT_EQ)
    np.testing.assert_allclose(act, exp_u, **FLOAT_EQ)
    np.testing.assert_allclose(full, exp_res, **FLOAT_EQ)


def test_model_large_literal(oracle32):
    case = kats.MODEL_CASES[2]
    lit = kats.large_testing_literal()
    for got, want in zip(kats.model_expected(case), lit):
        np.testing.assert_allclose(got, want, **FLOAT_EQ)


def test_model_py_three_steps(oracle64):
    c = kats.py_step3_expected()
    x = c["state"]
    for _ in range(3):
        x = oracle64.model_step(c["m"], c["dt"], 6, 3, x, c["action"])
    np.testing.assert_allclose(x, c["expected"], rtol=1e-6, atol=1e-6)   # assertAllClose default


# ---- CostBase ------------------------------------------------------------------------------
@pytest.mark.parametrize("case", kats.COST_CASES, ids=lambda c: c["name"])
def test_cost(oracle32, case):
    st = oracle32.cost_state(case["state"], case["goal"], case["q"])
    np.testing.assert_allclose(st, np.array(case["exp_state"], np.float32), **FLOAT_EQ)
    step = oracle32.cost_step(case["lam"], case["sigma"], case["goal"], case["q"], case["state"],
                              case["action"], case["noise"])
    np.testing.assert_allclose(step, np.array(case["exp_step"], np.float32), **FLOAT_EQ)
    assert st.shape == (case["k"],) and step.shape == (case["k"],)


def test_cost_action_general_sigma(oracle64):
    """lambda u^T Sigma^-1 eps with a non-diagonal Sigma, against numpy."""
    rng = np.random.default_rng(3)
    a = 3
    L = rng.standard_normal((a, a))
    sigma = L @ L.T + a * np.eye(a)
    u = rng.standard_normal(a)
    eps = rng.standard_normal((7, a))
    want = 0.7 * eps @ np.linalg.inv(sigma).T @ u
    np.testing.assert_allclose(oracle64.cost_action(0.7, sigma, u, eps), want, rtol=1e-12)


# ---- ControllerBase stages -------------------------------------------------------------------
@pytest.mark.parametrize("case", kats.PY_ACTION_COST_CASES, ids=lambda c: c["name"])
def test_python_action_cost(oracle64, case):
    """The gamma / upsilon action cost against the reference's own known answers (scripts/test.py:685-838)."""
    a = len(case["action"])
    got = oracle64.cost_action_py(case["lam"], case["gamma"], case["upsilon"], np.eye(a), case["action"], case["noise"])
    np.testing.assert_allclose(got, case["expected"], rtol=1e-6, atol=1e-6)


def test_python_static_cost_s13(oracle64):
    """scripts/test.py:944-1095: 13-dimensional state, a = 6, diagonal Q; step cost = state + action cost."""
    c = kats.PY_STATIC13
    sc = oracle64.cost_state(c["state"], c["goal"], c["q"])
    ac = oracle64.cost_action_py(c["lam"], c["gamma"], c["upsilon"], np.eye(6), c["action"], c["noise"])
    np.testing.assert_allclose(sc, c["expected_state"], rtol=1e-6)
    np.testing.assert_allclose(ac, c["expected_action"], rtol=1e-6)


@pytest.mark.parametrize("case", kats.ELLIPSE_CASES, ids=lambda c: c["name"])
def test_ellipse_cost(oracle64, oracle32, case):
    """scripts/test.py:1098-1161 (assertAllClose: rtol 1e-6)."""
    for orc in (oracle64, oracle32):
        got = orc.cost_state_ellipse(case["state"], kats.ELLIPSE_PARAMS)
        np.testing.assert_allclose(got, case["expected"], rtol=1e-6, atol=1e-6)


def test_data_prep(oracle32):
    for t in range(3):
        np.testing.assert_allclose(oracle32.prepare_action(kats.CTRL["action"], t),
                                   np.array(kats.CTRL_PREP["a"][t], np.float32), **FLOAT_EQ)
        got = oracle32.prepare_noise(kats.CTRL["noise"], t)
        assert got.shape == (5, 2)
        np.testing.assert_allclose(got, kats.CTRL_PREP["n"][t].astype(np.float32), **FLOAT_EQ)


def test_update_stages(oracle32):
    r = oracle32.update_stages(kats.CTRL["lam"], kats.CTRL["cost"], kats.CTRL["noise"])
    e = kats.CTRL_UPDATE
    f = lambda v: np.asarray(v, np.float32)
    np.testing.assert_allclose(r["beta"], f(e["beta"]), **FLOAT_EQ)
    np.testing.assert_allclose(r["exp_arg"], f(e["exp_arg"]), **FLOAT_EQ)
    np.testing.assert_allclose(r["exp"], f(e["exp"]), **FLOAT_EQ)
    np.testing.assert_allclose(r["nabla"], f(e["nabla"]), **FLOAT_EQ)
    np.testing.assert_allclose(r["weights"], f(e["weights"]), **FLOAT_EQ)
    np.testing.assert_allclose(r["weighted_noise"], f(e["weighted_noise"]), **FLOAT_EQ)
    np.testing.assert_allclose(r["weights"].sum(dtype=np.float32), f(e["sum_w"]), **FLOAT_EQ)


def test_get_new(oracle32):
    for nb, want in kats.CTRL_NEW.items():
        got = oracle32.get_new(kats.CTRL["action"], nb)
        assert got.shape == (nb, 2)
        np.testing.assert_allclose(got, np.asarray(want, np.float32).reshape(nb, 2), **FLOAT_EQ)


def test_shift(oracle32):
    for c in kats.CTRL_SHIFT:
        got = oracle32.shift(kats.CTRL["action"], c["init"], c["nb"])
        np.testing.assert_allclose(got, c["expected"].astype(np.float32), **FLOAT_EQ)


# ---- composition: C oracle vs the graph-faithful torch restatement ------------------------------
def _cfg(k, T, s, a, lam=1.0, sig=0.25, mass=1.0, dt=0.1, seed=0):
    rng = np.random.default_rng(seed)
    sigma = sig * np.eye(a) + 0.05 * rng.standard_normal((a, a))
    return dict(k=k, tau=T, s_dim=s, a_dim=a, dt=dt, mass=mass, **{"lambda": lam},
                sigma=sigma.astype(np.float32), goal=np.tile([1.0, 0.0], a).astype(np.float32),
                q=(1 + rng.random(s)).astype(np.float32))


@pytest.mark.parametrize("k,T,s,a", [(64, 5, 2, 1), (257, 11, 4, 2), (1024, 20, 2, 1), (300, 17, 6, 3)])
def test_full_update_c_vs_graph(oracle32, oracle64, k, T, s, a):
    import torch
    from oracle.graph_oracle import GraphOracle
    cfg = _cfg(k, T, s, a, seed=k)
    rng = np.random.default_rng(1234)
    z = rng.standard_normal((k, T, a)).astype(np.float32)
    eps = oracle32.scale_noise(cfg["sigma"], z)
    x0 = rng.uniform(-1, 1, s).astype(np.float32)
    U = (0.1 * rng.standard_normal((T, a))).astype(np.float32)
    c32 = oracle32.mppi_update(cfg, x0, U, eps)
    c64 = oracle64.mppi_update(cfg, x0, U, eps)
    g = GraphOracle(k, T, cfg["dt"], cfg["mass"], s, a, cfg["lambda"], cfg["sigma"], cfg["goal"],
                    cfg["q"], dtype=torch.float64).next(x0, U, eps)
    for key in ("costs", "U_new", "next", "U_shift"):
        np.testing.assert_allclose(c64[key], g[key], rtol=1e-10, atol=1e-12, err_msg=key)
        # fp32 op-for-op vs fp64: norm-wise, the reference's own fp32 rounding level
        scale = np.abs(c64[key]).max()
        assert np.abs(c32[key] - c64[key]).max() <= 3e-5 * scale, key
    # shift/next structure
    np.testing.assert_array_equal(c32["next"], c32["U_new"][0])
    np.testing.assert_array_equal(c32["U_shift"][:-1], c32["U_new"][1:])
    np.testing.assert_array_equal(c32["U_shift"][-1], 0)


def test_terminal_cost_counted_twice(oracle64):
    """src/controller_base.cpp:264-272: q(x_T) enters once as step T-1's state cost and once more
    as the terminal cost."""
    cfg = _cfg(4, 1, 2, 1)
    cfg["sigma"] = np.eye(1, dtype=np.float32)
    x0 = np.array([0.3, -0.2])
    U = np.array([[0.5]])
    eps = np.array([[[0.1]], [[-0.4]], [[0.0]], [[1.0]]])
    got = oracle64.rollout_costs(cfg, x0, U, eps)
    for k in range(4):
        x1 = oracle64.model_step(cfg["mass"], cfg["dt"], 2, 1, x0, [U[0] + eps[k, 0]])
        q = oracle64.cost_state(x1, cfg["goal"], cfg["q"])[0]
        ac = oracle64.cost_action(cfg["lambda"], cfg["sigma"], U[0], eps[k])[0]
        np.testing.assert_allclose(got[k], 2 * q + ac, rtol=1e-12)


def test_partials_merge_equals_update(oracle64):
    """The (beta_r, eta_r, N_r) rank partials merged with log-sum-exp rescaling reproduce the
    single-pass weighted noise."""
    rng = np.random.default_rng(7)
    k, T, a, lam = 96, 6, 2, 0.7
    costs = rng.uniform(0, 30, k)
    eps = rng.standard_normal((k, T, a))
    want = oracle64.update_stages(lam, costs, eps)["weighted_noise"]
    parts = [oracle64.partial(lam, costs, eps, lo, lo + 32) for lo in range(0, k, 32)]
    beta = min(p[0] for p in parts)
    eta = sum(np.exp(-(p[0] - beta) / lam) * p[1] for p in parts)
    N = sum(np.exp(-(p[0] - beta) / lam) * p[2] for p in parts)
    np.testing.assert_allclose(N / eta, want, rtol=1e-12, atol=1e-14)


# ---- noise stream contract --------------------------------------------------------------------
def test_philox_kats():
    from oracle.pyoracle import philox4x32_10
    for kat in kats.PHILOX_KATS:
        assert philox4x32_10(kat["ctr"], kat["key"]) == kat["out"]
    for kat in kats.PHILOX7_KATS:                       # the seven-round variant (mppi_config.philox_rounds = 7)
        assert philox4x32_10(kat["ctr"], kat["key"], rounds=7) == kat["out"]


def test_philox_normals_moments():
    from oracle.pyoracle import philox_normals
    z = philox_normals(seed=1, update=0, stream=0, k0=0, k1=4096, n_per_sample=60).astype(np.float64)
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1) < 0.01
    assert abs((z ** 4).mean() - 3) < 0.1
    # rows independent of how the sample range is split (global sample index in the counter)
    z2 = philox_normals(seed=1, update=0, stream=0, k0=100, k1=110, n_per_sample=60)
    np.testing.assert_array_equal(z2, z[100:110].astype(np.float32))
    z7 = philox_normals(seed=1, update=0, stream=0, k0=0, k1=4096, n_per_sample=60, rounds=7).astype(np.float64)
    assert abs(z7.mean()) < 0.01 and abs(z7.var() - 1) < 0.01 and abs((z7 ** 4).mean() - 3) < 0.1
    assert np.abs(z7 - z).max() > 1.0                    # a different stream


# ---- MLP dynamics (parity unpinned in the reference; structure check only) ------------------------
def test_mlp_step_matches_numpy(oracle64):
    rng = np.random.default_rng(4)
    s, a, H = 6, 3, 16
    mlp = dict(W1=rng.standard_normal((s + a, H)), b1=rng.standard_normal(H),
               W2=rng.standard_normal((H, H)), b2=rng.standard_normal(H),
               W3=rng.standard_normal((H, s)), b3=rng.standard_normal(s),
               Xmean=rng.standard_normal(s + a), Xstd=1 + rng.random(s + a),
               Ymean=rng.standard_normal(s), Ystd=1 + rng.random(s))
    x, u = rng.standard_normal(s), rng.standard_normal(a)
    X = (np.concatenate([x, u]) - mlp["Xmean"]) / mlp["Xstd"]
    h = np.maximum(X @ mlp["W1"] + mlp["b1"], 0)
    h = np.maximum(h @ mlp["W2"] + mlp["b2"], 0)
    want = x + (h @ mlp["W3"] + mlp["b3"]) * mlp["Ystd"] + mlp["Ymean"]
    np.testing.assert_allclose(oracle64.mlp_step(mlp, x, u), want, rtol=1e-12)
