"""ElipseCost3D: golden vectors from the reference's own class run on the numpy shim
(tests/golden/gen_ellipse3d_fixtures.py, which first replays the reference's known-answer tests for the class), against
the C restatement (CPU) and the CUDA functor (GPU)."""
import os

import numpy as np
import pytest

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ellipse3d_fixtures.npz")
CASES = ["e3_xy", "e3_tilt", "e3_gen"]


def load(name):
    d = np.load(FIX)
    return lambda key: d[f"{name}_{key}"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_class(oracle64, name):
    g = load(name)
    R, q = oracle64.ellipse3d_prep(g("normal"), g("aVec"))
    np.testing.assert_allclose(R, g("R"), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(q, g("q"), rtol=1e-12, atol=1e-13)
    speed, ms, mv = g("scal")
    e3 = oracle64.ellipse3d_pack(g("normal"), g("aVec"), g("axis"), speed, ms, mv)
    np.testing.assert_allclose(oracle64.cost_state_ellipse3d(g("state"), e3), g("cost"), rtol=1e-10, atol=1e-10)


def test_oracle_reference_kats(oracle64):
    """scripts/test.py:1183-1300 on the restatement: prep_const, and orientation / position errors through states already
    expressed in the plane frame (identity plane quaternion)."""
    R, _ = oracle64.ellipse3d_prep([0, 1, 1], [1, 0, 0])
    np.testing.assert_allclose(R, np.array([[1, 0, 0], [0, .5, -.5], [0, .5, .5]]).T, atol=1e-12)
    e3 = oracle64.ellipse3d_pack([0, 0, 1], [1, 0, 0], [2.0, 1.5], 1.0, 1.0, 0.0)            # m_vel = 0: position + orientation
    st = np.zeros((3, 13))
    st[:, :7] = [[0.1, 0.4, 0.2, 0, 0, 0, 1], [1, 1, -2, 0.48038446, 0.32025631, 0.16012815, 0.80064077],
                 [2, 1, -2, 0.20628425, -0.30942637, -0.92827912, 0.]]
    pos = np.array([0.8863888888888889, 3.6944444444444446, 0.0])
    pos[2] = abs((2 / 2.0) ** 2 + (1 / 1.5) ** 2 + 4.0 - 1.0)
    ori = np.array([3.0018837793006306, 2.4098026419889416, 1.1216620246733544])
    np.testing.assert_allclose(oracle64.cost_state_ellipse3d(st, e3), pos + ori, rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference_class(name):
    from mppi_tf_b200 import ellipse3dStateCost
    g = load(name)
    speed, ms, mv = g("scal")
    got = ellipse3dStateCost(g("state"), g("normal"), g("aVec"), g("axis"), g("center"), speed, ms, mv)
    np.testing.assert_allclose(got, g("cost"), rtol=2e-5, atol=2e-5)


@pytest.mark.gpu
def test_cuda_update_with_ellipse3d_cost(oracle64, oracle32):
    """Full AUV update with ElipseCost3D as the state cost (injected noise, then Philox store-then-replay).  The reference
    cannot produce this vector itself (its k > 1 cost sum mis-broadcasts), so the checker composes the pinned pieces."""
    import json
    from mppi_tf_b200 import ControllerBase
    from tests.util import assert_update_close, rel_err
    d = np.load(os.path.join(os.path.dirname(FIX), "auv_fixtures.npz"))
    prm = json.loads(bytes(d["params_json"]).decode())["full"]
    g = load("e3_gen")
    speed, ms, mv = g("scal")
    rng = np.random.default_rng(11)
    k, tau, lam, gamma, ups = 3000, 9, 0.9, 0.6, 1.3
    L = 4.0 * rng.standard_normal((6, 6))
    sigma = (L @ L.T + 60.0 * np.eye(6)).astype(np.float32)
    U = (20.0 * rng.standard_normal((tau, 6))).astype(np.float32)
    x = rng.uniform(-0.5, 0.5, 13)
    x[3:7] /= np.linalg.norm(x[3:7])
    x = x.astype(np.float32)
    eps = np.einsum("ij,ktj->kti", ups * sigma, rng.standard_normal((k, tau, 6))).astype(np.float32)
    e3 = oracle64.ellipse3d_pack(g("normal"), g("aVec"), g("axis"), speed, ms, mv)
    kw = dict(gamma=gamma, upsilon=ups, ellipse3d=e3)
    ctrl = ControllerBase(k, tau, 0.1, 1.0, 13, 6, lam=lam, sigma=sigma, model="auv")
    try:
        ctrl.setAuvModel(prm, rk=2)
        ctrl.setEllipse3dCost(g("normal"), g("aVec"), g("axis"), g("center"), speed, ms, mv)
        ctrl.setActionCost("python", gamma=gamma, upsilon=ups)
        for mode in ("injected", "philox"):
            ctrl.setSequence(U)
            if mode == "injected":
                ctrl.nextWithNoise(x, eps)
                e = eps
            else:
                ctrl.next(x)
                e = ctrl.dumpNoise().reshape(k, tau, 6)
            r64 = oracle64.mppi_update_auv(prm, 0.1, 2, lam, sigma, np.zeros(13), np.ones(13), x, U, e, **kw)
            r32 = oracle32.mppi_update_auv(prm, 0.1, 2, lam, sigma, np.zeros(13), np.ones(13), x, U, e, **kw)
            assert rel_err(ctrl.getCosts(), r64["costs"]) < 2e-5, mode
            assert_update_close(ctrl.getUpdate(), r64["U_new"], r32["U_new"], what="ellipse3d " + mode)
    finally:
        ctrl.close()
