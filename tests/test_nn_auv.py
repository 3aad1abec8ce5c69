"""The reference's own learned model, NNAUVModel (scripts/src/models/nn_model.py:181-304: 16 -> 32 -> 32 -> 32 -> 13 on the
AUV state without its position, next = state + delta), as a dynamics functor of the AUV controller.  Fixtures come from
the reference's class and its ControllerBase on the numpy TF shim (tests/golden/gen_nn_auv_fixtures.py): the model's forward
values are PINNED here, unlike the point-mass MLP of BASELINE config 4."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = np.load(os.path.join(HERE, "golden", "nn_auv_fixtures.npz"))


def _nn():
    return dict(W=[FIX[f"W{i}"] for i in range(4)], b=[FIX[f"b{i}"] for i in range(4)], Xmean=FIX["Xmean"], Xstd=FIX["Xstd"],
                Ymean=FIX["Ymean"], Ystd=FIX["Ystd"])


def _upd(name):
    g = lambda k: FIX[f"{name}_{k}"]
    k, tau, lam, gamma, upsilon, norm = g("meta")
    return g, int(k), int(tau), float(lam), float(gamma), float(upsilon), bool(norm)


def test_oracle_step_matches_the_reference(oracle64):
    got = oracle64.nn_auv_step(_nn(), FIX["step_state"], FIX["step_action"])
    np.testing.assert_allclose(got, FIX["step_next"], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("name", ["upd1", "upd2"])
def test_oracle_update_matches_the_reference(oracle64, name):
    g, k, tau, lam, gamma, upsilon, norm = _upd(name)
    r = oracle64.mppi_update_nn_auv(_nn(), lam, g("sigma"), g("goal"), g("q"), g("x"), g("U"), g("eps"), gamma=gamma,
                                    upsilon=upsilon, normalize=norm)
    np.testing.assert_allclose(r["costs"], g("costs"), rtol=1e-10)
    np.testing.assert_allclose(r["U_new"], g("U_new"), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(r["next"], g("next"), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(r["U_shift"], g("U_shift"), rtol=1e-9, atol=1e-11)


@pytest.mark.gpu
def test_cuda_nn_auv_predict_matches_the_reference():
    from mppi_tf_b200 import ControllerBase
    ctrl = ControllerBase(64, 4, 0.1, 1.0, 13, 6, model="auv")
    try:
        ctrl.setNnAuvModel(_nn())
        got = ctrl.auvPredict(FIX["step_state"], FIX["step_action"])
        d_ref = FIX["step_next"] - FIX["step_state"]
        assert np.abs((got - FIX["step_state"].astype(np.float32)) - d_ref).max() <= 1e-5 * np.abs(d_ref).max()
        got1 = ctrl.auvPredict(FIX["step_state"][:1], FIX["step_action"])      # one state broadcast over the actions
        assert got1.shape == (64, 13) and np.abs(got1[0] - got[0]).max() == 0
    finally:
        ctrl.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["upd1", "upd2"])
def test_cuda_nn_auv_update_matches_the_reference(oracle64, oracle32, name):
    """Injected noise against the reference's own controller output, then Philox store-then-replay against the oracle."""
    from mppi_tf_b200 import ControllerBase
    from tests.util import assert_update_close, rel_err
    g, k, tau, lam, gamma, upsilon, norm = _upd(name)
    ctrl = ControllerBase(k, tau, 0.1, 1.0, 13, 6, lam=lam, sigma=g("sigma"), goal=g("goal"), Q=g("q"), model="auv")
    try:
        ctrl.setNnAuvModel(_nn())
        ctrl.setActionCost("python", gamma=gamma, upsilon=upsilon)
        ctrl.setNormalizeCost(norm)
        ctrl.setSequence(g("U"))
        act = ctrl.nextWithNoise(g("x"), g("eps"))
        assert rel_err(ctrl.getCosts(), g("costs")) < 1e-5
        r32 = oracle32.mppi_update_nn_auv(_nn(), lam, g("sigma"), g("goal"), g("q"), g("x"), g("U"), g("eps"), gamma=gamma,
                                          upsilon=upsilon, normalize=norm)
        assert_update_close(ctrl.getUpdate(), g("U_new"), r32["U_new"], what=name + " U_new", allow_fp32_distance=True)
        assert np.abs(act - g("next")).max() <= max(1e-5, 3 * rel_err(r32["U_new"], g("U_new"))) * np.abs(g("U_new")).max()
        ctrl.setSequence(g("U"))
        ctrl.next(g("x"))
        eps = ctrl.dumpNoise()
        r64 = oracle64.mppi_update_nn_auv(_nn(), lam, g("sigma"), g("goal"), g("q"), g("x"), g("U"), eps, gamma=gamma,
                                          upsilon=upsilon, normalize=norm)
        r32 = oracle32.mppi_update_nn_auv(_nn(), lam, g("sigma"), g("goal"), g("q"), g("x"), g("U"), eps, gamma=gamma,
                                          upsilon=upsilon, normalize=norm)
        assert rel_err(ctrl.getCosts(), r64["costs"]) < 1e-5
        assert_update_close(ctrl.getUpdate(), r64["U_new"], r32["U_new"], what=name + " philox U_new", allow_fp32_distance=True)
    finally:
        ctrl.close()


@pytest.mark.gpu
def test_cuda_nn_auv_larger_rollout(oracle64):
    """Ragged multi-CTA size, narrower network (hidden 16, two hidden layers: zero padded to the kernel's 32-wide layers)."""
    from mppi_tf_b200 import ControllerBase
    from tests.util import rel_err
    rng = np.random.default_rng(3)
    dims = [16, 16, 16, 13]
    nn = dict(W=[rng.uniform(-0.4, 0.4, (dims[i], dims[i + 1])) for i in range(3)], b=[0.05 * rng.standard_normal(dims[i + 1]) for i in range(3)],
              Xmean=0.1 * rng.standard_normal(16), Xstd=1 + rng.random(16), Ymean=0.002 * rng.standard_normal(13), Ystd=0.01 + 0.02 * rng.random(13))
    k, tau, lam = 5000, 12, 2.0
    sigma = np.diag(10.0 + 20.0 * rng.random(6))
    goal = rng.uniform(-1, 1, 13); goal[3:7] /= np.linalg.norm(goal[3:7])
    x = goal + 0.3 * rng.standard_normal(13); x[3:7] /= np.linalg.norm(x[3:7])
    U = 5.0 * rng.standard_normal((tau, 6))
    ctrl = ControllerBase(k, tau, 0.1, 1.0, 13, 6, lam=lam, sigma=sigma, goal=goal, model="auv", seed=4)
    try:
        ctrl.setNnAuvModel(nn)
        ctrl.setActionCost("python", gamma=lam, upsilon=1.0)
        ctrl.setSequence(U)
        ctrl.next(x)
        eps = ctrl.dumpNoise()
        r64 = oracle64.mppi_update_nn_auv(nn, lam, sigma, goal, np.ones(13), x, U, eps)
        assert rel_err(ctrl.getCosts(), r64["costs"]) < 1e-5
        assert rel_err(ctrl.getUpdate(), r64["U_new"]) < 2e-5
    finally:
        ctrl.close()
