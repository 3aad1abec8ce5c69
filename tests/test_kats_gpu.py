"""The reference's unit KATs (test/test_{utile,model,cost,controller}.cpp) replayed against the
CUDA stage entry points of the C-ABI.  Tolerance as in the reference: EXPECT_FLOAT_EQ (4 ULP)."""
import numpy as np
import pytest

from tests.golden import kats

pytestmark = pytest.mark.gpu

FLOAT_EQ = dict(rtol=4 * np.finfo(np.float32).eps, atol=1e-30)
f32 = lambda v: np.asarray(v, np.float32)


@pytest.fixture(scope="module")
def ctrl():
    from mppi_tf_b200 import ControllerBase
    # fixture of test/test_controller.cpp:17-19: k=5, tau=3, dt=0.01, mass=1, s=4, a=2
    c = ControllerBase(5, 3, 0.01, 1.0, 4, 2)
    yield c
    c.close()


@pytest.mark.parametrize("nb", [1, 2, 3, 4])
def test_block_diag(nb):
    from mppi_tf_b200 import blockDiag
    dt, m = np.float32(kats.UTILE_DT), np.float32(kats.UTILE_M)
    A_blk = f32([[1, dt], [0, 1]])
    B_blk = f32([[dt * dt / (np.float32(2) * m)], [dt / m]])
    exp_a, exp_b = kats.blockdiag_expected(nb)
    got_a, got_b = blockDiag(A_blk, nb), blockDiag(B_blk, nb)
    assert got_a.shape == (2 * nb, 2 * nb) and got_b.shape == (2 * nb, nb)
    np.testing.assert_allclose(got_a, exp_a, **FLOAT_EQ)
    np.testing.assert_allclose(got_b, exp_b, **FLOAT_EQ)


@pytest.mark.parametrize("case", kats.MODEL_CASES, ids=lambda c: c["name"])
def test_model(case):
    from mppi_tf_b200 import ModelBase
    exp_s, exp_u, exp_res = kats.model_expected(case)
    model = ModelBase(case["m"], case["dt"], case["s"], case["a"])
    free = model.freeStep(case["state"])
    act = model.actionStep(case["action"])
    full = model.predict(case["state"], case["action"])
    assert free.shape == (len(case["state"]), case["s"])
    assert act.shape == (case["k"], case["s"]) and full.shape == (case["k"], case["s"])
    np.testing.assert_allclose(free, exp_s, **FLOAT_EQ)
    np.testing.assert_allclose(act, exp_u, **FLOAT_EQ)
    np.testing.assert_allclose(full, exp_res, **FLOAT_EQ)


def test_model_three_steps_py_twin():
    from mppi_tf_b200 import ModelBase
    c = kats.py_step3_expected()
    model = ModelBase(c["m"], c["dt"], 6, 3)
    x = c["state"]
    for _ in range(3):
        x = model.predict(x, c["action"])
    np.testing.assert_allclose(x, c["expected"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("case", kats.COST_CASES, ids=lambda c: c["name"])
def test_cost(case):
    from mppi_tf_b200 import CostBase
    cost = CostBase(case["lam"], case["sigma"], case["goal"], case["q"])
    st = cost.finalCost(case["state"])
    step = cost.stepCost(case["state"], case["action"], case["noise"])
    assert st.shape == (case["k"],) and step.shape == (case["k"],)
    np.testing.assert_allclose(st, f32(case["exp_state"]), **FLOAT_EQ)
    np.testing.assert_allclose(step, f32(case["exp_step"]), **FLOAT_EQ)


def test_cost_set_goal_takes_effect():
    from mppi_tf_b200 import CostBase
    case = kats.COST_CASES[1]
    cost = CostBase(case["lam"], case["sigma"], case["goal"], case["q"])
    assert cost.setGoal([0, 0.5, 2, 0]) is True         # goal == state -> zero cost
    np.testing.assert_array_equal(cost.stateCost(case["state"]), [0.0])
    assert cost.setGoal([1, 2, 3]) is False             # wrong size, src/controller_base.cpp:127-130


@pytest.mark.parametrize("case", kats.PY_ACTION_COST_CASES, ids=lambda c: c["name"])
def test_python_action_cost(case):
    """gamma / upsilon action cost KATs of the Python twin, scripts/test.py:685-838."""
    from mppi_tf_b200 import actionCostPython
    a = len(case["action"])
    got = actionCostPython(case["lam"], case["gamma"], case["upsilon"], np.eye(a), case["action"], case["noise"])
    np.testing.assert_allclose(got, case["expected"], rtol=1e-6, atol=1e-6)


def test_python_static_cost_s13():
    """scripts/test.py:944-1095 (13-dimensional state, a = 6, diagonal Q) on the GPU stages."""
    from mppi_tf_b200 import CostBase, actionCostPython
    c = kats.PY_STATIC13
    cost = CostBase(c["lam"], np.eye(6), c["goal"], c["q"])
    np.testing.assert_allclose(cost.stateCost(np.asarray(c["state"])), c["expected_state"], rtol=1e-6)
    got = actionCostPython(c["lam"], c["gamma"], c["upsilon"], np.eye(6), c["action"], c["noise"])
    np.testing.assert_allclose(got, c["expected_action"], rtol=1e-6)


@pytest.mark.parametrize("case", kats.ELLIPSE_CASES, ids=lambda c: c["name"])
def test_ellipse_cost(case):
    """ElipseCost.state_cost KATs, scripts/test.py:1098-1161 (assertAllClose: rtol 1e-6)."""
    from mppi_tf_b200 import ellipseStateCost
    got = ellipseStateCost(case["state"], *kats.ELLIPSE_PARAMS)
    np.testing.assert_allclose(got, case["expected"], rtol=1e-6, atol=1e-6)


def test_data_prep(ctrl):
    for t in range(3):
        np.testing.assert_allclose(ctrl.prepareAction(kats.CTRL["action"], t), f32(kats.CTRL_PREP["a"][t]), **FLOAT_EQ)
        got = ctrl.prepareNoise(kats.CTRL["noise"], t)
        assert got.shape == (5, 2)
        np.testing.assert_allclose(got, f32(kats.CTRL_PREP["n"][t]), **FLOAT_EQ)


def test_update_stages(ctrl):
    r = ctrl.updateStages(kats.CTRL["cost"], kats.CTRL["noise"], lam=kats.CTRL["lam"])
    e = kats.CTRL_UPDATE
    np.testing.assert_allclose(r["beta"], f32(e["beta"]), **FLOAT_EQ)
    np.testing.assert_allclose(r["exp_arg"], f32(e["exp_arg"]), **FLOAT_EQ)
    np.testing.assert_allclose(r["exp"], f32(e["exp"]), **FLOAT_EQ)
    np.testing.assert_allclose(r["nabla"], f32(e["nabla"]), **FLOAT_EQ)
    np.testing.assert_allclose(r["weights"], f32(e["weights"]), **FLOAT_EQ)
    np.testing.assert_allclose(r["weighted_noise"], f32(e["weighted_noise"]), **FLOAT_EQ)
    np.testing.assert_allclose(r["weights"].sum(dtype=np.float32), f32(e["sum_w"]), **FLOAT_EQ)


def test_get_new(ctrl):
    for nb, want in kats.CTRL_NEW.items():
        got = ctrl.getNew(kats.CTRL["action"], nb)
        assert got.shape == (nb, 2)                      # nb = 0: the empty [0,2,1] tensor
        np.testing.assert_allclose(got, f32(want).reshape(nb, 2), **FLOAT_EQ)


def test_shift(ctrl):
    for c in kats.CTRL_SHIFT:
        got = ctrl.shift(kats.CTRL["action"], c["init"], c["nb"])
        np.testing.assert_allclose(got, f32(c["expected"]), **FLOAT_EQ)


def test_set_goal_size_check(ctrl):
    assert ctrl.setGoal([1, 1, 1]) is False
    assert ctrl.setGoal([1, 1, 1, 2]) is True


def test_philox_raw_bit_exact():
    """Integer contract of the noise stream: device Philox4x32-10 == oracle == Random123 KATs."""
    from mppi_tf_b200 import philox_raw
    from oracle.pyoracle import philox4x32_10
    kat = kats.PHILOX_KATS[2]
    seed = kat["key"][0] | (kat["key"][1] << 32)
    got = philox_raw(seed, kat["ctr"][0], kat["ctr"][1], kat["ctr"][2], kat["ctr"][3], 1)
    assert [int(v) for v in got[0]] == kat["out"]
    got = philox_raw(0, 0, 0, 0, 0, 1)
    assert [int(v) for v in got[0]] == kats.PHILOX_KATS[0]["out"]
    got = philox_raw(12345678901234567, 5, 77, 3, 9, 64)
    for i in range(64):
        want = philox4x32_10([5 + i, 77, 3, 9], [12345678901234567 & 0xFFFFFFFF, 12345678901234567 >> 32])
        assert [int(v) for v in got[i]] == want
