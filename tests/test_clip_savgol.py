"""The last two Python-twin extras of SURVEY.md section 8(f) row N1: action clipping (clip_act,
controllers/controller_base.py:500-504) and the Savitzky-Golay pass over the sequence (:281-291).  Fixtures come from
the reference's own code on the numpy TF shim (tests/golden/gen_clip_savgol_fixtures.py)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = np.load(os.path.join(HERE, "golden", "clip_savgol_fixtures.npz"))
CASES = ["clip1d", "clip2d", "clip3d"]


def _case(name):
    g = lambda k: FIX[f"{name}_{k}"]
    k, tau, s, a, mass, dt, lam = g("meta")
    return g, int(k), int(tau), int(s), int(a), float(mass), float(dt), float(lam)


@pytest.mark.parametrize("name", CASES)
def test_oracle_update_then_clip_matches_the_reference(oracle64, name):
    """The oracle's update followed by the clip reproduces the reference's update -> clip_act -> get_next -> shift."""
    g, k, tau, s, a, mass, dt, lam = _case(name)
    cfg = dict(k=k, tau=tau, s_dim=s, a_dim=a, dt=dt, mass=mass, sigma=g("sigma"), goal=g("goal"), q=g("q"))
    cfg["lambda"] = lam
    r = oracle64.mppi_update_py(cfg, g("x"), g("U"), g("eps"), gamma=lam, upsilon=1.0)
    np.testing.assert_allclose(r["U_new"], g("U_raw"), rtol=1e-10, atol=1e-12)
    lo, hi = np.broadcast_to(g("lim_min"), (a,)), np.broadcast_to(g("lim_max"), (a,))
    clipped = np.clip(r["U_new"], lo, hi)
    np.testing.assert_allclose(clipped, g("U_new"), rtol=1e-10, atol=1e-12)
    assert (g("U_new") != g("U_raw")).any()                    # the case really clips
    np.testing.assert_allclose(clipped[0], g("next"), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(np.vstack([clipped[1:], np.zeros((1, a))]), g("U_shift"), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_savgol_matches_scipy_and_the_fixture(name):
    """mppi_savgol_filter (host code of the C-ABI, no GPU needed) against the reference's exact scipy call."""
    import ctypes as C
    import scipy.signal
    from mppi_tf_b200 import _capi
    lib = _capi.load()
    g, k, tau, s, a, mass, dt, lam = _case(name)
    fp = C.POINTER(C.c_float)
    U = np.ascontiguousarray(g("U_new"), np.float32)
    out = np.empty_like(U)
    assert lib.mppi_savgol_filter(tau, a, U.ctypes.data_as(fp), 10, 9, out.ctypes.data_as(fp)) == 0
    # a degree-9 fit through 10 points is ill-conditioned: scipy's own result moves by ~1e-6 with the least-squares driver
    np.testing.assert_allclose(out, g("savgol_10_9"), rtol=0, atol=2e-5 * np.abs(U).max())
    rng = np.random.default_rng(3)
    x = rng.standard_normal((40, 3)).astype(np.float32)
    for window, order in ((5, 3), (7, 2), (11, 4), (6, 3), (8, 4), (40, 5)):
        y = np.empty_like(x)
        assert lib.mppi_savgol_filter(40, 3, x.ctypes.data_as(fp), window, order, y.ctypes.data_as(fp)) == 0
        want = scipy.signal.savgol_filter(x.astype(np.float64), window, order, deriv=0, delta=1.0, axis=0)
        np.testing.assert_allclose(y, want, rtol=0, atol=2e-6)
    assert lib.mppi_savgol_filter(8, 1, x.ctypes.data_as(fp), 10, 9, out.ctypes.data_as(fp)) != 0     # window > T
    assert lib.mppi_savgol_filter(40, 3, x.ctypes.data_as(fp), 5, 5, out.ctypes.data_as(fp)) != 0     # polyorder >= window


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("per_axis", [False, True])
def test_cuda_clip_act_matches_the_reference(name, per_axis):
    from mppi_tf_b200 import ControllerBase
    g, k, tau, s, a, mass, dt, lam = _case(name)
    lo, hi = g("lim_min"), g("lim_max")
    if per_axis and lo.size == 1:
        lo, hi = np.repeat(lo, a), np.repeat(hi, a)
    if not per_axis and lo.size != 1:
        pytest.skip("this fixture has one limit per axis")
    ctrl = ControllerBase(k, tau, dt, 1.0, s, a, lam=lam, sigma=g("sigma"), goal=g("goal"), Q=g("q"), model_mass=mass)
    try:
        ctrl.setActionCost("python", gamma=lam, upsilon=1.0)
        ctrl.setActionLimits(lo, hi)
        ctrl.setSequence(g("U"))
        act = ctrl.nextWithNoise(g("x"), g("eps"))
        scale = np.abs(g("U_raw")).max()
        assert np.abs(ctrl.getUpdate() - g("U_new")).max() <= 1e-5 * scale
        assert np.abs(act - g("next")).max() <= 1e-5 * scale
        assert np.abs(ctrl.getSequence() - g("U_shift")).max() <= 1e-5 * scale
        lo32, hi32 = np.broadcast_to(lo, (a,)).astype(np.float32), np.broadcast_to(hi, (a,)).astype(np.float32)
        assert (ctrl.getUpdate() <= hi32).all() and (ctrl.getUpdate() >= lo32).all()
        sg = ctrl.filterSequence(10, 9) if tau >= 10 else None
        if sg is not None:
            import scipy.signal
            want = scipy.signal.savgol_filter(ctrl.getSequence().astype(np.float64), 10, 9, deriv=0, delta=1.0, axis=0)
            np.testing.assert_allclose(sg, want, rtol=0, atol=2e-5 * scale)
        # Philox mode clips too, and turning the limits off restores the raw update
        ctrl.setSequence(g("U"))
        ctrl.next(g("x"))
        un = ctrl.getUpdate()
        assert (un <= hi32).all() and (un >= lo32).all()
        ctrl.setActionLimits(None, None)
        ctrl.setSequence(g("U"))
        ctrl.nextWithNoise(g("x"), g("eps"))
        assert np.abs(ctrl.getUpdate() - g("U_raw")).max() <= 1e-5 * scale
    finally:
        ctrl.close()
