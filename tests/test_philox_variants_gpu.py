"""The Philox-mode kernel family (round 2): the superposition kernels (mppi_linear.cuh) in their regenerating and
resident-tile forms, the direct-form kernel, and the 7 / 10 round generator.  Every variant is checked by
store-then-replay against the fp64 oracle (flat 1e-5) and against the other variants on the same stream.
MPPI_PHILOX_KERNEL (read at mppi_create) forces a variant: regen | resident | direct."""
import contextlib
import os

import numpy as np
import pytest

from tests.golden import kats
from tests.util import assert_update_close, controller_from_cfg, make_cfg, rel_err

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def kernel_variant(name):
    old = os.environ.get("MPPI_PHILOX_KERNEL")
    if name:
        os.environ["MPPI_PHILOX_KERNEL"] = name
    else:
        os.environ.pop("MPPI_PHILOX_KERNEL", None)
    try:
        yield
    finally:
        if old is None:
            os.environ.pop("MPPI_PHILOX_KERNEL", None)
        else:
            os.environ["MPPI_PHILOX_KERNEL"] = old


def test_philox7_raw_bit_exact():
    """Device Philox4x32-7 == oracle == Random123's seven-round known-answer vectors."""
    from mppi_tf_b200 import philox_raw
    from oracle.pyoracle import philox4x32_10
    for kat in kats.PHILOX7_KATS:
        seed = kat["key"][0] | (kat["key"][1] << 32)
        got = philox_raw(seed, kat["ctr"][0], kat["ctr"][1], kat["ctr"][2], kat["ctr"][3], 1, rounds=7)
        assert [int(v) for v in got[0]] == kat["out"]
    got = philox_raw(987654321987654321, 3, 1234567, 8, 2, 48, rounds=7)
    for i in range(48):
        want = philox4x32_10([3 + i, 1234567, 8, 2], [987654321987654321 & 0xFFFFFFFF, 987654321987654321 >> 32], rounds=7)
        assert [int(v) for v in got[i]] == want


def _one(cfg, x0, U0, variant, rounds, seed=21):
    with kernel_variant(variant):
        ctrl = controller_from_cfg(cfg, seed=seed, philox_rounds=rounds)
    try:
        ctrl.setSequence(U0)
        act = ctrl.next(x0)
        out = dict(next=act, U_new=ctrl.getUpdate(), U_shift=ctrl.getSequence(), costs=ctrl.getCosts(), eps=ctrl.dumpNoise(),
                   stats=ctrl.getWeightStats())
    finally:
        ctrl.close()
    return out


SHAPES = [
    # k, tau, a, lam
    (1024, 20, 1, 1.0),      # config 1
    (8192, 50, 2, 1.0),      # config 2 rows (25 calls per row)
    (4096, 30, 2, 1.0),      # config 5 rows (15 calls per row)
    (2048, 100, 3, 1.0),     # config 3 rows: too long for the resident kernel
    (1000, 7, 2, 0.3),       # T not a multiple of 4, ragged K
    (33, 3, 3, 1.0), (1, 1, 1, 1.0), (31, 5, 4, 2.0),
    (3000, 35, 4, 0.5),      # 35 calls per row: two accumulators per lane in the resident kernel
    (5000, 16, 5, 1.0), (700, 9, 6, 1.0), (640, 6, 7, 1.0), (515, 8, 8, 1.0),
    (40000, 12, 2, 0.02),    # sparse weights, several tiles per warp
    (40000, 12, 2, 500.0),   # dense weights
]


@pytest.mark.parametrize("rounds", [7, 10])
@pytest.mark.parametrize("k,tau,a,lam", SHAPES)
def test_variants_store_then_replay(oracle32, oracle64, k, tau, a, lam, rounds):
    rng = np.random.default_rng(k * 31 + tau)
    cfg = make_cfg(k, tau, 2 * a, a, lam=lam, goal=rng.uniform(-1, 1, 2 * a), q=0.5 + 2 * rng.random(2 * a), mass=1.5)
    x0 = rng.uniform(-1, 1, 2 * a).astype(np.float32)
    U0 = (0.2 * rng.standard_normal((tau, a))).astype(np.float32)
    outs = {}
    for variant in ("regen", "resident", "direct", None):
        from mppi_tf_b200 import MppiError, _capi
        try:
            outs[variant] = _one(cfg, x0, U0, variant, rounds)
        except MppiError as e:                      # rows too long for the resident kernel: the forced variant must say so
            assert variant == "resident" and e.code == _capi.MPPI_ERR_UNSUPPORTED, (variant, str(e))
            continue
    ref_eps = outs["direct"]["eps"]
    r64 = oracle64.mppi_update(cfg, x0, U0, ref_eps)
    for variant, o in outs.items():
        what = f"{variant} r{rounds}"
        # the same stream in every variant (z_scale * n against the plain Box-Muller: a few ulp)
        np.testing.assert_allclose(o["eps"], ref_eps, rtol=0, atol=2e-6 * np.abs(ref_eps).max())
        ro = oracle64.mppi_update(cfg, x0, U0, o["eps"])
        for key in ("U_new", "next", "U_shift"):
            assert np.abs(np.asarray(o[key], np.float64) - ro[key]).max() <= 1e-5 * np.abs(ro["U_new"]).max(), f"{key} {what}"
        assert rel_err(o["costs"], ro["costs"]) < 1e-5, what
        beta, eta = o["stats"]
        assert abs(float(beta[0]) - ro["costs"].min()) <= 1e-5 * np.abs(ro["costs"]).max(), what
        # against the direct form on (nearly) the same noise
        assert rel_err(o["U_new"], outs["direct"]["U_new"]) < 2e-5, what
    assert rel_err(outs["direct"]["U_new"], r64["U_new"]) < 1e-5


@pytest.mark.parametrize("variant", ["regen", "resident"])
def test_superposition_batched_controllers(oracle64, variant):
    """Independent controllers (own state, goal, sequence, noise stream) through the superposition kernels."""
    n, k, tau, a = 19, 1024, 30, 2
    rng = np.random.default_rng(6)
    goals = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
    xs = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
    U0 = (0.1 * rng.standard_normal((n, tau, a))).astype(np.float32)
    cfg = make_cfg(k, tau, 4, a)
    from mppi_tf_b200 import ControllerBase
    with kernel_variant(variant):
        ctrl = ControllerBase(k, tau, cfg["dt"], cfg["mass"], 4, a, lam=1.0, sigma=cfg["sigma"], goal=goals, Q=cfg["q"],
                              n_controllers=n, goal_per_controller=True, philox_rounds=7, seed=5)
    try:
        ctrl.setSequence(U0)
        act = ctrl.next(xs)
        Un, costs, eps = ctrl.getUpdate(), ctrl.getCosts(), ctrl.dumpNoise()
    finally:
        ctrl.close()
    assert not np.array_equal(eps[0], eps[1])
    for c in range(n):
        ref = oracle64.mppi_update(dict(cfg, goal=goals[c]), xs[c], U0[c], eps[c])
        assert_update_close(Un[c], ref["U_new"], what=f"U_new[{c}] {variant}")
        assert np.abs(act[c] - ref["next"]).max() <= 1e-5 * np.abs(ref["U_new"]).max()
        assert rel_err(costs[c], ref["costs"]) < 1e-5


def test_superposition_normalised_and_python_cost(oracle64):
    """Cost normalisation (two launches) and the gamma action cost run in the superposition form too (regenerating
    kernel); the noise-quadratic term (upsilon != 1) falls back to the direct form."""
    k, tau, a = 6000, 20, 2
    cfg = make_cfg(k, tau, 4, a, lam=0.7)
    rng = np.random.default_rng(3)
    x0 = rng.uniform(-1, 1, 4).astype(np.float32)
    U0 = (0.2 * rng.standard_normal((tau, a))).astype(np.float32)
    for gamma, upsilon, norm in ((0.4, 1.0, False), (0.4, 1.0, True), (0.9, 1.7, True)):
        ctrl = controller_from_cfg(cfg, seed=13, philox_rounds=7)
        try:
            ctrl.setActionCost("python", gamma=gamma, upsilon=upsilon)
            ctrl.setNormalizeCost(norm)
            ctrl.setSequence(U0)
            act = ctrl.next(x0)
            Un, costs, eps = ctrl.getUpdate(), ctrl.getCosts(), ctrl.dumpNoise()
        finally:
            ctrl.close()
        ref = oracle64.mppi_update_py(cfg, x0, U0, eps, gamma=gamma, upsilon=upsilon, normalize=norm)
        assert_update_close(Un, ref["U_new"], what=f"gamma {gamma} upsilon {upsilon} norm {norm}")
        assert rel_err(costs, ref["costs"]) < 1e-5
        assert np.abs(act - ref["next"]).max() <= 1e-5 * np.abs(ref["U_new"]).max()


def test_zero_q_takes_the_direct_form(oracle64):
    """q_i = 0 cannot be folded into the scaled noise state: such a handle must run (direct form) and match."""
    k, tau, a = 2048, 12, 2
    cfg = make_cfg(k, tau, 4, a, q=[1.0, 0.0, 2.0, 0.0])
    rng = np.random.default_rng(4)
    x0 = rng.uniform(-1, 1, 4).astype(np.float32)
    U0 = (0.2 * rng.standard_normal((tau, a))).astype(np.float32)
    o = _one(cfg, x0, U0, None, 7)
    ref = oracle64.mppi_update(cfg, x0, U0, o["eps"])
    assert_update_close(o["U_new"], ref["U_new"], what="q with zeros")
    assert rel_err(o["costs"], ref["costs"]) < 1e-5


def test_philox7_noise_statistics_and_stream():
    from oracle.pyoracle import philox_normals
    k, tau, a = 65536, 20, 2
    cfg = make_cfg(k, tau, 4, a, sigma=np.eye(a))
    ctrl = controller_from_cfg(cfg, seed=77, philox_rounds=7)
    try:
        ctrl.setUpdateCounter(9)
        ctrl.next(np.zeros(4, np.float32))
        z = ctrl.dumpNoise().astype(np.float64)
    finally:
        ctrl.close()
    want = philox_normals(seed=77, update=9, stream=0, k0=0, k1=512, n_per_sample=tau * a, rounds=7).reshape(512, tau, a)
    np.testing.assert_allclose(z[:512], want, rtol=0, atol=2e-5)
    flat = z.reshape(-1, a)
    assert np.abs(flat.mean(0)).max() < 3e-3
    np.testing.assert_allclose(flat.T @ flat / flat.shape[0], np.eye(a), atol=3e-3)
    assert abs((flat ** 4).mean() - 3.0) < 0.05
    assert abs(np.mean(flat[:-1, 0] * flat[1:, 0])) < 3e-3          # along the stream
    assert abs(np.mean(z[:-1, :, :] * z[1:, :, :])) < 1e-3           # between neighbouring samples (counter word 1)
