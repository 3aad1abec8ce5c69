"""Philox mode (the hot path: noise regenerated in registers).  The TF RandomNormal stream cannot be
restated, so parity is by store-then-replay: dump the eps the kernel used and feed exactly that
tensor to the oracle, plus statistical checks of the generator and
rank-count independence of the sharded update."""
import numpy as np
import pytest

from tests.util import controller_from_cfg, make_cfg, rel_err, assert_update_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,tau,a,sigma_full", [
    (1024, 20, 1, False),     # config 1
    (8192, 50, 2, False),     # config 2 rows
    (4096, 100, 3, False),    # config 3 rows
    (3000, 25, 3, True),      # non-diagonal Sigma, ragged K
    (1000, 7, 2, False),      # T not a multiple of 4
    (777, 9, 4, True), (515, 6, 5, False), (300, 5, 8, True),
])
def test_store_then_replay(oracle32, oracle64, k, tau, a, sigma_full):
    rng = np.random.default_rng(k)
    sigma = 0.25 * np.eye(a)
    if sigma_full:
        L = 0.2 * rng.standard_normal((a, a))
        sigma = L @ L.T + 0.2 * np.eye(a)
    cfg = make_cfg(k, tau, 2 * a, a, sigma=sigma, lam=0.8)
    x0 = rng.uniform(-1, 1, 2 * a).astype(np.float32)
    U0 = (0.2 * rng.standard_normal((tau, a))).astype(np.float32)
    ctrl = controller_from_cfg(cfg, seed=42)
    try:
        ctrl.setSequence(U0)
        act = ctrl.next(x0)
        got = dict(next=act, U_new=ctrl.getUpdate(), U_shift=ctrl.getSequence(), costs=ctrl.getCosts())
        eps = ctrl.dumpNoise()
    finally:
        ctrl.close()
    assert eps.shape == (k, tau, a)
    r64 = oracle64.mppi_update(cfg, x0, U0, eps)
    r32 = oracle32.mppi_update(cfg, x0, U0, eps)
    for key in ("U_new", "next", "U_shift"):
        assert_update_close(got[key], r64[key], r32[key], what=key)
    np.testing.assert_allclose(got["costs"], r64["costs"], rtol=2e-5, atol=2e-5)


def test_replay_through_injected_path():
    """The dumped tensor fed back through the injected-noise kernel reproduces the Philox update."""
    cfg = make_cfg(4096, 40, 4, 2)
    rng = np.random.default_rng(1)
    x0 = rng.uniform(-1, 1, 4).astype(np.float32)
    U0 = (0.2 * rng.standard_normal((40, 2))).astype(np.float32)
    ctrl = controller_from_cfg(cfg, seed=7)
    try:
        ctrl.setSequence(U0)
        a1 = ctrl.next(x0)
        u1, c1 = ctrl.getUpdate(), ctrl.getCosts()
        eps = ctrl.dumpNoise()
        ctrl.setSequence(U0)
        a2 = ctrl.nextWithNoise(x0, eps)
        u2, c2 = ctrl.getUpdate(), ctrl.getCosts()
    finally:
        ctrl.close()
    assert rel_err(u2, u1) < 1e-5 and rel_err(a2, a1) < 1e-5
    np.testing.assert_allclose(c2, c1, rtol=1e-5, atol=1e-5)


def test_noise_matches_stream_specification():
    """z = Sigma^-1 eps equals the oracle's statement of the stream (Philox bits exact; Box-Muller
    through MUFU approximations, so a small absolute tolerance)."""
    from oracle.pyoracle import philox_normals
    k, tau, a = 512, 11, 3
    cfg = make_cfg(k, tau, 6, a, sigma=np.eye(a))
    ctrl = controller_from_cfg(cfg, seed=99)
    try:
        ctrl.setUpdateCounter(5)
        ctrl.next(np.zeros(6, np.float32))
        eps = ctrl.dumpNoise()
    finally:
        ctrl.close()
    want = philox_normals(seed=99, update=5, stream=0, k0=0, k1=k, n_per_sample=tau * a).reshape(k, tau, a)
    np.testing.assert_allclose(eps, want, rtol=0, atol=2e-5)


def test_noise_statistics():
    k, tau, a = 65536, 20, 2
    sigma = np.array([[0.5, 0.2], [0.0, 0.3]], np.float32)
    cfg = make_cfg(k, tau, 4, a, sigma=sigma)
    ctrl = controller_from_cfg(cfg, seed=1)
    try:
        ctrl.next(np.zeros(4, np.float32))
        e1 = ctrl.dumpNoise().astype(np.float64)
        ctrl.next(np.zeros(4, np.float32))
        e2 = ctrl.dumpNoise().astype(np.float64)
    finally:
        ctrl.close()
    flat = e1.reshape(-1, a)
    cov = flat.T @ flat / flat.shape[0]
    np.testing.assert_allclose(cov, sigma.astype(np.float64) @ sigma.T.astype(np.float64), atol=3e-3)
    assert np.abs(flat.mean(0)).max() < 3e-3
    z = np.linalg.solve(sigma.astype(np.float64), flat.T).T
    assert abs((z ** 4).mean() - 3.0) < 0.05                    # kurtosis of a normal
    assert abs(np.mean(z[:-1, 0] * z[1:, 0])) < 3e-3             # no lag-1 correlation along the stream
    # fresh noise on every update (the reference's RandomNormal is stateful, :196-199)
    assert np.abs(e1 - e2).max() > 0.1
    assert abs(np.mean(e1 * e2)) < 1e-3


def test_same_seed_same_update_is_deterministic():
    cfg = make_cfg(5000, 16, 4, 2)
    outs = []
    for _ in range(2):
        ctrl = controller_from_cfg(cfg, seed=3)
        try:
            outs.append((ctrl.next(np.full(4, 0.1, np.float32)), ctrl.getUpdate(), ctrl.getCosts()))
        finally:
            ctrl.close()
    for x, y in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("philox", [True, False])
def test_sample_sharding_is_rank_count_independent(world, philox):
    """K sharded over `world` ranks (emulated as `world` handles on one GPU, launched one after the
    other, with the all-gather done by a device copy) gives the 1-rank result: the Philox counter
    uses the global sample index and the (beta, eta, N) merge is exact."""
    import torch
    k, tau, a = 6000, 24, 3
    cfg = make_cfg(k, tau, 6, a, lam=0.7)
    rng = np.random.default_rng(2)
    x0 = rng.uniform(-1, 1, 6).astype(np.float32)
    U0 = (0.2 * rng.standard_normal((tau, a))).astype(np.float32)
    eps_full = None
    single = controller_from_cfg(cfg, seed=11)
    try:
        single.setSequence(U0)
        if philox:
            a_ref = single.next(x0)
            eps_full = single.dumpNoise()
        else:
            from tests.util import parity_noise
            eps_full = parity_noise(k, tau, a, cfg["sigma"])
            a_ref = single.nextWithNoise(x0, eps_full)
        u_ref, c_ref = single.getUpdate(), single.getCosts()
    finally:
        single.close()

    ranks = [controller_from_cfg(cfg, seed=11, rank=r, world=world) for r in range(world)]
    try:
        stride = ranks[0].exchangeStride()
        gathered = torch.zeros(world, stride, device="cuda")
        sends = [torch.zeros(stride, device="cuda") for _ in range(world)]
        keep = []
        for r, c in enumerate(ranks):
            c.setExchangeBuffers(sends[r].data_ptr(), gathered.data_ptr())
            c.setSequence(U0)
            c.setState(x0)
            if philox:
                c.enqueueUpdate()
            else:
                e = torch.from_numpy(np.ascontiguousarray(eps_full[c.k_offset:c.k_offset + c.k_local])).cuda()
                keep.append(e)
                c.enqueueUpdate(e.data_ptr())
            c.synchronize()
        for r in range(world):
            gathered[r].copy_(sends[r])          # the all-gather
        torch.cuda.synchronize()
        costs = []
        for c in ranks:
            c.enqueueFinish()
            act = c.fetchAction()
            assert rel_err(act, a_ref) < 1e-5
            assert rel_err(c.getUpdate(), u_ref) < 1e-5
            costs.append(c.getCosts())
        np.testing.assert_array_equal(np.concatenate(costs), c_ref)   # same samples, same arithmetic
        # every rank ends with the identical sequence (bit-wise: same merge order everywhere)
        seqs = [c.getSequence() for c in ranks]
        for s in seqs[1:]:
            np.testing.assert_array_equal(s, seqs[0])
    finally:
        for c in ranks:
            c.close()


@pytest.mark.parametrize("k,tau,a,lam,what", [
    (400000, 9, 3, 0.05, "sparse weights, several samples per thread: CTA-wide compaction"),
    (400000, 9, 3, 1e3, "dense weights: the direct loop"),
    (1500000, 4, 1, 2e-4, "more than one compaction batch per CTA (over 4096 samples per CTA)"),
    (70000, 12, 2, 0.02, "sparse weights, less than one sample per thread"),
])
def test_weighted_sum_paths_store_then_replay(oracle32, oracle64, k, tau, a, lam, what):
    """Phase 2 revisits only the samples whose fp32 weight is non-zero (0 * z adds nothing); which code path runs
    depends on lambda against the cost spread.  Every path must reproduce the oracle on the dumped noise."""
    rng = np.random.default_rng(k + tau)
    cfg = make_cfg(k, tau, 2 * a, a, lam=lam)
    x0 = rng.uniform(-1, 1, 2 * a).astype(np.float32)
    U0 = (0.2 * rng.standard_normal((tau, a))).astype(np.float32)
    ctrl = controller_from_cfg(cfg, seed=9)
    try:
        ctrl.setSequence(U0)
        act = ctrl.next(x0)
        U_new, costs, eps = ctrl.getUpdate(), ctrl.getCosts(), ctrl.dumpNoise()
        beta, eta = ctrl.getWeightStats()
    finally:
        ctrl.close()
    d = (costs.astype(np.float64) - costs.min()) / lam * 1.4426950408889634
    frac_nonzero = np.mean(d < 126)
    if lam < 1:
        assert frac_nonzero < 0.9, (what, frac_nonzero)          # the case really exercises the sparse path
    else:
        assert frac_nonzero == 1.0
    r64 = oracle64.mppi_update(cfg, x0, U0, eps)
    r32 = oracle32.mppi_update(cfg, x0, U0, eps)
    assert rel_err(costs, r64["costs"]) < 1e-5            # norm-wise: with lambda = 1e3 the action cost dominates
    assert_update_close(U_new, r64["U_new"], r32["U_new"], what="U_new: " + what)
    assert_update_close(act, r64["next"], r32["next"], what="next: " + what)
    assert abs(float(np.ravel(beta)[0]) - r64["costs"].min()) <= 2e-5 * abs(r64["costs"].min()) + 2e-5


def test_long_rows_and_the_shared_memory_limit(oracle64, oracle32):
    """tau * a_dim = 1800 still runs in Philox mode (store-then-replay against the checker); a row of 4000 floats does
    not fit the per-CTA tables and must fail with MPPI_ERR_UNSUPPORTED, not with a launch error."""
    from mppi_tf_b200 import MppiError, _capi
    k, tau, a = 256, 600, 3
    cfg = make_cfg(k, tau, 6, a, lam=5.0)
    rng = np.random.default_rng(8)
    x0 = rng.uniform(-1, 1, 6).astype(np.float32)
    U0 = (0.05 * rng.standard_normal((tau, a))).astype(np.float32)
    ctrl = controller_from_cfg(cfg, seed=3)
    try:
        ctrl.setSequence(U0)
        ctrl.next(x0)
        eps = ctrl.dumpNoise()
        ref = oracle64.mppi_update(cfg, x0, U0, eps)
        r32 = oracle32.mppi_update(cfg, x0, U0, eps)
        from tests.util import assert_update_close
        assert_update_close(ctrl.getUpdate(), ref["U_new"], r32["U_new"], what="T*a = 1800")
    finally:
        ctrl.close()
    big = controller_from_cfg(make_cfg(64, 1000, 8, 4))
    try:
        with pytest.raises(MppiError) as e:
            big.next(np.zeros(8, np.float32))
        assert e.value.code == _capi.MPPI_ERR_UNSUPPORTED
    finally:
        big.close()
