"""world_size-2 (and 3) gloo test of the multi-rank protocol on CPU: the shard ranges and payload
layout come from the C-ABI's host helpers (mppi_shard_range / mppi_payload_stride), each rank
computes its (beta_r, eta_r, N_r) partial with the oracle, the payloads are all-gathered with
torch.distributed (gloo) and merged with log-sum-exp rescaling exactly as finish_kernel does on
the device; the result must equal the single-rank update."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, k, tau, a, lam, out):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from mppi_tf_b200 import _capi
    from oracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _capi.load()
    off, cnt = C.c_int(), C.c_int()
    assert lib.mppi_shard_range(k, rank, world, C.byref(off), C.byref(cnt)) == 0
    stride = lib.mppi_payload_stride(tau * a)
    orc = Oracle("f64")
    rng = np.random.default_rng(3)                     # same inputs on every rank
    costs = rng.uniform(0, 25, k)
    eps = rng.standard_normal((k, tau, a))
    U = 0.1 * rng.standard_normal((tau, a))
    beta, eta, N = orc.partial(lam, costs, eps, off.value, off.value + cnt.value)
    payload = torch.zeros(stride, dtype=torch.float64)     # {beta, eta, 0, 0, N[TA] padded}
    payload[0], payload[1] = beta, eta
    payload[4:4 + tau * a] = torch.from_numpy(N.reshape(-1))
    gathered = [torch.zeros(stride, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, payload)
    g = torch.stack(gathered).numpy()
    b = g[:, 0].min()
    scale = np.exp(-(g[:, 0] - b) / lam)
    eta_all = (scale * g[:, 1]).sum()
    N_all = (scale[:, None] * g[:, 4:4 + tau * a]).sum(0)
    U_new = U + (N_all / eta_all).reshape(tau, a)
    want = U + orc.update_stages(lam, costs, eps)["weighted_noise"]
    ranges = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(ranges, torch.tensor([off.value, cnt.value]))
    if rank == 0:
        r = torch.stack(ranges).numpy()
        ok_cover = (r[0, 0] == 0 and (r[:-1, 0] + r[:-1, 1] == r[1:, 0]).all() and r[-1, 0] + r[-1, 1] == k)
        out.put((float(np.abs(U_new - want).max()), bool(ok_cover), int(stride)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,k", [(2, 1000), (3, 1001)])
def test_sharded_merge_over_gloo(world, k):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    tau, a, lam = 7, 3, 0.6
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, tau, a, lam, out)) for r in range(world)]
    for p in procs:
        p.start()
    err, cover, stride = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert cover, "shards must tile [0, k) without gaps"
    assert stride == 4 + ((tau * a + 3) // 4) * 4
    assert err < 1e-12


def test_shard_range_rejects_bad_arguments():
    import ctypes as C
    from mppi_tf_b200 import _capi
    lib = _capi.load()
    off, cnt = C.c_int(), C.c_int()
    assert lib.mppi_shard_range(10, 2, 2, C.byref(off), C.byref(cnt)) == _capi.MPPI_ERR_BAD_ARG
    assert lib.mppi_shard_range(1, 0, 2, C.byref(off), C.byref(cnt)) == _capi.MPPI_ERR_BAD_ARG
    assert lib.mppi_shard_range(1048576, 7, 8, C.byref(off), C.byref(cnt)) == 0
    assert (off.value, cnt.value) == (917504, 131072)
