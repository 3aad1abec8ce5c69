"""Fused exchange over peer memory (mppi_peer_handle / mppi_peer_attach): needs two GPUs of one box, one process
each.  Skipped on single-GPU boxes; `scripts_dev/scale.sh` runs the same worker on 8 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:                      # noqa: BLE001
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_peer_exchange_matches_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "peer_check_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PEER CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_peer_handle_needs_a_sharded_handle():
    from mppi_tf_b200 import ControllerBase, MppiError, _capi
    c = ControllerBase(1024, 8, 0.1, 1.0, 2, 1)
    try:
        with pytest.raises(MppiError) as e:
            c.peerHandle()
        assert e.value.code == _capi.MPPI_ERR_UNSUPPORTED
    finally:
        c.close()
    c = ControllerBase(1024, 8, 0.1, 1.0, 2, 1, rank=1, world=2)
    try:
        assert len(c.peerHandle()) == 64
    finally:
        c.close()
