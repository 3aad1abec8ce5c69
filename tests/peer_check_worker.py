"""Worker of tests/test_peer_exchange_gpu.py (needs >= 2 GPUs): the fused peer-memory exchange must give the same
sequences as the NCCL path (to the last bits of an fp32 sum), bit-identical on every rank.
python -m torch.distributed.run --nproc-per-node N tests/peer_check_worker.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mppi_tf_b200 import ControllerBase, comm_unique_id

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
K, T, s, a = 65536, 40, 6, 3
x = np.linspace(-0.5, 0.5, s).astype(np.float32)


def make():
    return ControllerBase(K, T, 0.1, 1.0, s, a, lam=1.0, sigma=0.25 * np.eye(a, dtype=np.float32), seed=3, device=lr,
                          rank=rank, world=world)


c1 = make()
uid = [comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
c1.commInit(uid[0])
c2 = make()
hs = [None] * world
dist.all_gather_object(hs, c2.peerHandle())
c2.peerAttach(hs)
ok = True
for it in range(5):
    a1 = c1.next(x)
    a2 = c2.next(x)
    u1, u2 = c1.getSequence(), c2.getSequence()
    # the two transports merge the same rank payloads, but in different kernels (finish_kernel / the update kernel's last CTA)
    # whose merge slices the records by their own CTA size: same sum, last-bit differences in its association
    scale = max(np.abs(u1).max(), 1e-30)
    err = max(np.abs(u1 - u2).max(), np.abs(a1 - a2).max()) / scale
    same = err <= 2e-6
    ok &= bool(same)
    if rank == 0:
        print(f"update {it}: nccl vs peer rel diff {err:.1e} (identical: {np.array_equal(u1, u2)}); action {a2}", flush=True)
# all ranks hold the same sequence
t = torch.tensor(c2.getSequence(), device="cuda")
g = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(g, t)
ok &= all(torch.equal(g[0], gi) for gi in g)
# cost normalisation over the mailboxes (two exchanges per update): the sharded result must agree with an
# unsharded controller on the same seed (Philox noise does not depend on the rank count)
c3 = make()
hs = [None] * world
dist.all_gather_object(hs, c3.peerHandle())
c3.peerAttach(hs)
c3.setActionCost("python", gamma=0.5, upsilon=1.2)
c3.setNormalizeCost(True)
c0 = None
if rank == 0:
    c0 = ControllerBase(K, T, 0.1, 1.0, s, a, lam=1.0, sigma=0.25 * np.eye(a, dtype=np.float32), seed=3, device=lr)
    c0.setActionCost("python", gamma=0.5, upsilon=1.2)
    c0.setNormalizeCost(True)
for it in range(3):
    a3 = c3.next(x)
    if rank == 0:
        a0 = c0.next(x)
        u3, u0 = c3.getSequence(), c0.getSequence()
        err = np.abs(u3 - u0).max() / max(np.abs(u0).max(), 1e-30)
        same = err < 5e-5 and np.abs(a3 - a0).max() <= 5e-5 * max(np.abs(u0).max(), 1e-30)
        ok &= bool(same)
        print(f"normalised update {it}: sharded vs unsharded rel err {err:.2e}", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, src=0)
ok = bool(flag.item())
c3.close()
if c0 is not None:
    c0.close()
if rank == 0:
    print("PEER CHECK", "OK" if ok else "FAILED", flush=True)
c1.close(); c2.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
