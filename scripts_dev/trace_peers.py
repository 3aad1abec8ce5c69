"""Dev helper: phase stamps of a sharded update with the fused peer exchange (per rank, own clock).
python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts_dev/trace_peers.py [K]"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from mppi_tf_b200 import ControllerBase

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
T, s, a = 100, 6, 3
c = ControllerBase(K, T, 0.1, 1.0, s, a, sigma=0.25 * np.eye(a, dtype=np.float32), device=lr, rank=rank, world=world, philox_rounds=7)
hs = [None] * world
dist.all_gather_object(hs, c.peerHandle())
c.peerAttach(hs)
x = np.zeros(s, np.float32)
for _ in range(10):
    c.next(x)
names = ["start", "tables", "rollout", "wsum", "published", "merged", "peers", "applied"]
rows = []
for rep in range(6):
    c.debugTrace(True)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    c.next(x)
    tr = c.getTrace().astype(np.int64).reshape(-1, 12)
    t0 = tr[:, 0][tr[:, 0] > 0].min()
    last = {nm: (tr[:, i][tr[:, i] > 0].max() - t0) / 1e3 for i, nm in enumerate(names) if (tr[:, i] > 0).any()}
    rows.append(last)
out = [None] * world
dist.all_gather_object(out, rows)
if rank == 0:
    for rep in range(6):
        print(f"-- update {rep} (us after the rank's own first CTA start; last stamp of each phase)")
        for r in range(world):
            d = out[r][rep]
            print(f"   rank {r}: " + "  ".join(f"{k} {v:7.2f}" for k, v in d.items()) +
                  f"   | exchange {d.get('peers', 0) - d.get('merged', 0):6.2f}  apply {d.get('applied', 0) - d.get('peers', 0):5.2f}")
c.close()
dist.destroy_process_group()
