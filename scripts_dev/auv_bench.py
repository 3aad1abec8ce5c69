"""Developer timing of the AUV rollout kernel (not a bench line): K x T update in Philox and injected mode.
  MPPI_AUV_GEOM=0..3 python scripts_dev/auv_bench.py [K] [T] [rk]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mppi_tf_b200 import ControllerBase  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
T = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rk = int(sys.argv[3]) if len(sys.argv) > 3 else 2
d = np.load(os.path.join(ROOT, "tests", "golden", "auv_fixtures.npz"))
prm = json.loads(bytes(d["params_json"]).decode())["full"]
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
sigma = 80.0 * np.eye(6, dtype=np.float32)
ctrl = ControllerBase(K, T, 0.1, 1.0, 13, 6, lam=1.0, sigma=sigma, model="auv", stream=st.cuda_stream)
ctrl.setAuvModel(prm, rk=rk)
ctrl.setActionCost("python", gamma=1.0, upsilon=1.0)
x = np.zeros(13, np.float32); x[6] = 1.0; x[0] = 1.0
ctrl.setState(x)


def timeit(ptr=None, n=20):
    for _ in range(3):
        ctrl.enqueueUpdate(ptr)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); ctrl.enqueueUpdate(ptr); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return ms[len(ms) // 2]


ms = timeit()
print(f"geom={os.environ.get('MPPI_AUV_GEOM', 'default')} K={K} T={T} rk={rk} philox: {ms:.4f} ms  {K * T / ms * 1e3:.3e} sample-steps/s")
eps = torch.randn(K * T * 6, device=dev) * 80.0
ms = timeit(eps.data_ptr())
print(f"   injected: {ms:.4f} ms  {K * T / ms * 1e3:.3e} sample-steps/s  ({K * T * 24 / ms / 1e6:.0f} GB/s of eps)")
ctrl.close()
