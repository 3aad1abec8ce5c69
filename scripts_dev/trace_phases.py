"""Dev helper: phase timeline of one update from the %globaltimer stamps (mppi_debug_trace).
usage: python scripts_dev/trace_phases.py K T a [n_ctrl] [rounds]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from mppi_tf_b200 import ControllerBase
K, T, a = (int(v) for v in sys.argv[1:4])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rounds = int(sys.argv[5]) if len(sys.argv) > 5 else 7
rng = np.random.default_rng(5)
goal = rng.uniform(-1, 1, (n, 2 * a)).astype(np.float32) if n > 1 else None
c = ControllerBase(K, T, 0.1, 1.0, 2 * a, a, sigma=0.25 * np.eye(a, dtype=np.float32), goal=goal, n_controllers=n,
                   goal_per_controller=n > 1, philox_rounds=rounds)
x = np.zeros((n, 2 * a), np.float32) if n == 1 else rng.uniform(-1, 1, (n, 2 * a)).astype(np.float32)
for _ in range(5):
    c.next(x)
c.debugTrace(True)
names = ["start", "tables", "rollout", "wsum", "published", "merged", "peers", "applied", "listed", "walked", "-", "-"]
for rep in range(3):
    c.debugTrace(True)            # clears the stamps of earlier updates
    c.next(x)
    tr = c.getTrace().astype(np.int64).reshape(-1, 12)
    t0 = tr[:, 0][tr[:, 0] > 0].min()
    print(f"-- update {rep}: grid {tr.shape[0]} CTAs; us after the first CTA's start: min / median / max over the CTAs that stamped")
    for i, nm in enumerate(names):
        v = tr[:, i][tr[:, i] > 0]
        if v.size:
            d = (v - t0) / 1e3
            print(f"   {nm:10s} n={v.size:5d}  {d.min():8.2f} {np.median(d):8.2f} {d.max():8.2f}")
c.close()
