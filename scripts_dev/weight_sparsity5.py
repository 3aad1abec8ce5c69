import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mppi_tf_b200 import ControllerBase
K, T, s, a, n = 1024, 30, 4, 2, 4096
rng = np.random.default_rng(5)
goal = rng.uniform(-1, 1, (n, s)).astype(np.float32)
x = rng.uniform(-1, 1, (n, s)).astype(np.float32)
c = ControllerBase(K, T, 0.1, 1.0, s, a, lam=1.0, sigma=0.25 * np.eye(a, dtype=np.float32), goal=goal, seed=1, n_controllers=n, goal_per_controller=True)
for it in range(12):
    c.next(x)
    if it in (0, 3, 11):
        S = c.getCosts().astype(np.float64)
        d = S - S.min(1, keepdims=True)
        nz = d <= 87.3
        print(f"update {it}: nonzero frac {nz.mean():.3f}; per-warp cnt mean {nz.reshape(n, 16, 2, 32).sum((2,3)).mean():.1f}; warps with cnt<=32: {(nz.reshape(n,16,2,32).sum((2,3))<=32).mean():.3f}")
c.close()
