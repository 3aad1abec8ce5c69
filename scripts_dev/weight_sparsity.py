"""Dev: how many samples of the bench workload carry a weight that is exactly zero in fp32?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mppi_tf_b200 import ControllerBase
K, T, s, a = 1048576, 100, 6, 3
c = ControllerBase(K, T, 0.1, 1.0, s, a, lam=1.0, sigma=0.25 * np.eye(a, dtype=np.float32), seed=1)
x = np.zeros(s, np.float32)
for it in range(25):
    act = c.next(x)
    if it in (0, 1, 2, 5, 10, 24):
        S = c.getCosts().astype(np.float64)
        d = S - S.min()
        print(f"update {it}: beta {S.min():.2f} median S-beta {np.median(d):.2f}  frac(e==0 ftz: >87.3) {np.mean(d > 87.3):.4f} "
              f" frac(>103.3) {np.mean(d > 103.3):.4f}  frac(e<1e-7) {np.mean(d > 16.1):.4f}  warps all-zero {np.mean((d.reshape(-1, 32) > 87.3).all(1)):.4f}")
c.close()
