"""Dev: small end-to-end updates of every kernel family for compute-sanitizer (memcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mppi_tf_b200 import ControllerBase
from bench import glorot_mlp
rng = np.random.default_rng(0)
for (k, tau, a, n) in [(1500, 20, 1, 1), (3000, 33, 3, 1), (4096, 100, 3, 1), (1024, 30, 2, 7), (2048, 12, 4, 1)]:
    s = 2 * a
    goal = rng.uniform(-1, 1, (n, s)).astype(np.float32) if n > 1 else None
    c = ControllerBase(k, tau, 0.1, 1.0, s, a, lam=1.0, sigma=0.25 * np.eye(a, dtype=np.float32), goal=goal, n_controllers=n,
                       goal_per_controller=n > 1)
    x = rng.uniform(-1, 1, (n, s)).astype(np.float32) if n > 1 else rng.uniform(-1, 1, s).astype(np.float32)
    eps = (0.25 * rng.standard_normal((n, k, tau, a))).astype(np.float32)
    for lam in (1.0, 50.0):                     # sparse and dense weights
        c.setLambda(lam)
        c.next(x); c.next(x)
        c.nextWithNoise(x, eps if n > 1 else eps[0])
    c.setActionCost("python", gamma=0.5, upsilon=1.5)
    c.setNormalizeCost(True)
    c.next(x); c.nextWithNoise(x, eps if n > 1 else eps[0])
    if a == 2 and n == 1:
        c.setEllipseCost(1.2, 0.8, 0, 0, 1, 1, 1)
        c.next(x)
    c.close()
    print("ok", k, tau, a, n, flush=True)
k, tau, a = 1000, 12, 3
c = ControllerBase(k, tau, 0.1, 1.0, 6, 3, sigma=0.25 * np.eye(3, dtype=np.float32))
c.setMlp(glorot_mlp(6, 3))
x = np.zeros(6, np.float32)
c.next(x); c.nextWithNoise(x, (0.25 * rng.standard_normal((k, tau, a))).astype(np.float32))
c.mlpPredict(rng.uniform(-1, 1, (300, 6)), rng.uniform(-1, 1, (300, 3)))
c.close()
print("ok mlp", flush=True)
