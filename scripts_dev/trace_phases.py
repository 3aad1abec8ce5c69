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
names = ["start", "tables", "rollout", "wsum", "published", "merged", "peers", "applied", "listed", "walked"]
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
    if rep == 2 and tr[:, 10].max() > 0 and n == 1:
        # placement: CTAs per SM, and the rollout end of the first / second CTA of every SM
        sm = tr[:, 10] - 1
        slot = tr[:, 11] - 1
        ro = (tr[:, 2] - t0) / 1e3
        per = {}
        for i in range(tr.shape[0]):
            per.setdefault(int(sm[i]), []).append((float(ro[i]), int(slot[i]), i))
        cnt = np.bincount([len(v) for v in per.values()])
        print("   CTAs per SM histogram:", cnt.tolist(), " SMs used:", len(per))
        firsts = [sorted(v)[0][0] for v in per.values() if len(v) >= 2]
        seconds = [sorted(v)[1][0] for v in per.values() if len(v) >= 2]
        if firsts:
            print(f"   first CTA of an SM ends rollout at  {min(firsts):.1f} / {np.median(firsts):.1f} / {max(firsts):.1f}")
            print(f"   second CTA of an SM ends rollout at {min(seconds):.1f} / {np.median(seconds):.1f} / {max(seconds):.1f}")
        print("   warp slot of thread 0:", sorted(set(slot.tolist()))[:8], " rollout-end deciles:", np.percentile(ro, [0, 10, 25, 50, 75, 90, 100]).round(1).tolist())
c.close()
