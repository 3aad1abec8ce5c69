import sys
import numpy as np
sys.path.insert(0, ".")
from mppi_tf_b200 import ControllerBase
K, T, a = 131072, 100, 3
c = ControllerBase(K, T, 0.1, 1.0, 2 * a, a, sigma=0.25 * np.eye(a, dtype=np.float32), philox_rounds=7)
x = np.zeros((1, 2 * a), np.float32)
for _ in range(5):
    c.next(x)
for rep in range(4):
    c.debugTrace(True)
    c.next(x)
    tr = c.getTrace().astype(np.int64).reshape(-1, 12)
    f = int(np.argmax(tr[:, 5]))
    t4 = tr[f, 4]
    print(f"finisher CTA {f}: published(own) 0.00, group-elected {(tr[f,8]-t4)/1e3:.2f}, group-merged {(tr[f,9]-t4)/1e3:.2f}, top-elected {(tr[f,3]-t4)/1e3:.2f}, merged {(tr[f,5]-t4)/1e3:.2f}, applied {(tr[f,7]-t4)/1e3:.2f}; last published overall {(tr[:,4].max()-t4)/1e3:.2f}")
c.close()
