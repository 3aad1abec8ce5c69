"""Dev: run one cfg4-like MLP update with the trace build and print the hand-over timeline of CTA 0."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MPPI_B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mppi_tf_b200/_build_alt/libmppi_b200.so")
from bench import glorot_mlp
from mppi_tf_b200 import ControllerBase, _capi
K, T, s, a = 262144, 50, 6, 3
c = ControllerBase(K, T, 0.1, 1.0, s, a, sigma=0.25 * np.eye(a, dtype=np.float32), device=0)
c.setMlp(glorot_mlp(s, a))
x = np.zeros(s, np.float32)
lib = C.CDLL(os.environ["MPPI_B200_LIB"])
out = (C.c_longlong * 2048)(); n = (C.c_int * 2)()
for it in range(3):
    c.next(x)
    lib.mppi_debug_mlp_trace(out, n)
ev = []
for who in range(2):
    for i in range(n[who]):
        v = out[who * 1024 + i]
        ev.append((v >> 8, who, v & 255))
ev.sort()
t0 = ev[0][0]
RN = {1: "X arrived(sent)", 2: "wait D0(L1)", 3: "got D0(L1)", 4: "packed h0", 5: "A0 sent", 6: "got D1(L1)", 7: "A1 sent",
      8: "wait D0(L2)", 9: "got D0(L2)", 10: "packed h0", 11: "got D1(L2)", 12: "A0 sent", 13: "A1 sent", 14: "wait D3", 15: "got D3"}
MN = {0: "wait X", 1: "got X", 2: "L1 issued; wait A0", 3: "got A0", 4: "L2 k0-3 issued; wait A1", 5: "got A1",
      6: "L2 issued; wait A0", 7: "got A0", 8: "L3 k0-3 issued; wait A1", 9: "got A1", 10: "L3 issued"}
last = {0: t0, 1: t0}
for t, who, e in ev[:int(sys.argv[1]) if len(sys.argv) > 1 else 160]:
    print(f"{t - t0:8d}  +{t - last[who]:5d}  {'ROW' if who == 0 else '        MMA'}  {(RN if who == 0 else MN)[e]}")
    last[who] = t
c.close()
