"""Generates tests/golden/auv_update_fixtures.npz by running the REFERENCE'S OWN Python controller
(/root/reference/scripts/src/controllers/controller_base.py: build_model / update / get_next / shift) with the
reference's AUVModel (scripts/src/models/auv_model.py) and StaticCost / StaticQuatCost
(scripts/src/costs/static_cost.py) on fixed inputs with the noise tensor injected, against tests/golden/tf_shim.
Also stores StaticQuatCost.state_cost on random states.  Run here only (never on the GPU box):
    python tests/golden/gen_auv_update_fixtures.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, HERE)

from scripts.src.models.auv_model import AUVModel                    # noqa: E402
from scripts.src.costs.static_cost import StaticCost, StaticQuatCost  # noqa: E402
from scripts.src.controllers.controller_base import ControllerBase   # noqa: E402
from gen_auv_fixtures import params                                  # noqa: E402

CASES = [
    dict(name="auv_rk1", prm="test", k=96, tau=8, rk=1, lam=0.8, gamma=0.8, upsilon=1.0, normalize=False, quat=False),
    dict(name="auv_rk2", prm="full", k=128, tau=10, rk=2, lam=1.2, gamma=0.5, upsilon=1.6, normalize=False, quat=False),
    dict(name="auv_quat", prm="full", k=128, tau=9, rk=2, lam=0.6, gamma=0.6, upsilon=1.0, normalize=False, quat=True),
    dict(name="auv_quatn", prm="test", k=96, tau=7, rk=1, lam=1.5, gamma=0.9, upsilon=0.7, normalize=True, quat=True),
    dict(name="auv_rk4", prm="test", k=64, tau=6, rk=4, lam=1.0, gamma=1.0, upsilon=1.0, normalize=False, quat=False),
]


def run_case(c, seed):
    rng = np.random.default_rng(seed)
    k, tau, s, a = c["k"], c["tau"], 13, 6
    L = 4.0 * rng.standard_normal((a, a))
    sigma = L @ L.T + 60.0 * np.eye(a)                            # forces of tens of newtons
    goal = rng.uniform(-1, 1, (s, 1))
    goal[3:7] /= np.linalg.norm(goal[3:7])
    x = rng.uniform(-0.5, 0.5, (s, 1))
    x[3:7] = goal[3:7] + 0.3 * rng.standard_normal((4, 1))        # within reach of the goal attitude
    x[3:7] /= np.linalg.norm(x[3:7])
    U = 20.0 * rng.standard_normal((tau, a, 1))
    z = rng.standard_normal((k, tau, a, 1))
    eps = np.matmul(c["upsilon"] * sigma, z)                      # build_noise (controller_base.py:368)
    prm = params(c["prm"])
    model = AUVModel({}, actionDim=a, dt=0.1, parameters=dict(prm, rk=c["rk"]))
    model._k = k
    if c["quat"]:
        q = 1.0 + 4.0 * rng.random(10)
        cost = StaticQuatCost(c["lam"], c["gamma"], c["upsilon"], sigma, goal, q, diag=True)
    else:
        q = 1.0 + 4.0 * rng.random(13)
        cost = StaticCost(c["lam"], c["gamma"], c["upsilon"], sigma, goal, np.diag(q))
    ctrl = ControllerBase(model, cost, k=k, tau=tau, sDim=s, aDim=a, lam=c["lam"], upsilon=c["upsilon"], sigma=sigma,
                          initSeq=U.copy())
    costs = ctrl.build_model("rollout", k, x, eps, U)
    update = ctrl.update("update", costs, eps, normalize=c["normalize"])
    nxt = ctrl.get_next("next", update, 1)
    shifted = ctrl.shift("shift", update, ctrl.init_zeros("init", 1), 1)
    p = c["name"] + "_"
    out = {p + "sigma": sigma, p + "goal": goal[:, 0], p + "q": q, p + "x": x[:, 0], p + "U": U[..., 0],
           p + "eps": eps[..., 0], p + "costs_py": np.asarray(costs).reshape(k), p + "U_new": np.asarray(update)[..., 0],
           p + "next": np.asarray(nxt).reshape(a), p + "U_shift": np.asarray(shifted)[..., 0],
           p + "meta": np.array([k, tau, c["rk"], c["lam"], c["gamma"], c["upsilon"], float(c["normalize"]), float(c["quat"])])}
    if c["quat"]:                                                 # the cost functor alone, on random unit quaternions
        st = rng.uniform(-1, 1, (40, 13, 1))
        st[:, 3:7] /= np.linalg.norm(st[:, 3:7], axis=1, keepdims=True)
        out[p + "qc_state"] = st[..., 0]
        out[p + "qc_cost"] = np.asarray(cost.state_cost("c", st)).reshape(40)
    return out, c["prm"]


def main():
    store, which = {}, {}
    for i, c in enumerate(CASES):
        out, w = run_case(c, 500 + i)
        store.update(out)
        which[c["name"]] = w
    store["params_json"] = np.frombuffer(json.dumps({"which": which, "prm": {w: params(w) for w in ("test", "full")}}).encode(),
                                         dtype=np.uint8)
    path = os.path.join(HERE, "auv_update_fixtures.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
