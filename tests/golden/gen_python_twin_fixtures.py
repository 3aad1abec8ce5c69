"""Generates tests/golden/python_twin_fixtures.npz by running the REFERENCE'S OWN Python twin
(/root/reference/scripts/src: ControllerBase.build_model / update / get_next / shift,
PointMassModel, StaticCost) on fixed inputs with the noise tensor injected.

TensorFlow, cpprb, matplotlib and imageio are not installable here, so the reference modules are
imported against tests/golden/tf_shim (a numpy stand-in for the ~25 TF ops they call).  What this
pins is the reference's op SEQUENCE and constants for a composed update — rollout -> cost -> softmin
update -> next / shift — which the reference's own tests never check end to end
(test/test_controller.cpp:224-226 `testAll` is empty).

One repair is needed to import HEAD at all: PointMassModel.__init__ calls self.add_model_vars
(scripts/src/models/point_mass_model.py:61), which ModelBase no longer defines; it is restored as
the one-line dict insert the name implies.

Run here (never on the GPU box; /root/reference does not exist there):
    python tests/golden/gen_python_twin_fixtures.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(1, "/root/reference")

from scripts.src.models import model_base as ref_model_base          # noqa: E402
ref_model_base.ModelBase.add_model_vars = lambda self, name, var: self._modelVars.__setitem__(name, var)
from scripts.src.models.point_mass_model import PointMassModel       # noqa: E402
from scripts.src.costs.static_cost import StaticCost                 # noqa: E402
from scripts.src.costs.elipse_cost import ElipseCost                 # noqa: E402
from scripts.src.controllers.controller_base import ControllerBase   # noqa: E402

CASES = [
    dict(name="pm1d", k=64, tau=8, s=2, a=1, mass=1.0, dt=0.1, lam=1.0, full_sigma=False),
    dict(name="pm2d", k=96, tau=10, s=4, a=2, mass=2.0, dt=0.05, lam=0.7, full_sigma=True),
    dict(name="pm3d", k=128, tau=12, s=6, a=3, mass=1.5, dt=0.1, lam=2.0, full_sigma=True),
    # Python-twin extras: gamma != lambda, upsilon != 1, cost normalisation
    dict(name="tw1d", k=64, tau=8, s=2, a=1, mass=1.0, dt=0.1, lam=1.3, full_sigma=False, gamma=0.4, upsilon=1.7,
         normalize=False),
    dict(name="tw2d", k=96, tau=10, s=4, a=2, mass=2.0, dt=0.05, lam=0.7, full_sigma=True, gamma=0.7, upsilon=1.0,
         normalize=True),
    dict(name="tw3d", k=128, tau=12, s=6, a=3, mass=1.5, dt=0.1, lam=2.0, full_sigma=True, gamma=0.9, upsilon=2.5,
         normalize=True),
    # ElipseCost as the state cost, point_mass2d
    dict(name="el2d", k=128, tau=14, s=4, a=2, mass=1.0, dt=0.1, lam=0.6, full_sigma=True, gamma=0.6, upsilon=1.0,
         normalize=False, ellipse=(1.5, 0.8, 0.2, -0.1, 0.7, 2.0, 0.5)),
    dict(name="el2n", k=96, tau=11, s=4, a=2, mass=2.0, dt=0.05, lam=1.1, full_sigma=False, gamma=0.3, upsilon=1.4,
         normalize=True, ellipse=(1.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0)),
    dict(name="tw3e", k=160, tau=9, s=6, a=3, mass=1.0, dt=0.1, lam=0.5, full_sigma=False, gamma=1.5, upsilon=0.6,
         normalize=False),
]


def run_case(c, seed):
    rng = np.random.default_rng(seed)
    k, tau, s, a = c["k"], c["tau"], c["s"], c["a"]
    sigma = 0.25 * np.eye(a)
    if c["full_sigma"]:
        L = 0.2 * rng.standard_normal((a, a))
        sigma = L @ L.T + 0.2 * np.eye(a)
    goal = rng.uniform(-1, 1, (s, 1))
    q = 1.0 + 4.0 * rng.random(s)
    x = rng.uniform(-1, 1, (s, 1))
    U = 0.2 * rng.standard_normal((tau, a, 1))
    z = rng.standard_normal((k, tau, a, 1))
    gamma, upsilon, normalize = c.get("gamma", c["lam"]), c.get("upsilon", 1.0), c.get("normalize", False)
    eps = np.matmul(upsilon * sigma, z)                         # build_noise: (upsilon * sigma) @ rng (:368)

    model = PointMassModel(None, mass=c["mass"], dt=c["dt"], stateDim=s, actionDim=a)
    if "ellipse" in c:
        ea, eb, ecx, ecy, espeed, ems, emv = c["ellipse"]
        cost = ElipseCost(c["lam"], gamma, upsilon, sigma, ea, eb, ecx, ecy, espeed, ems, emv)
    else:
        cost = StaticCost(c["lam"], gamma, upsilon, sigma, goal, np.diag(q))
    ctrl = ControllerBase(model, cost, k=k, tau=tau, sDim=s, aDim=a, lam=c["lam"], upsilon=upsilon, sigma=sigma,
                          initSeq=U.copy())
    costs = ctrl.build_model("rollout", k, x, eps, U)           # [k,1,1]
    update = ctrl.update("update", costs, eps, normalize=normalize)   # uses ctrl._actionSeq = U
    nxt = ctrl.get_next("next", update, 1)
    shifted = ctrl.shift("shift", update, ctrl.init_zeros("init", 1), 1)
    p = c["name"] + "_"
    return {p + "sigma": sigma, p + "goal": goal[:, 0], p + "q": q, p + "x": x[:, 0], p + "U": U[..., 0],
            p + "eps": eps[..., 0], p + "costs_py": np.asarray(costs).reshape(k),
            p + "U_new": np.asarray(update)[..., 0], p + "next": np.asarray(nxt).reshape(a),
            p + "U_shift": np.asarray(shifted)[..., 0],
            p + "meta": np.array([k, tau, s, a, c["mass"], c["dt"], c["lam"]], np.float64),
            p + "twin": np.array([gamma, upsilon, float(normalize)], np.float64),
            p + "ellipse": np.array(c.get("ellipse", ()), np.float64)}


def main():
    out = {}
    for i, c in enumerate(CASES):
        out.update(run_case(c, 100 + i))
    path = os.path.join(HERE, "python_twin_fixtures.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
