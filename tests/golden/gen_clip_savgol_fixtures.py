"""Generates tests/golden/clip_savgol_fixtures.npz: the two remaining Python-twin extras of SURVEY.md section 8(f) row N1.

  clip_act   the REFERENCE'S OWN ControllerBase.build_model / update / clip_act / get_next / shift
             (/root/reference/scripts/src/controllers/controller_base.py:436-462,500-504 with the limits of
             models/model_base.py:121-126) on the numpy TF shim, noise injected.  At the reference's HEAD the call to
             clip_act inside update() is commented out (:455-456); the fixture applies it where that line stands.
  savgol     the exact call of controller_base.py:281-291, scipy.signal.savgol_filter(seq, 10, 9, deriv=0, delta=1.0,
             axis=0), on the updated sequences (scipy is the reference's third-party dependency for this pass).

Run here (never on the GPU box; /root/reference does not exist there):
    python tests/golden/gen_clip_savgol_fixtures.py
"""
import os
import sys

import numpy as np
import scipy.signal

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(1, "/root/reference")

from scripts.src.models import model_base as ref_model_base          # noqa: E402
ref_model_base.ModelBase.add_model_vars = lambda self, name, var: self._modelVars.__setitem__(name, var)
from scripts.src.models.point_mass_model import PointMassModel       # noqa: E402
from scripts.src.costs.static_cost import StaticCost                 # noqa: E402
from scripts.src.controllers.controller_base import ControllerBase   # noqa: E402

CASES = [
    dict(name="clip1d", k=64, tau=12, s=2, a=1, mass=1.0, dt=0.1, lam=1.0, lim_min=[-0.15], lim_max=[0.2]),
    dict(name="clip2d", k=96, tau=16, s=4, a=2, mass=2.0, dt=0.05, lam=0.7, lim_min=[-0.3], lim_max=[0.1]),
    dict(name="clip3d", k=128, tau=20, s=6, a=3, mass=1.5, dt=0.1, lam=2.0, lim_min=[-0.2, -0.05, -0.4], lim_max=[0.25, 0.3, 0.02]),
]


def run_case(c, seed):
    rng = np.random.default_rng(seed)
    k, tau, s, a = c["k"], c["tau"], c["s"], c["a"]
    sigma = 0.25 * np.eye(a)
    goal = rng.uniform(-1, 1, (s, 1))
    q = 1.0 + 4.0 * rng.random(s)
    x = rng.uniform(-1, 1, (s, 1))
    U = 0.2 * rng.standard_normal((tau, a, 1))
    eps = np.matmul(sigma, rng.standard_normal((k, tau, a, 1)))
    lo = np.asarray(c["lim_min"], np.float64).reshape(-1, 1)
    hi = np.asarray(c["lim_max"], np.float64).reshape(-1, 1)
    model = PointMassModel(None, mass=c["mass"], dt=c["dt"], stateDim=s, actionDim=a)
    # PointMassModel drops its limMax / limMin arguments (point_mass_model.py:55 does not forward them, so max_act / min_act
    # answer the ModelBase default of +-1): the limits are set where ModelBase keeps them (model_base.py:32-33)
    model._actMax, model._actMin = hi, lo
    cost = StaticCost(c["lam"], c["lam"], 1.0, sigma, goal, np.diag(q))
    ctrl = ControllerBase(model, cost, k=k, tau=tau, sDim=s, aDim=a, lam=c["lam"], upsilon=1.0, sigma=sigma, initSeq=U.copy())
    costs = ctrl.build_model("rollout", k, x, eps, U)
    raw = ctrl.update("update", costs, eps, normalize=False)
    clipped = ctrl.clip_act("clipping", raw)                     # the line commented out at :455
    nxt = ctrl.get_next("next", clipped, 1)
    shifted = ctrl.shift("shift", clipped, ctrl.init_zeros("init", 1), 1)
    sg = scipy.signal.savgol_filter(np.asarray(clipped)[:, :, 0], 10, 9, deriv=0, delta=1.0, axis=0)     # :284-291
    p = c["name"] + "_"
    return {p + "sigma": sigma, p + "goal": goal[:, 0], p + "q": q, p + "x": x[:, 0], p + "U": U[..., 0], p + "eps": eps[..., 0],
            p + "lim_min": lo[:, 0], p + "lim_max": hi[:, 0], p + "U_raw": np.asarray(raw)[..., 0],
            p + "U_new": np.asarray(clipped)[..., 0], p + "next": np.asarray(nxt).reshape(a),
            p + "U_shift": np.asarray(shifted)[..., 0], p + "savgol_10_9": sg,
            p + "meta": np.array([k, tau, s, a, c["mass"], c["dt"], c["lam"]], np.float64)}


def main():
    out = {}
    for i, c in enumerate(CASES):
        out.update(run_case(c, 300 + i))
    path = os.path.join(HERE, "clip_savgol_fixtures.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
