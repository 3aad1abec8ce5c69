"""Generates tests/golden/nn_auv_fixtures.npz by running the REFERENCE'S OWN learned AUV model, NNAUVModel
(/root/reference/scripts/src/models/nn_model.py:181-304: prepare_data -> Sequential(Dense 32 relu x 3, Dense 13) ->
denormalizeY -> next_state = state + delta), and its ControllerBase.build_model / update / get_next / shift with that
model and StaticCost, on the numpy TF shim (tests/golden/tf_shim; tf.keras there is a forward-only Dense / Sequential).
Weights and normalisation constants are drawn here and stored with the outputs.

One repair is needed to construct the class at HEAD: NNAUVModel.__init__ passes `limMax=limMax, limMin=limMin` to its
base class without defining them (nn_model.py:203-206, a NameError); the two names are provided as module globals with
the values every other model defaults to (+-1).

Run here only (never on the GPU box):  python tests/golden/gen_nn_auv_fixtures.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(1, "/root/reference")

from scripts.src.models import nn_model as ref_nn                      # noqa: E402
ref_nn.limMax, ref_nn.limMin = np.ones(1), -np.ones(1)
from scripts.src.costs.static_cost import StaticCost                   # noqa: E402
from scripts.src.controllers.controller_base import ControllerBase     # noqa: E402


def make_model(rng, scale):
    m = ref_nn.NNAUVModel(None)
    for layer in m.nn.layers:
        n_in, n_out = layer.kernel.shape
        lim = np.sqrt(6.0 / (n_in + n_out)) * scale
        layer.kernel = rng.uniform(-lim, lim, (n_in, n_out))
        layer.bias = 0.05 * rng.standard_normal(n_out)
    m.set_Xmean_Xstd(0.1 * rng.standard_normal(16), 1.0 + rng.random(16))
    m.set_Ymean_Ystd(0.002 * rng.standard_normal(13), 0.01 + 0.02 * rng.random(13))
    return m


def main():
    rng = np.random.default_rng(77)
    out = {}
    model = make_model(rng, 0.8)
    for i, layer in enumerate(model.nn.layers):
        out[f"W{i}"], out[f"b{i}"] = layer.kernel, layer.bias
    out["Xmean"], out["Xstd"], out["Ymean"], out["Ystd"] = (np.asarray(v) for v in (model.Xmean, model.Xstd, model.Ymean, model.Ystd))
    # one-step predictions on a batch (build_step_graph)
    k = 64
    st = rng.uniform(-1, 1, (k, 13, 1))
    st[:, 3:7] /= np.linalg.norm(st[:, 3:7], axis=1, keepdims=True)
    ac = 30.0 * rng.standard_normal((k, 6, 1))
    out["step_state"], out["step_action"] = st[..., 0], ac[..., 0]
    out["step_next"] = np.asarray(model.build_step_graph("step", st, ac))[..., 0]
    # full controller updates with the learned model
    for name, kk, tau, lam, gamma, upsilon, norm in (("upd1", 96, 10, 1.2, 1.2, 1.0, False), ("upd2", 128, 14, 0.8, 0.5, 1.0, True)):
        sigma = np.diag(10.0 + 20.0 * rng.random(6))
        goal = rng.uniform(-1, 1, (13, 1))
        goal[3:7] /= np.linalg.norm(goal[3:7])
        q = 1.0 + 4.0 * rng.random(13)
        x = goal + 0.3 * rng.standard_normal((13, 1))
        x[3:7] /= np.linalg.norm(x[3:7])
        U = 5.0 * rng.standard_normal((tau, 6, 1))
        eps = np.matmul(upsilon * sigma, rng.standard_normal((kk, tau, 6, 1)))
        cost = StaticCost(lam, gamma, upsilon, sigma, goal, np.diag(q))
        ctrl = ControllerBase(model, cost, k=kk, tau=tau, sDim=13, aDim=6, lam=lam, upsilon=upsilon, sigma=sigma, initSeq=U.copy())
        costs = ctrl.build_model("rollout", kk, x, eps, U)
        update = ctrl.update("update", costs, eps, normalize=norm)
        nxt = ctrl.get_next("next", update, 1)
        shifted = ctrl.shift("shift", update, ctrl.init_zeros("init", 1), 1)
        p = name + "_"
        out.update({p + "sigma": sigma, p + "goal": goal[:, 0], p + "q": q, p + "x": x[:, 0], p + "U": U[..., 0], p + "eps": eps[..., 0],
                    p + "costs": np.asarray(costs).reshape(kk), p + "U_new": np.asarray(update)[..., 0],
                    p + "next": np.asarray(nxt).reshape(6), p + "U_shift": np.asarray(shifted)[..., 0],
                    p + "meta": np.array([kk, tau, lam, gamma, upsilon, float(norm)], np.float64)})
    path = os.path.join(HERE, "nn_auv_fixtures.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
