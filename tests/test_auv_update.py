"""MPPI update with the AUV (Fossen) model and StaticCost / StaticQuatCost:
golden vectors produced by the reference's own Python controller, AUVModel and cost classes
(tests/golden/gen_auv_update_fixtures.py -> auv_update_fixtures.npz), against the C restatement (CPU) and the
CUDA path (GPU, through the C-ABI)."""
import json
import os

import numpy as np
import pytest

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "auv_update_fixtures.npz")
CASES = ["auv_rk1", "auv_rk2", "auv_quat", "auv_quatn", "auv_rk4"]


def load(name):
    d = np.load(FIX)
    info = json.loads(bytes(d["params_json"]).decode())
    prm = info["prm"][info["which"][name]]
    k, tau, rk, lam, gamma, upsilon, normalize, quat = d[f"{name}_meta"]
    meta = dict(k=int(k), tau=int(tau), rk=int(rk), lam=float(lam), gamma=float(gamma), upsilon=float(upsilon),
                normalize=bool(normalize), quat=bool(quat))
    return prm, meta, (lambda key: d[f"{name}_{key}"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_controller(oracle64, name):
    prm, m, g = load(name)
    r = oracle64.mppi_update_auv(prm, 0.1, m["rk"], m["lam"], g("sigma"), g("goal"), g("q"), g("x"), g("U"), g("eps"),
                                 gamma=m["gamma"], upsilon=m["upsilon"], normalize=m["normalize"], quat_cost=m["quat"])
    np.testing.assert_allclose(r["costs"], g("costs_py"), rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(r["U_new"], g("U_new"), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(r["next"], g("next"), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(r["U_shift"], g("U_shift"), rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("name", ["auv_quat", "auv_quatn"])
def test_oracle_quat_cost(oracle64, name):
    _, _, g = load(name)
    got = oracle64.cost_state_quat(g("qc_state"), g("goal"), g("q"))
    np.testing.assert_allclose(got, g("qc_cost"), rtol=1e-12, atol=1e-12)
