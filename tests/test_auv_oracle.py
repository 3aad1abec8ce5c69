"""AUV (Fossen) dynamics: the C restatement (oracle/mppi_oracle_impl.h, orc_auv_*) against golden vectors
produced by the reference's own AUVModel run on tests/golden/tf_shim (tests/golden/gen_auv_fixtures.py), and
against the rotation matrices the reference's test expects (scripts/test.py:268-300)."""
import json
import os

import numpy as np
import pytest

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "auv_fixtures.npz")


def load(which):
    d = np.load(FIX)
    prm = json.loads(bytes(d["params_json"]).decode())[which]
    return prm, (lambda key: d[f"{which}_{key}"])


@pytest.mark.parametrize("which", ["test", "full"])
@pytest.mark.parametrize("rk", [1, 2, 4])
def test_step_matches_reference(oracle64, which, rk):
    prm, g = load(which)
    got = oracle64.auv_step(prm, 0.1, rk, g("state"), g("action"))
    np.testing.assert_allclose(got, g(f"next_rk{rk}"), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(np.linalg.norm(got[:, 3:7], axis=1), 1.0, rtol=1e-12)


@pytest.mark.parametrize("which", ["test", "full"])
def test_state_dot_and_components(oracle64, which):
    prm, g = load(which)
    st, ac = g("state"), g("action")
    xd = oracle64.auv_state_dot(prm, st, ac)
    np.testing.assert_allclose(xd, g("state_dot"), rtol=1e-10, atol=1e-11)
    # the components the reference's tests look at, recombined the way acc() does (auv_model.py:544-559)
    nu = st[:, 7:13]
    rhs = ac - np.einsum("kij,kj->ki", g("coriolis"), nu) - np.einsum("kij,kj->ki", g("damping"), nu) - g("restoring")
    acc = np.einsum("ij,kj->ki", np.linalg.inv(g("mtot")), rhs)
    np.testing.assert_allclose(xd[:, 7:], acc, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(xd[:, :3], np.einsum("kij,kj->ki", g("rot"), nu[:, :3]), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(xd[:, 3:7], np.einsum("kij,kj->ki", g("tquat"), nu[:, 3:]), rtol=1e-10, atol=1e-12)


def test_rotation_known_answers(oracle64):
    """scripts/test.py:268-300: quaternions (x, y, z, w) and the body-to-inertial rotations 'from lib'."""
    prm, _ = load("test")
    quats = np.array([[0., 0., 0., 1.],
                      [0.0438308910967523, 0.25508068761447, 0.171880267220619, 0.950510320581509],
                      [-0.111618880991033, 0.633022223770408, 0.492403876367579, 0.586824089619078]])
    for q in quats:
        for axis in range(3):
            st = np.zeros(13)
            st[3:7] = q
            st[7 + axis] = 1.0                                # body velocity e_axis -> pose_dot[:3] = R[:, axis]
            xd = oracle64.auv_state_dot(prm, st[None], np.zeros((1, 6)))[0]
            x, y, z, w = q
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            np.testing.assert_allclose(xd[:3], R[:, axis], rtol=1e-12, atol=1e-14)
    # second quaternion: the reference's literal (scripts/test.py:288-292)
    st = np.zeros(13); st[3:7] = quats[1]; st[7] = 1.0
    xd = oracle64.auv_state_dot(prm, st[None], np.zeros((1, 6)))[0]
    np.testing.assert_allclose(xd[:3], [0.8107820, 0.3491088, -0.4698463], atol=2e-7)
    st = np.zeros(13); st[3:7] = quats[2]; st[9] = 1.0      # third quaternion, third column (:295-299)
    xd = oracle64.auv_state_dot(prm, st[None], np.zeros((1, 6)))[0]
    np.testing.assert_allclose(xd[:3], [0.6330222, 0.7544065, 0.1736482], atol=2e-7)
