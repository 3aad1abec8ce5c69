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


def _ref_test_expectations(prm, vel, off_diag_sign=-1.0):
    """The expected damping and Coriolis matrices exactly as the reference's tests build them
    (scripts/test.py:500-541), for one body velocity `vel` [6].  test_corrolis writes the inertia products with a minus
    sign (:520-522) while the model's get_inertial (auv_model.py:265-280) uses plus; its only input has zero angular
    velocity, where the sign cannot matter.  off_diag_sign = +1 selects the model's convention for other inputs."""
    ld, lf, qd = (np.asarray(prm[k], float) for k in ("linear_damping", "linear_damping_forward_speed", "quad_damping"))
    D = -np.diag(ld) - vel[0] * np.diag(lf) + (-np.diag(qd)) * np.abs(vel)[:, None]           # :510-514
    i, m = dict(prm["inertial"]), prm["mass"]
    for key in ("ixy", "ixz", "iyz"):
        i[key] = -off_diag_sign * i[key]
    Iv = [i["ixx"] * vel[3] - i["ixy"] * vel[4] - i["ixz"] * vel[5], -i["ixy"] * vel[3] + i["iyy"] * vel[4] - i["iyz"] * vel[5],
          -i["ixz"] * vel[3] - i["iyz"] * vel[4] + i["izz"] * vel[5]]                             # :520-522
    Mav = -np.asarray(prm["Ma"], float) @ vel                                                   # :523
    crb = np.array([[0, 0, 0, 0, m * vel[2], -m * vel[1]], [0, 0, 0, -m * vel[2], 0, m * vel[0]], [0, 0, 0, m * vel[1], -m * vel[0], 0],
                    [0, m * vel[2], -m * vel[1], 0, Iv[2], -Iv[1]], [-m * vel[2], 0, m * vel[0], -Iv[2], 0, Iv[0]],
                    [m * vel[1], -m * vel[0], 0, Iv[1], -Iv[0], 0]], float)                      # :525-530
    ca = np.array([[0, 0, 0, 0, -Mav[2], Mav[1]], [0, 0, 0, Mav[2], 0, -Mav[0]], [0, 0, 0, -Mav[1], Mav[0], 0],
                   [0, -Mav[2], Mav[1], 0, -Mav[5], Mav[4]], [Mav[2], 0, -Mav[0], Mav[5], 0, -Mav[3]],
                   [-Mav[1], Mav[0], 0, -Mav[4], Mav[3], 0]], float)                             # :532-537
    return D, crb + ca


def test_reference_damping_and_coriolis_kats(oracle64):
    """scripts/test.py:500-541 (test_damping, test_corrolis): with u = 0 the acceleration is M^-1(-C nu - D nu - g), so
    M acc + g(identity attitude) must equal -(C + D) nu with C and D built the way the reference's tests build them."""
    prm, g = load("test")
    M = g("mtot")
    W, B = prm["mass"] * 9.81, prm["volume"] * prm["density"] * 9.81
    g_id = np.array([0, 0, -(B - W), 0, 0, 0.0])          # restoring at identity attitude: -(fbb + fbg), no moment (cob on the z axis)
    vels = np.array([[1, 1, 1, 1, 1, 1], [2, 1.5, 1, 3, 3.5, 2.5], [-2, -1.5, -1, -3, -3.5, -2.5], [1, 1, 1, 0, 0, 0]], float)
    for vel in vels:
        st = np.zeros(13)
        st[6] = 1.0
        st[7:] = vel
        acc = oracle64.auv_state_dot(prm, st[None], np.zeros((1, 6)))[0, 7:]
        literal = not vel[3:].any()                           # the reference's own Coriolis input (:519)
        D, C = _ref_test_expectations(prm, vel, off_diag_sign=-1.0 if literal else 1.0)
        np.testing.assert_allclose(M @ acc + g_id, -(C + D) @ vel, rtol=1e-10, atol=1e-8)


def test_reference_restoring_kat(oracle64):
    """scripts/test.py:417-498 (test_restoring): Euler-angle rotation matrices of the test against the quaternion poses it
    feeds (7-digit literals), restoring vector -(fbb + fbg, r_b x fbb + r_g x fbg)."""
    prm, g = load("test")
    M = g("mtot")
    roll, pitch, yaw = (np.array(v) * np.pi / 180 for v in ([13., 280.], [110., 50.], [25., 325.]))
    quats = np.array([[-0.1127657, 0.8086476, 0.0328141, 0.5764513], [-0.4582488, 0.4839407, 0.0503092, 0.7438269]])
    W, B = prm["mass"] * 9.81, prm["volume"] * prm["density"] * 9.81
    for i in range(2):
        cr, sr, cp, sp, cy, sy = np.cos(roll[i]), np.sin(roll[i]), np.cos(pitch[i]), np.sin(pitch[i]), np.cos(yaw[i]), np.sin(yaw[i])
        RbI = np.array([[cy * cp, -sy * cr + cy * sp * sr, sy * sr + cy * cr * sp],
                        [sy * cp, cy * cr + sr * sp * sy, -cy * sr + sp * sy * cr],
                        [-sp, cp * sr, cp * cr]])                                                # :446-461
        fbg, fbb = RbI.T @ [0, 0, -W], RbI.T @ [0, 0, B]
        rest = -np.concatenate([fbb + fbg, np.cross(prm["cob"], fbb) + np.cross(prm["cog"], fbg)])   # :472-491
        st = np.zeros(13)
        st[:3] = [1.5, 2.3, 0.7]
        st[3:7] = quats[i]
        acc = oracle64.auv_state_dot(prm, st[None], np.zeros((1, 6)))[0, 7:]
        np.testing.assert_allclose(M @ acc, -rest, rtol=2e-6, atol=2e-6 * W)
