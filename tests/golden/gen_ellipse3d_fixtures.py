"""Generates tests/golden/ellipse3d_fixtures.npz from the REFERENCE'S OWN ElipseCost3D
(/root/reference/scripts/src/costs/elipse_cost.py:99-246) run on tests/golden/tf_shim.

tensorflow_graphics is not installable, so its five quaternion functions are restated in the shim; before anything is
stored, the reference's known-answer tests for this class (scripts/test.py:1183-1359: prep_const, position_error,
orientation_error, velocity_error, tf_rot) are replayed through the reference class on the shim and must hold.

state_cost is called one state at a time: for k > 1 the reference adds a [k] orientation term to [k, 1, 1] position and
velocity terms (elipse_cost.py:196), which broadcasts to [k, 1, k] — its own test_state_cost (scripts/test.py:1307)
asserts nothing.  The k = 1 call is well defined and is what the per-sample cost means.
    python tests/golden/gen_ellipse3d_fixtures.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(1, "/root/reference")

import tensorflow_graphics as tfg                                   # noqa: E402
from scripts.src.costs.elipse_cost import ElipseCost3D              # noqa: E402

SIG = np.eye(6)
AXIS = np.array([[2.], [1.5]])


def mk(normal, aVec, center, axis=AXIS, speed=1., ms=1., mv=1.):
    return ElipseCost3D(1., 1., 1., SIG, np.array(normal, float).reshape(3, 1), np.array(aVec, float).reshape(3, 1), axis,
                        np.array(center, float).reshape(3, 1), speed, 1., ms, mv)


def replay_reference_kats():
    c = mk([0, 0, 1], [1, 0, 0], [0, 0, 0])
    np.testing.assert_allclose(c.R, np.eye(3), atol=1e-12)                                   # test.py:1195-1201
    c = mk([0, 1, 1], [1, 0, 0], [0, 1, -2])
    np.testing.assert_allclose(c.R, np.array([[1, 0, 0], [0, .5, -.5], [0, .5, .5]]).T, atol=1e-12)   # :1212-1219
    pos = np.array([[[0.1], [0.4], [0.2]], [[1.], [1.], [-2]], [[2.], [1.], [0.]]])
    np.testing.assert_allclose(c.position_error(pos), np.array([[[0.8863888888888889]], [[3.6944444444444446]], [[0.4444444444444444]]]),
                               rtol=1e-6)                                                    # :1236-1240
    ori = np.array([[[0.1], [0.4], [0.2], [0.0], [0.0], [0.0], [1.]],
                    [[1.], [1.], [-2], [0.48038446], [0.32025631], [0.16012815], [0.80064077]],
                    [[2.], [1.], [-2], [0.20628425], [-0.30942637], [-0.92827912], [0.]]])
    np.testing.assert_allclose(c.orientation_error(ori), [3.0018837793006306, 2.4098026419889416, 1.1216620246733544],
                               rtol=1e-6)                                                    # :1273-1278
    vel = np.array([[[0.1], [0.4], [0.2], [0.0], [0.0], [0.0]], [[1.], [1.], [-2], [0.3], [0.2], [0.1]],
                    [[2.], [1.], [-2], [0.2], [-0.3], [-0.9]]])
    np.testing.assert_allclose(c.velocity_error(vel), np.abs(np.array([[[0.21 - 1]], [[6 - 1]], [[9 - 1]]])), rtol=1e-6)   # :1303-1305
    q = np.array([0., 0.7071068, 0., 0.7071068])
    np.testing.assert_allclose(tfg.geometry.transformation.quaternion.rotate(np.array([1., 2., 3.]), q), [3., 2., -1.], atol=1e-6)
    np.testing.assert_allclose(tfg.geometry.transformation.quaternion.multiply(q, np.array([0.7071068, 0., 0., 0.7071068])),
                               [0.5, 0.5, -0.5, 0.5], atol=1e-6)                             # :1342-1359
    print("reference KATs for ElipseCost3D hold on the shim")


CASES = [
    dict(name="e3_xy", normal=[0, 0, 1], aVec=[1, 0, 0], center=[0, 0, 0], axis=[2.0, 1.5], speed=1.0, ms=1.0, mv=1.0),
    dict(name="e3_tilt", normal=[0, 1, 1], aVec=[1, 0, 0], center=[0, 1, -2], axis=[2.0, 1.5], speed=0.7, ms=2.0, mv=0.5),
    dict(name="e3_gen", normal=[0.36, -0.48, 0.8], aVec=[0.8, 0.6, 0.0], center=[0.3, -0.2, 1.0], axis=[3.0, 1.2], speed=1.3, ms=0.8, mv=1.7),
]


def main():
    replay_reference_kats()
    out = {}
    for i, c in enumerate(CASES):
        rng = np.random.default_rng(900 + i)
        cost = mk(c["normal"], c["aVec"], c["center"], np.array(c["axis"]).reshape(2, 1), c["speed"], c["ms"], c["mv"])
        st = rng.uniform(-2, 2, (48, 13, 1))
        st[:, 3:7] /= np.linalg.norm(st[:, 3:7], axis=1, keepdims=True)
        if i == 0:                                                                           # test.py:1308-1326 (no expected values there)
            st[0, :, 0] = [0.1, 0.4, 0.2, 0, 0, 0, 1, 0.3, 0.7, 2., 1., 2.4, 5.0]
            st[1, :, 0] = [1., 1., -2, 0.3, 0.2, 0.1, 0.5, 0.4, 2.7, 2., 0., 0., 0.]
            st[2, :, 0] = [2., 1., -2, 0.2, -0.3, -0.9, 0.0, 2.3, 1.7, 0., 0.1, 0.4, 0.01]
        costs = np.array([np.asarray(cost.state_cost("c", st[j:j + 1])).reshape(()) for j in range(st.shape[0])])
        p = c["name"] + "_"
        out.update({p + "state": st[..., 0], p + "cost": costs, p + "normal": np.array(c["normal"], float), p + "aVec": np.array(c["aVec"], float),
                    p + "center": np.array(c["center"], float), p + "axis": np.array(c["axis"], float),
                    p + "scal": np.array([c["speed"], c["ms"], c["mv"]]), p + "R": np.asarray(cost.R), p + "q": np.asarray(cost.q)})
    path = os.path.join(HERE, "ellipse3d_fixtures.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
