"""numpy stand-in for the five tensorflow_graphics quaternion functions ElipseCost3D calls
(scripts/src/costs/elipse_cost.py:166-246).  TEST TOOLING ONLY.  tensorflow_graphics is an unpinned, un-vendored
dependency of the reference (imported at elipse_cost.py:3) and is not installable here; the functions below restate
its published algorithms (tensorflow_graphics/geometry/transformation/quaternion.py; quaternions are (x, y, z, w)).
They are pinned by the reference's own known-answer tests, replayed by tests/golden/gen_ellipse3d_fixtures.py:
test_tf_rot (rotate, multiply), test_orientation_error (between_two_vectors_3d, relative_angle) and test_prep_const.
Not restated: the (1 - eps) shrink relative_angle applies to the dot product before acos (safe_ops.safe_shrink) and the
eps safe_unsigned_div adds to denominators — both below 1e-14."""
import types

import numpy as np


def _multiply(q1, q2, name=None):
    q1, q2 = np.asarray(q1, np.float64), np.asarray(q2, np.float64)
    x1, y1, z1, w1 = np.moveaxis(q1, -1, 0)
    x2, y2, z2, w2 = np.moveaxis(q2, -1, 0)
    x = x1 * w2 + y1 * z2 - z1 * y2 + w1 * x2
    y = -x1 * z2 + y1 * w2 + z1 * x2 + w1 * y2
    z = x1 * y2 - y1 * x2 + z1 * w2 + w1 * z2
    w = -x1 * x2 - y1 * y2 - z1 * z2 + w1 * w2
    return np.stack((x, y, z, w), axis=-1)


def _conjugate(q):
    q = np.asarray(q, np.float64)
    return np.concatenate((-q[..., :3], q[..., 3:]), axis=-1)


def _rotate(point, quaternion, name=None):
    point, quaternion = np.asarray(point, np.float64), np.asarray(quaternion, np.float64)
    p = np.concatenate((point, np.zeros(point.shape[:-1] + (1,))), axis=-1)
    p = _multiply(quaternion, p)
    p = _multiply(p, _conjugate(quaternion))
    return p[..., :3]


def _from_rotation_matrix(R, name=None):
    R = np.asarray(R, np.float64)
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        sq = np.sqrt(tr + 1.0) * 2.0
        return np.array([(R[2, 1] - R[1, 2]) / sq, (R[0, 2] - R[2, 0]) / sq, (R[1, 0] - R[0, 1]) / sq, 0.25 * sq])
    if R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        sq = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2.0
        return np.array([0.25 * sq, (R[0, 1] + R[1, 0]) / sq, (R[0, 2] + R[2, 0]) / sq, (R[2, 1] - R[1, 2]) / sq])
    if R[1, 1] > R[2, 2]:
        sq = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2.0
        return np.array([(R[0, 1] + R[1, 0]) / sq, 0.25 * sq, (R[1, 2] + R[2, 1]) / sq, (R[0, 2] - R[2, 0]) / sq])
    sq = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2.0
    return np.array([(R[0, 2] + R[2, 0]) / sq, (R[1, 2] + R[2, 1]) / sq, 0.25 * sq, (R[1, 0] - R[0, 1]) / sq])


def _l2n(v):
    return v / np.sqrt(np.maximum(np.sum(v * v, axis=-1, keepdims=True), 1e-12))


def _between_two_vectors_3d(v1, v2, name=None):
    v1, v2 = np.asarray(v1, np.float64), np.asarray(v2, np.float64)
    v1, v2 = np.broadcast_arrays(_l2n(v1), _l2n(v2))
    cos = np.sum(v1 * v2, axis=-1, keepdims=True)
    real = 1.0 + cos
    axis = np.cross(v1, v2)
    x, y, z = v1[..., 0:1], v1[..., 1:2], v1[..., 2:3]
    anti = np.where(np.abs(x) > np.abs(y), np.concatenate((-z, np.zeros_like(x), x), -1), np.concatenate((np.zeros_like(x), -z, y), -1))
    rot = np.where(real < 1e-6, np.concatenate((anti, np.zeros_like(real)), -1), np.concatenate((axis, real), -1))
    return _l2n(rot)


def _relative_angle(q1, q2, name=None):
    dot = np.sum(np.asarray(q1, np.float64) * np.asarray(q2, np.float64), axis=-1)
    return 2.0 * np.arccos(np.minimum(np.abs(dot), 1.0))


geometry = types.SimpleNamespace(transformation=types.SimpleNamespace(quaternion=types.SimpleNamespace(
    multiply=_multiply, rotate=_rotate, from_rotation_matrix=_from_rotation_matrix,
    between_two_vectors_3d=_between_two_vectors_3d, relative_angle=_relative_angle, conjugate=_conjugate)))


# nn_model.py:10-17 patches tensorflow_graphics.util.shape._get_dim through sys.modules at import time
import sys as _sys                                                                          # noqa: E402
_shape_mod = types.ModuleType("tensorflow_graphics.util.shape")
_util_mod = types.ModuleType("tensorflow_graphics.util")
_util_mod.shape = _shape_mod
_sys.modules.setdefault("tensorflow_graphics.util", _util_mod)
_sys.modules.setdefault("tensorflow_graphics.util.shape", _shape_mod)
util = _util_mod
