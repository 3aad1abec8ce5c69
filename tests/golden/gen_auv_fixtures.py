"""Generates tests/golden/auv_fixtures.npz by running the REFERENCE'S OWN AUVModel
(/root/reference/scripts/src/models/auv_model.py: state_dot, step with rk = 1, 2, 4, normalize_quat and the
component matrices) against tests/golden/tf_shim, on the parameter set of the reference's TestAUVModel
(scripts/test.py:237-266) and on a second one with an off-centre centre of gravity, full added mass and full
damping matrices.  Run here only (never on the GPU box):  python tests/golden/gen_auv_fixtures.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(1, "/root/reference")

from scripts.src.models.auv_model import AUVModel       # noqa: E402


def params(which):
    if which == "test":                                   # scripts/test.py:239-263
        return dict(mass=1000, volume=1.5, density=1000, height=1.6, length=2.5, width=1.5, cog=[0, 0, 0], cob=[0, 0, 0.5],
                    Ma=(500. * np.eye(6)).tolist(), linear_damping=[-70., -70., -700., -300., -300., -100.],
                    quad_damping=[-740., -990., -1800., -670., -770., -520.],
                    linear_damping_forward_speed=[1., 2., 3., 4., 5., 6.],
                    inertial=dict(ixx=650.0, iyy=750.0, izz=550.0, ixy=1.0, ixz=2.0, iyz=3.0))
    rng = np.random.default_rng(42)
    A = rng.uniform(-30, 30, (6, 6))
    return dict(mass=1862.87, volume=1.8382, density=1028.0, height=1.6, length=2.6, width=1.5, cog=[0.02, -0.01, 0.05],
                cob=[0.0, 0.01, 0.3], Ma=(np.diag([779.8, 1222., 3659.9, 534.9, 842.7, 224.3]) + 0.5 * (A + A.T)).tolist(),
                linear_damping=(np.diag([-74.82, -69.48, -728.4, -268.8, -309.77, -105.]) + rng.uniform(-5, 5, (6, 6))).tolist(),
                quad_damping=[-748.22, -992.53, -1821.01, -672., -774.44, -523.27],
                linear_damping_forward_speed=rng.uniform(-3, 3, (6, 6)).tolist(),
                inertial=dict(ixx=525.39, iyy=794.2, izz=691.23, ixy=1.44, ixz=33.41, iyz=2.6))


def run(which, k, seed):
    prm = params(which)
    rng = np.random.default_rng(seed)
    st = rng.uniform(-1, 1, (k, 13, 1))
    st[:, 3:7] /= np.linalg.norm(st[:, 3:7], axis=1, keepdims=True)
    st[0, :, 0] = [0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]                       # scripts/test.py:545
    ac = rng.uniform(-200, 200, (k, 6, 1))
    ac[0, :, 0] = 1.0
    out = {"state": st[..., 0], "action": ac[..., 0]}
    for rk in (1, 2, 4):
        m = AUVModel({}, actionDim=6, dt=0.1, parameters=dict(prm, rk=rk))
        m._k = k
        out[f"next_rk{rk}"] = np.asarray(m.build_step_graph("step", st, ac))[..., 0]
    m = AUVModel({}, actionDim=6, dt=0.1, parameters=dict(prm, rk=1))
    m._k = k
    out["state_dot"] = np.asarray(m.state_dot(st, ac))[..., 0]
    pose, speed = m.prepare_data(st)
    m.body2inertial_transform(pose)
    out["rot"] = np.asarray(m._rotBtoI)
    out["tquat"] = np.asarray(m._TBtoIquat)
    out["damping"] = np.asarray(m.damping_matrix("d", speed))
    out["coriolis"] = np.asarray(m.coriolis_matrix("c", speed))
    out["restoring"] = np.asarray(m.restoring_forces("r"))[..., 0]
    out["mtot"] = np.asarray(m._mTot)
    flat = {f"{which}_{key}": v for key, v in out.items()}
    return flat, prm


def main():
    import json
    store, prms = {}, {}
    for i, which in enumerate(("test", "full")):
        flat, prm = run(which, 48, 300 + i)
        store.update(flat)
        prms[which] = prm
    # the quaternions of test_B2I_transform_and_jacobian with the rotation matrices the reference expects (:283-300)
    store["params_json"] = np.frombuffer(json.dumps(prms).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "auv_fixtures.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
