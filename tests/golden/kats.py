"""Known-answer vectors of the reference's own tests for the MPPI update path, transcribed
by hand (TensorFlow is not installable here, so the reference's gtest/tf.test suites cannot be
run).  Every entry cites the file:line under /root/reference it comes
from.  Shapes drop the reference's trailing singleton dimension.

These vectors pin (a) the CPU oracle (tests/test_oracle_kats.py) and (b) the CUDA stage
entry points of the C-ABI (tests/test_kats_gpu.py).
"""
import numpy as np

# ---------------------------------------------------------------------------------------------
# utile::blockDiag — test/test_utile.cpp:63-173 (dt = 0.01, m = 1.5, nb = 1..4)
# ---------------------------------------------------------------------------------------------
UTILE_DT = 0.01
UTILE_M = 1.5


def blockdiag_expected(nb, dt=UTILE_DT, m=UTILE_M):
    """Expected A, B of blockDiagTest{nb}: the literal exp_a/exp_b arrays of
    test/test_utile.cpp:66-67 (nb=1), :86-93 (nb=2), :113-124 (nb=3), :144-159 (nb=4) are
    exactly the nb-fold block diagonals of A=[[1,dt],[0,1]] and B=[[dt^2/(2m)],[dt/m]]."""
    A = np.zeros((2 * nb, 2 * nb), np.float32)
    B = np.zeros((2 * nb, nb), np.float32)
    for i in range(nb):
        A[2 * i, 2 * i] = 1.0
        A[2 * i, 2 * i + 1] = np.float32(dt)
        A[2 * i + 1, 2 * i + 1] = 1.0
        B[2 * i, i] = np.float32(dt) * np.float32(dt) / (np.float32(2.0) * np.float32(m))
        B[2 * i + 1, i] = np.float32(dt) / np.float32(m)
    return A, B


# Literal nb=2 arrays, test/test_utile.cpp:86-92, kept verbatim as a check on the helper above.
BLOCKDIAG2_A_LITERAL = lambda dt: np.array([[1, dt, 0, 0], [0, 1, 0, 0], [0, 0, 1, dt], [0, 0, 0, 1]],
                                           np.float32)

# ---------------------------------------------------------------------------------------------
# ModelBase — test/test_model.cpp
# ---------------------------------------------------------------------------------------------
MODEL_CASES = []

# StepTesting1 — test/test_model.cpp:120-146 (fixture :32-39): k=1,s=2,a=1,m=1,dt=.01
MODEL_CASES.append(dict(name="StepTesting1", cite="test/test_model.cpp:120-146",
                        k=1, s=2, a=1, m=1.0, dt=0.01,
                        state=[[0., 0.]], action=[[1.]]))
# StepTesting2 — test/test_model.cpp:148-176 (fixture :42-49): k=1,s=4,a=2,m=2
MODEL_CASES.append(dict(name="StepTesting2", cite="test/test_model.cpp:148-176",
                        k=1, s=4, a=2, m=2.0, dt=0.01,
                        state=[[0., 0., 0., 0.]], action=[[1., 1.]]))
# LargeTesting — test/test_model.cpp:178-214 (fixture :51-70): k=5,s=6,a=3,m=1.5
_STATE3 = [[0., 0., 0., 0., 0., 0.],
           [2., 1., 5., 0., -1., -2.],
           [0.5, 0.5, 0.5, 0.5, 0.5, 0.5],
           [1., 0., 1., 0., 1., 0.],
           [-1, 0.5, -3, 2., 0., 0.]]
_ACTION3 = [[1., 1., 1.], [2., 0., -1.], [0., 0., 0.], [0.5, -0.5, 0.5], [3., 3., 3.]]
MODEL_CASES.append(dict(name="LargeTesting", cite="test/test_model.cpp:178-214",
                        k=5, s=6, a=3, m=1.5, dt=0.01, state=_STATE3, action=_ACTION3))
# InitTest — test/test_model.cpp:216-255: state [1,s,1] broadcast against [k,a,1] actions
MODEL_CASES.append(dict(name="InitTest", cite="test/test_model.cpp:216-255",
                        k=5, s=6, a=3, m=1.5, dt=0.01,
                        state=[[-1, 0.5, -3, 2., 0., 0.]], action=_ACTION3))


def model_expected(case):
    """exp_s (free step), exp_u (action step), exp_res as the tests build them
    (test/test_model.cpp:123-129,151-158,182-199,221-238): float32 arithmetic,
    acc = dt*dt/(2*m), vel = dt/m."""
    f = np.float32
    dt, m = f(case["dt"]), f(case["m"])
    acc = (dt * dt) / (f(2.0) * m)
    vel = dt / m
    st = np.array(case["state"], np.float32)
    ac = np.array(case["action"], np.float32)
    exp_u = np.empty((ac.shape[0], case["s"]), np.float32)
    exp_u[:, 0::2] = ac * acc
    exp_u[:, 1::2] = ac * vel
    exp_s = st.copy()
    exp_s[:, 0::2] = st[:, 0::2] + st[:, 1::2] * dt
    exp_res = exp_u + exp_s
    return exp_s, exp_u, exp_res


# Literal expectations of LargeTesting (test/test_model.cpp:185-197, dt3 = 0.01), kept verbatim
# so the formula in model_expected is itself checked against the reference's numbers.
def large_testing_literal():
    f = np.float32
    dt3, m3 = f(0.01), f(1.5)
    acc = (dt3 * dt3) / (f(2.) * m3)
    vel = dt3 / m3
    exp_u = np.array([acc, vel, acc, vel, acc, vel,
                      f(2.) * acc, f(2.) * vel, f(0.) * acc, 0 * vel, f(-1.) * acc, f(-1.) * vel,
                      f(0.) * acc, f(0.) * vel, f(0.) * acc, 0 * vel, f(0.) * acc, f(0.) * vel,
                      f(.5) * acc, f(.5) * vel, f(-.5) * acc, f(-.5) * vel, f(.5) * acc, f(.5) * vel,
                      f(3.) * acc, f(3.) * vel, f(3.) * acc, f(3.) * vel, f(3.) * acc, f(3.) * vel],
                     np.float32).reshape(5, 6)
    exp_s = np.array([0., 0., 0., 0., 0., 0.,
                      f(2.) + dt3, 1., 5., 0., f(-1.) - f(2.) * dt3, -2.,
                      f(.5) + dt3 / f(2.), .5, f(.5) + dt3 / f(2.), .5, f(.5) + dt3 / f(2.), .5,
                      1., 0., 1., 0., 1., 0.,
                      f(-1.) + dt3 / f(2.), .5, f(-3.) + f(2.) * dt3, 2., 0., 0.],
                     np.float32).reshape(5, 6)
    return exp_s, exp_u, exp_u + exp_s


# Python twin, 3-step rollout — scripts/test.py:173-218 (dt = 0.1, m = 1.5, fp64)
def py_step3_expected():
    dt, m = 0.1, 1.5
    acc = dt * dt / (2. * m)
    vel = dt / m
    st = np.array(_STATE3, np.float64)
    ac = np.array(_ACTION3, np.float64)
    B_u = np.empty((5, 6))
    B_u[:, 0::2] = ac * 3 * (acc + vel * dt)
    B_u[:, 1::2] = ac * vel * 3
    exp_x = st.copy()
    exp_x[:, 0::2] = st[:, 0::2] + st[:, 1::2] * dt * 3
    return dict(dt=dt, m=m, state=st, action=ac, expected=B_u + exp_x)


# ---------------------------------------------------------------------------------------------
# CostBase — test/test_cost.cpp (Sigma = I, lambda = 1)
# ---------------------------------------------------------------------------------------------
COST_CASES = [
    # scenario 1 — fixture test/test_cost.cpp:29-54; expectations :172-173 (state) / :208 (step)
    dict(name="scenario1", cite="test/test_cost.cpp:29-54,169-178,205-215",
         k=1, s=2, a=2, lam=1.0,
         state=[[0., 1.]], goal=[1., 1.], action=[1., 1.], noise=[[1., 1.]],
         sigma=[[1., 0.], [0., 1.]], q=[1., 1.],
         exp_state=[1.], exp_step=[3.]),
    # scenario 2 — fixture :56-84; expectations :181-182 / :218
    dict(name="scenario2", cite="test/test_cost.cpp:56-84,180-189,217-225",
         k=1, s=4, a=2, lam=1.0,
         state=[[0., 0.5, 2., 0.]], goal=[1., 1., 1., 2.], action=[0.5, 2.], noise=[[0.5, 1.]],
         sigma=[[1., 0.], [0., 1.]], q=[1., 1., 10., 10.],
         exp_state=[51.25], exp_step=[53.5]),
    # scenario 3 — fixture :86-127; expectations :193-194 / :228
    dict(name="scenario3", cite="test/test_cost.cpp:86-127,191-201,227-237",
         k=5, s=4, a=3, lam=1.0,
         state=[[0., 0.5, 2., 0.], [0., 2., 0., 0.], [10., 2., 2., 3], [1., 1., 1., 2.],
                [3., 4., 5., 6.]],
         goal=[1., 1., 1., 2.], action=[0.5, 2., 0.25],
         noise=[[0.5, 1., 2.], [0.5, 2., 0.25], [-2, -0.2, -1], [0, 0, 0], [1., 0.5, 3.]],
         sigma=[[1., 0., 0.], [0., 1., 0.], [0., 0., 1.]], q=[1., 1., 10., 10.],
         exp_state=[51.25, 52, 102, 0., 333],
         exp_step=[51.25 + 2.75, 52 + 4.3125, 102 - 1.65, 0. + 0, 333 + 2.25]),
]

# Python twin StaticCost, s=13, a=6 — scripts/test.py:944-1095 gives state costs
# {911.25, 917.5}; that case uses a non-point-mass state (quaternion AUV) and a full Q matrix,
# which is outside the C++ API (Q = Diag(q), src/cost_base.cpp:40) — listed, not replayed.

# ---------------------------------------------------------------------------------------------
# Python-twin action cost — scripts/test.py:685-838 (CostBase.action_cost, Sigma = I):
#   0.5 * (gamma * (u.u + 2 u.eps) + lambda * (1 - 1/upsilon) * eps.eps), u = action (a), eps = noise [k][a]
# The expected values are transcribed as the reference writes them (its literals for u.u, u.eps, eps.eps).
# ---------------------------------------------------------------------------------------------
_PY_ACTION3 = [0.5, 2.0, 0.25]
_PY_NOISE3 = [[0.5, 1., 2.], [0.5, 2., 0.25], [-2, -0.2, -1], [0., 0., 0.], [1., 0.5, 3.]]
_PY_TERMS3 = [(4.3125, 2.75, 5.25), (4.3125, 4.3125, 4.3125), (4.3125, -1.65, 5.04), (4.3125, 0.0, 0.0), (4.3125, 2.25, 10.25)]


def _py_ac(lam, gamma, upsilon, terms):
    return [0.5 * (gamma * (uu + 2. * ue) + lam * (1 - 1. / upsilon) * ee) for uu, ue, ee in terms]


PY_ACTION_COST_CASES = [
    dict(name="testStepCost_s2_a2_l1", lam=1., gamma=1., upsilon=1., action=[1., 1.], noise=[[1., 1.]],        # :689-708
         expected=_py_ac(1., 1., 1., [(2., 2., 0.)])),
    dict(name="testStepCost_s4_a2_l1", lam=1., gamma=1., upsilon=1., action=[0.5, 2.], noise=[[0.5, 1.]],      # :710-730
         expected=_py_ac(1., 1., 1., [(4.25, 2.25, 1.25)])),
    dict(name="testStepCost_s4_a3_l1", lam=1., gamma=1., upsilon=1., action=_PY_ACTION3, noise=_PY_NOISE3,     # :732-766
         expected=_py_ac(1., 1., 1., _PY_TERMS3)),
    dict(name="testStepCost_s4_a3_l10_g2_u3", lam=10., gamma=2., upsilon=3., action=_PY_ACTION3,               # :768-802
         noise=_PY_NOISE3, expected=_py_ac(10., 2., 3., _PY_TERMS3)),
    dict(name="testStepCost_s4_a3_l15_g20_u30", lam=15., gamma=20., upsilon=30., action=_PY_ACTION3,           # :804-838
         noise=_PY_NOISE3, expected=_py_ac(15., 20., 30., _PY_TERMS3)),
]

# StaticCost on the 13-dimensional AUV state with a = 6 — scripts/test.py:944-1095.  Q is diagonal there
# (1 on pose, 10 on velocities), so the case replays through the diagonal-Q state cost and the a = 6 action cost.
PY_STATIC13 = dict(
    state=[[0., 0.5, 2., 0., 0., 0., 1., 1., 2., 3., 4., 5., 6.], [0., 2., 0., 0., 0.5, 0.5, 0., 4., 5., 6., 1., 2., 3.]],
    goal=[1., 1., 2., 0., 0., 0., 1., 0., 0., 0., 0., 0., 0.],
    q=[1.] * 7 + [10.] * 6,
    action=[0.5, 2., 0.25, 4., 1., 1.5],
    noise=[[0.5, 1., 2., 3., 4., 5.], [0.5, 2., 0.25, 1.25, 2.5, 0.75]],
    lam=1., gamma=1., upsilon=1.,
    expected_action=_py_ac(1., 1., 1., [(23.5625, 26.25, 55.25), (23.5625, 12.9375, 12.6875)]),   # :1073-1080
    expected_state=[911.25, 917.5],                                                               # :1082-1089
)

# ---------------------------------------------------------------------------------------------
# ElipseCost (Python twin) — scripts/test.py:1098-1161: a = b = 1, centre (0, 0), speed 1,
# m_state = m_vel = 1; states (x, vx, y, vy)
# ---------------------------------------------------------------------------------------------
ELLIPSE_PARAMS = (1.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0)        # a, b, cx, cy, speed, m_state, m_vel (:1102-1109)
ELLIPSE_CASES = [
    dict(name="testStepElipseCost_s4_l1_k1",                # :1111-1131
         state=[[0., 0.5, 1., 0.]], expected=[0.25]),
    dict(name="testStepElipseCost_s4_l1_k5",                # :1133-1161
         state=[[0., 0.5, 1., 0.], [0., 2., 0., 0.], [10., 2., 2., 3.], [1., 1., 1., 2.], [3., 4., 5., 6.]],
         expected=[0.25, 2, 103 + 6.788897449072021, 1 + 1.5278640450004208, 33 + 38.57779489814404]),
]

# ---------------------------------------------------------------------------------------------
# ControllerBase — test/test_controller.cpp (k=5, tau=3, a_dim=2, s_dim=4, dt=0.01, lambda=1)
# ---------------------------------------------------------------------------------------------
CTRL = dict(
    cite="test/test_controller.cpp:17-39",
    k=5, tau=3, a=2, s=4, lam=1.0,
    cost=[3., 10., 0., 1., 5.],                                          # :25
    noise=np.array([1., -0.5, 1., -0.5, 2., 1.,
                    0.3, 0, 2., 0.2, 1.2, 3.,
                    0.5, 0.5, 0.5, 0.5, 0.5, 0.5,
                    0.6, 0.7, 0.2, -0.3, 0.1, -0.4,
                    -2., -3., -4., -1., 0., 0.]).reshape(5, 3, 2),        # :26-30
    action=np.array([1., 0.5, 2.3, 4.5, 2.1, -0.4]).reshape(3, 2),      # :32
)

# testDataPrep — test/test_controller.cpp:71-107
CTRL_PREP = dict(
    a=[[1., 0.5], [2.3, 4.5], [2.1, -0.4]],                              # :77-79
    n=[np.array([1., -0.5, 0.3, 0, 0.5, 0.5, 0.6, 0.7, -2., -3.]).reshape(5, 2),    # :80
       np.array([1., -0.5, 2., 0.2, 0.5, 0.5, 0.2, -0.3, -4, -1]).reshape(5, 2),    # :81
       np.array([2., 1., 1.2, 3., 0.5, 0.5, 0.1, -0.4, 0., 0.]).reshape(5, 2)],     # :82
)

# testUpdate — test/test_controller.cpp:109-167
_W = [0.034951787275480706, 3.1871904480408675e-05, 0.7020254138530686, 0.2582607169364174,
      0.004730210030553017]                                              # :128-132
CTRL_UPDATE = dict(
    beta=0.0,                                                            # :112
    exp_arg=[-3., -10., 0, -1., -5.],                                    # :115
    exp=[0.049787068367863944, 4.5399929762484854e-05, 1, 0.36787944117144233,
         0.006737946999085467],                                          # :118-122
    nabla=1.424449856468154,                                             # :125
    weights=_W,
    weighted_noise=np.array([                                            # :135-140
        _W[0] * 1. + _W[1] * 0.3 + _W[2] * 0.5 + _W[3] * 0.6 + _W[4] * (-2),
        _W[0] * (-0.5) + _W[1] * 0 + _W[2] * 0.5 + _W[3] * 0.7 + _W[4] * (-3),
        _W[0] * 1 + _W[1] * 2 + _W[2] * 0.5 + _W[3] * 0.2 + _W[4] * (-4),
        _W[0] * (-0.5) + _W[1] * 0.2 + _W[2] * 0.5 + _W[3] * (-0.3) + _W[4] * (-1),
        _W[0] * 2 + _W[1] * 1.2 + _W[2] * 0.5 + _W[3] * 0.1 + _W[4] * 0,
        _W[0] * 1 + _W[1] * 3 + _W[2] * 0.5 + _W[3] * (-0.4) + _W[4] * 0]).reshape(3, 2),
    sum_w=1.0,                                                           # :164
)

# testNew — test/test_controller.cpp:169-193 (nb = 0..3; nb=0 is the empty [0,2,1] tensor)
CTRL_NEW = {0: np.zeros((0, 2)), 1: [[1, 0.5]], 2: [[1, 0.5], [2.3, 4.5]],
            3: [[1., 0.5], [2.3, 4.5], [2.1, -0.4]]}

# testShiftAndInit — test/test_controller.cpp:195-222
CTRL_SHIFT = [
    dict(nb=1, init=[[1, 0.5]], expected=np.array([2.3, 4.5, 2.1, -0.4, 1., 0.5]).reshape(3, 2)),
    dict(nb=2, init=[[1, 0.5], [2.3, 4.5]],
         expected=np.array([2.1, -0.4, 1., 0.5, 2.3, 4.5]).reshape(3, 2)),
]

# ---------------------------------------------------------------------------------------------
# Philox4x32-10 known-answer vectors (Random123 kat_vectors; the noise-stream integer contract,
# see oracle/mppi_oracle.c).  Third-party algorithm: Salmon, Moraes, Dror, Shaw, "Parallel random
# numbers: as easy as 1, 2, 3", SC'11 — the same generator TF's RandomNormal is built on.
# ---------------------------------------------------------------------------------------------
PHILOX_KATS = [
    dict(ctr=[0, 0, 0, 0], key=[0, 0],
         out=[0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    dict(ctr=[0xffffffff] * 4, key=[0xffffffff] * 2,
         out=[0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    dict(ctr=[0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], key=[0xa4093822, 0x299f31d0],
         out=[0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]

# Philox4x32-7 (mppi_config.philox_rounds = 7): the seven-round lines of the same Random123 kat_vectors file
# ("philox4x32 7 <ctr> <key> <out>": zeros, all ones, and the digits of pi).
PHILOX7_KATS = [
    dict(ctr=[0, 0, 0, 0], key=[0, 0],
         out=[0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48]),
    dict(ctr=[0xffffffff] * 4, key=[0xffffffff] * 2,
         out=[0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662]),
    dict(ctr=[0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], key=[0xa4093822, 0x299f31d0],
         out=[0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a]),
]
