"""Learner step on the GPU (mppi_mlp_train_step) against the fp64 numpy restatement
of LearnerBase._train_step + Keras Adam (oracle/train_oracle.py)."""
import numpy as np
import pytest

from oracle.train_oracle import KEYS, AdamTrainer
from tests.test_train_oracle import _problem
from tests.util import controller_from_cfg, make_cfg, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,a", [(257, 3), (4096, 3), (100, 1), (1000, 5)])
def test_train_steps_match_oracle(n, a):
    s = 2 * a
    mlp, x, u, xn = _problem(n + a, n=n, s=s, a=a)
    cfg = make_cfg(256, 4, s, a)
    c = controller_from_cfg(cfg)
    try:
        c.setMlp(mlp)
        tr = AdamTrainer(mlp)
        for it in range(5):
            got = c.mlpTrainStep(x, u, xn, 1e-3)
            want = tr.step(x, u, xn, 1e-3)
            assert abs(got - want) <= 2e-5 * abs(want), (it, got, want)
        w = c.mlpGetWeights()
    finally:
        c.close()
    for k in KEYS:
        # five Adam steps of 1e-3 move every weight by about 5e-3: compare the MOVEMENT, not the weight
        moved = tr.w[k] - np.asarray(mlp[k], np.float64)
        assert rel_err(w[k] - np.asarray(mlp[k], np.float32), moved) < 2e-2, k


def test_rollout_uses_the_trained_weights(oracle64):
    """After a step the bf16 copy is refreshed: predict follows the NEW weights (2e-2 bf16 bar)."""
    a, s = 3, 6
    mlp, x, u, xn = _problem(7, n=2048, s=s, a=a)
    c = controller_from_cfg(make_cfg(256, 4, s, a))
    try:
        c.setMlp(mlp)
        tr = AdamTrainer(mlp)
        losses = []
        for _ in range(40):
            losses.append(c.mlpTrainStep(x, u, xn, 3e-3))
            tr.step(x, u, xn, 3e-3)
        assert losses[-1] < 0.5 * losses[0]
        got = c.mlpPredict(x[:300], u[:300])
        c.mlpSetAdam(0.8, 0.99, 1e-6)                     # resets the moments; next step is a first step again
        l0 = c.mlpTrainStep(x, u, xn, 1e-3)
    finally:
        c.close()
    new = tr.weights()
    want = np.stack([oracle64.mlp_step(new, x[i], u[i]) for i in range(300)])
    old = np.stack([oracle64.mlp_step(mlp, x[i], u[i]) for i in range(300)])
    assert rel_err(got - x[:300], want - x[:300]) < 2e-2
    assert rel_err(got - x[:300], old - x[:300]) > 5e-2        # and it really is no longer the old model
    assert abs(l0 - losses[-1]) < 0.2 * losses[-1]


def test_train_needs_a_model():
    from mppi_tf_b200 import MppiError, _capi
    c = controller_from_cfg(make_cfg(128, 4, 6, 3))
    try:
        with pytest.raises(MppiError) as e:
            c.mlpTrainStep(np.zeros((4, 6)), np.zeros((4, 3)), np.zeros((4, 6)), 1e-3)
        assert e.value.code == _capi.MPPI_ERR_STATE
    finally:
        c.close()
