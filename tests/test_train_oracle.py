"""The numpy restatement of the learner step (oracle/train_oracle.py; learner_base.py:469-496) against an
independent automatic differentiation (torch.autograd, fp64) and torch's own Adam where the two formulations
coincide.  The reference holds no golden vector for a training step (parity unpinned)."""
import numpy as np
import pytest

from oracle.train_oracle import KEYS, AdamTrainer, loss_and_grads, normalise


def _problem(seed=0, n=257, s=6, a=3, H=128):
    rng = np.random.default_rng(seed)
    g = lambda i, o: rng.uniform(-1, 1, (i, o)) * np.sqrt(6.0 / (i + o))
    mlp = dict(W1=g(s + a, H), b1=0.1 * rng.standard_normal(H), W2=g(H, H), b2=0.1 * rng.standard_normal(H),
               W3=g(H, s), b3=0.05 * rng.standard_normal(s), Xmean=0.2 * rng.standard_normal(s + a),
               Xstd=1 + rng.random(s + a), Ymean=0.01 * rng.standard_normal(s), Ystd=0.05 + 0.1 * rng.random(s))
    x = rng.uniform(-1, 1, (n, s))
    u = rng.uniform(-1, 1, (n, a))
    # a learnable target: a fixed linear map of (x, u) plus a little noise
    Ax, Bu = 0.1 * rng.standard_normal((s, s)), 0.1 * rng.standard_normal((a, s))
    xn = x + x @ Ax + u @ Bu + 0.002 * rng.standard_normal((n, s))
    return mlp, x, u, xn


def test_unpinned_oracle_gradients_match_autograd():
    torch = pytest.importorskip("torch")
    mlp, x, u, xn = _problem()
    Xn, Yn = normalise(mlp, x, u, xn)
    w = {k: np.asarray(mlp[k], np.float64) for k in KEYS}
    loss, g = loss_and_grads(w, Xn, Yn)
    tw = {k: torch.tensor(w[k], dtype=torch.float64, requires_grad=True) for k in KEYS}
    tx, ty = torch.tensor(Xn), torch.tensor(Yn)
    out = torch.relu(torch.relu(tx @ tw["W1"] + tw["b1"]) @ tw["W2"] + tw["b2"]) @ tw["W3"] + tw["b3"]
    tl = torch.mean((out - ty) ** 2)
    tl.backward()
    assert abs(loss - tl.item()) < 1e-13 * max(1.0, abs(loss))
    for k in KEYS:
        np.testing.assert_allclose(g[k], tw[k].grad.numpy(), rtol=1e-10, atol=1e-14)


def test_adam_first_steps_follow_the_keras_formula():
    mlp, x, u, xn = _problem(1, n=64)
    tr = AdamTrainer(mlp)
    w0 = {k: tr.w[k].copy() for k in KEYS}
    Xn, Yn = normalise(mlp, x, u, xn)
    _, g = loss_and_grads(w0, Xn, Yn)
    tr.step(x, u, xn, 1e-3)
    # first step of Adam: m/(1-b1) = g, v/(1-b2) = g^2  ->  w1 = w0 - lr * g / (|g| + eps / sqrt(1 - b2)) up to the eps term
    for k in KEYS:
        lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
        want = w0[k] - lr_t * (0.1 * g[k]) / (np.sqrt(0.001 * g[k] ** 2) + 1e-7)
        np.testing.assert_allclose(tr.w[k], want, rtol=1e-12, atol=1e-15)


def test_training_reduces_the_loss():
    mlp, x, u, xn = _problem(2, n=512)
    tr = AdamTrainer(mlp)
    losses = [tr.step(x, u, xn, 1e-3) for _ in range(30)]
    assert losses[-1] < 0.7 * losses[0]
