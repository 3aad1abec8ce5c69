"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

CPU restatement (numpy, fp64) of one learner step for the MLP dynamics model:
LearnerBase._train_step (/root/reference/scripts/src/learners/learner_base.py:469-496) with the optimizer of
train_all (:325, tf.optimizers.Adam, Keras defaults beta1 = 0.9, beta2 = 0.999, epsilon = 1e-7), applied to the
network of orc_mlp_step (oracle/mppi_oracle_impl.h):

    Xn = (concat(x, u) - Xmean) / Xstd ;  Yn = ((x' - x) - Ymean) / Ystd          (normalised data)
    h1 = relu(Xn W1 + b1) ; h2 = relu(h1 W2 + b2) ; out = h2 W3 + b3                (model._predict_nn, :472)
    lossNorm = mean((out - Yn)^2)                                                   (:474-475)
    grads = d lossNorm / d weights ; optimizer.apply_gradients                      (:477-479)

Keras Adam (non-amsgrad), per tensor:  m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ;
    w <- w - lr sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps).

PARITY UNPINNED in the reference: TensorFlow (GradientTape, Keras) cannot run here and the reference holds no
golden vector for a training step.  tests/test_train_oracle.py checks the hand-written gradients below against
torch.autograd (an independent automatic differentiation) in fp64.
"""
import numpy as np

KEYS = ("W1", "b1", "W2", "b2", "W3", "b3")


def normalise(mlp, x, u, xn):
    s, a = x.shape[1], u.shape[1]
    one = lambda k, n, v: np.asarray(mlp.get(k, np.full(n, v)), np.float64)
    Xn = (np.concatenate([x, u], 1) - one("Xmean", s + a, 0.0)) / one("Xstd", s + a, 1.0)
    Yn = ((xn - x) - one("Ymean", s, 0.0)) / one("Ystd", s, 1.0)
    return Xn, Yn


def loss_and_grads(w, Xn, Yn):
    """w: dict of fp64 arrays (Keras layout [in][out]).  Returns (lossNorm, grads dict)."""
    p1 = Xn @ w["W1"] + w["b1"]
    h1 = np.maximum(p1, 0.0)
    p2 = h1 @ w["W2"] + w["b2"]
    h2 = np.maximum(p2, 0.0)
    out = h2 @ w["W3"] + w["b3"]
    diff = out - Yn
    loss = np.mean(diff * diff)
    d_out = 2.0 * diff / diff.size
    g = {"W3": h2.T @ d_out, "b3": d_out.sum(0)}
    d_h2 = (d_out @ w["W3"].T) * (h2 > 0)
    g["W2"], g["b2"] = h1.T @ d_h2, d_h2.sum(0)
    d_h1 = (d_h2 @ w["W2"].T) * (h1 > 0)
    g["W1"], g["b1"] = Xn.T @ d_h1, d_h1.sum(0)
    return loss, g


class AdamTrainer:
    """State of the optimizer across steps, as tf.optimizers.Adam keeps it."""

    def __init__(self, mlp, beta1=0.9, beta2=0.999, epsilon=1e-7):
        self.mlp = dict(mlp)
        self.w = {k: np.array(mlp[k], np.float64) for k in KEYS}
        self.m = {k: np.zeros_like(self.w[k]) for k in KEYS}
        self.v = {k: np.zeros_like(self.w[k]) for k in KEYS}
        self.t, self.b1, self.b2, self.eps = 0, beta1, beta2, epsilon

    def step(self, x, u, xn, lr):
        """One full-batch step; returns the loss before the update."""
        Xn, Yn = normalise(self.mlp, np.asarray(x, np.float64), np.asarray(u, np.float64), np.asarray(xn, np.float64))
        loss, g = loss_and_grads(self.w, Xn, Yn)
        self.t += 1
        lr_t = lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k in KEYS:
            self.m[k] = self.b1 * self.m[k] + (1.0 - self.b1) * g[k]
            self.v[k] = self.b2 * self.v[k] + (1.0 - self.b2) * g[k] * g[k]
            self.w[k] = self.w[k] - lr_t * self.m[k] / (np.sqrt(self.v[k]) + self.eps)
        return loss

    def weights(self):
        out = dict(self.mlp)
        out.update({k: self.w[k].copy() for k in KEYS})
        return out
