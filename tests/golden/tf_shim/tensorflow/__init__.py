"""Minimal numpy-backed stand-in for the handful of TensorFlow ops the reference's Python twin
(/root/reference/scripts/src) executes on the MPPI update path.  TEST TOOLING ONLY: it exists so that
tests/golden/gen_python_twin_fixtures.py can run the REFERENCE'S OWN Python code (its op sequence,
its constants, its shift/next logic) here, where TensorFlow cannot be installed, and record golden
vectors.  Each function implements the documented semantics of the TF op of the same name for the
argument patterns the reference uses; nothing here is imported by the product or by the tests."""
import contextlib
import types

import numpy as np

float64, float32, int32, int64 = np.float64, np.float32, np.int32, np.int64


class Variable(np.ndarray):
    def __new__(cls, initial_value, trainable=False, dtype=None, name=None):
        return np.asarray(initial_value, dtype=dtype).view(cls).copy()

    def assign(self, v):
        self[...] = v
        return self

    def numpy(self):
        return np.asarray(self)


class Module:
    pass


def _a(x, dtype=None):
    return np.asarray(x, dtype=dtype)


def convert_to_tensor(x, dtype=None, name=None):
    return np.array(x, dtype=dtype)


constant = convert_to_tensor


def is_tensor(x):
    return isinstance(x, np.ndarray)


def zeros(shape, dtype=float32, name=None):
    return np.zeros(tuple(int(s) for s in np.atleast_1d(shape)), dtype=dtype)


def ones(shape, dtype=float32, name=None):
    return np.ones(tuple(int(s) for s in np.atleast_1d(shape)), dtype=dtype)


def shape(x):
    return np.array(np.shape(x))


def cast(x, dtype=None):
    return np.asarray(x, dtype=dtype)


def add(a, b, name=None):
    return _a(a) + _a(b)


def divide(a, b, name=None):
    return _a(a) / _a(b)


realdiv = divide


def expand_dims(x, axis, name=None):
    return np.expand_dims(_a(x), axis)


def broadcast_to(x, shape, name=None):
    return np.broadcast_to(_a(x), tuple(int(s) for s in shape)).copy()


def squeeze(x, axis=None, name=None):
    return np.squeeze(_a(x), axis=axis)


def slice(x, begin, size, name=None):          # noqa: A001  (tf.slice)
    x = _a(x)
    idx = tuple(np.s_[b:(None if s == -1 else b + s)] for b, s in zip(begin, size))
    return x[idx]


def concat(values, axis, name=None):
    # rank-0 operands are taken as one-element vectors: AUVModel.get_inertial concatenates scalar Variables
    # (auv_model.py:274-276), which only makes sense as building the rows of the inertia matrix
    return np.concatenate([np.atleast_1d(_a(v)) for v in values], axis=axis)


def tensordot(a, b, axes, name=None):
    return np.tensordot(_a(a), _a(b), axes)


def gather(params, indices, axis=0, name=None):
    return np.take(_a(params), indices, axis=axis)


def norm(x, ord="euclidean", axis=None, keepdims=False, name=None):     # noqa: A002
    return np.linalg.norm(_a(x), axis=axis, keepdims=keepdims)


def reduce_min(x, axis=None, name=None):
    return np.min(_a(x), axis=axis)


def reduce_max(x, axis=None, name=None):
    return np.max(_a(x), axis=axis)


def reduce_sum(x, axis=None, name=None):
    return np.sum(_a(x), axis=axis)


def clip_by_value(x, lo, hi, name=None):
    return np.clip(_a(x), lo, hi)


def eye(n, dtype=float64, name=None):
    return np.eye(n, dtype=dtype)


def multiply(a, b, name=None):
    return _a(a) * _a(b)


def subtract(a, b, name=None):
    return _a(a) - _a(b)


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    return _matmul(a, b, transpose_a, transpose_b)


def transpose(x, perm=None, name=None):
    return np.transpose(_a(x), perm)


def sqrt(x, name=None):
    return np.sqrt(_a(x))


def pow(x, y, name=None):                      # noqa: A001  (tf.pow)
    return np.power(_a(x), y)


def abs(x, name=None):                         # noqa: A001  (tf.abs)
    return np.abs(_a(x))


@contextlib.contextmanager
def name_scope(name):
    yield name


def function(f=None, **kw):
    return f if f is not None else (lambda g: g)


def _matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _a(a), _a(b)
    if transpose_a:
        a = np.swapaxes(a, -1, -2)
    if transpose_b:
        b = np.swapaxes(b, -1, -2)
    return np.matmul(a, b)


def _diag(v, name=None):
    v = _a(v)
    if v.ndim == 1:
        return np.diag(v)
    out = np.zeros(v.shape + (v.shape[-1],), v.dtype)           # batched: [..., n] -> [..., n, n]
    idx = np.arange(v.shape[-1])
    out[..., idx, idx] = v
    return out


def _l2_normalize(x, axis=None, epsilon=1e-12, name=None):
    x = _a(x)
    return x / np.sqrt(np.maximum(np.sum(x * x, axis=axis, keepdims=True), epsilon))


linalg = types.SimpleNamespace(matmul=_matmul, inv=lambda m, name=None: np.linalg.inv(_a(m)), diag=_diag,
                               cross=lambda a, b, name=None: np.cross(_a(a), _a(b)),
                               norm=lambda x, axis=None, keepdims=False, name=None: np.linalg.norm(_a(x), axis=axis, keepdims=keepdims),
                               normalize=lambda x, ord="euclidean", axis=None, name=None: (
                                   _a(x) / np.linalg.norm(_a(x), axis=axis, keepdims=True),
                                   np.linalg.norm(_a(x), axis=axis, keepdims=True)))
math = types.SimpleNamespace(
    subtract=lambda a, b, name=None: _a(a) - _a(b), multiply=lambda a, b, name=None: _a(a) * _a(b),
    add=add, exp=lambda x, name=None: np.exp(_a(x)), reduce_sum=reduce_sum, reduce_min=reduce_min,
    reduce_max=reduce_max, divide=divide, l2_normalize=_l2_normalize, acos=lambda x, name=None: np.arccos(_a(x)),
    abs=lambda x, name=None: np.abs(_a(x)))


def _normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    return np.random.default_rng(seed).normal(mean, stddev, tuple(int(s) for s in shape)).astype(dtype)


random = types.SimpleNamespace(normal=_normal)
optimizers = types.SimpleNamespace(Adam=lambda **kw: None)
config = types.SimpleNamespace(experimental=types.SimpleNamespace(
    list_physical_devices=lambda kind=None: ["shim:0"], set_memory_growth=lambda dev, flag: None))
summary = types.SimpleNamespace(create_file_writer=lambda *a, **k: None, scalar=lambda *a, **k: None,
                                histogram=lambda *a, **k: None)
profiler = types.SimpleNamespace(experimental=types.SimpleNamespace(start=lambda *a, **k: None,
                                                                    stop=lambda *a, **k: None))


# ---- tf.keras: just enough for NNModel / NNAUVModel (scripts/src/models/nn_model.py:54-60): Sequential of Dense layers,
# forward only.  Weights are created on first use (Glorot uniform like Keras) and are plain numpy arrays the fixture
# generator overwrites.
class _Dense:
    def __init__(self, units, activation=None, input_shape=None, kernel_regularizer=None, name=None):
        self.units, self.activation, self.input_shape, self.name = int(units), activation, input_shape, name
        self.kernel, self.bias = None, None

    def build(self, n_in, rng):
        lim = np.sqrt(6.0 / (n_in + self.units))
        self.kernel = rng.uniform(-lim, lim, (n_in, self.units))
        self.bias = np.zeros(self.units)

    def __call__(self, x):
        y = _a(x) @ self.kernel + self.bias
        return np.maximum(y, 0.0) if self.activation == "relu" else y


class _Sequential:
    def __init__(self, layers):
        self.layers = list(layers)
        rng = np.random.default_rng(0)
        n_in = int(self.layers[0].input_shape[0])
        for layer in self.layers:
            layer.build(n_in, rng)
            n_in = layer.units

    def __call__(self, x):
        for layer in self.layers:
            x = layer(x)
        return x

    @property
    def trainable_variables(self):
        out = []
        for layer in self.layers:
            out += [layer.kernel, layer.bias]
        return out


keras = types.SimpleNamespace(Sequential=_Sequential, layers=types.SimpleNamespace(Dense=_Dense),
                              backend=types.SimpleNamespace(set_floatx=lambda name: None),
                              models=types.SimpleNamespace(load_model=lambda path: None))


def constant(x, dtype=None, name=None):
    return _a(x, dtype)


compat = types.SimpleNamespace(v1=types.SimpleNamespace(dimension_value=lambda d: d))
