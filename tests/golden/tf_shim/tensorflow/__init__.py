"""Minimal torch-backed stand-in for the handful of TensorFlow symbols that
`/root/reference/nif/{model,layers/*}.py` touch.

TEST INFRASTRUCTURE ONLY.  TensorFlow 2.11 cannot be installed in this image
(no network, no cp312 wheel), so `tests/golden/make_golden.py` puts this
directory on sys.path and imports the *unmodified* reference sources from
/root/reference on top of it.  Every arithmetic primitive is mapped to the
torch op of the same meaning (matmul, einsum, sin, sigmoid...), so the
reference's own slicing / reshaping / layer-wiring code is what produces the
golden vectors committed under tests/golden/.  Nothing in the product path
imports this.
"""
import sys
import types

import numpy as np
import torch

_DT = {
    "float64": torch.float64,
    "float32": torch.float32,
    "float16": torch.float16,
    "bfloat16": torch.bfloat16,
}

float32 = "float32"
float64 = "float64"
float16 = "float16"
bfloat16 = "bfloat16"


def _dt(d):
    if d is None:
        return None
    if isinstance(d, torch.dtype):
        return d
    if hasattr(d, "variable_dtype"):  # a Policy
        return _DT[d.variable_dtype]
    return _DT[str(d)]


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def cast(x, dtype, name=None):
    d = _dt(dtype)
    if isinstance(x, torch.Tensor):
        return x.to(d)
    return torch.as_tensor(x, dtype=d)


def reshape(x, shape, name=None):
    return torch.reshape(x, tuple(int(s) for s in shape))


def einsum(eq, *ops):
    return torch.einsum(eq, *ops)


def matmul(a, b):
    return torch.matmul(a, b)


def stack(xs, axis=0):
    return torch.stack(list(xs), dim=axis)


def gather(x, idx, axis=-1):
    idx = torch.as_tensor(list(idx), dtype=torch.long)
    return torch.index_select(x, axis if axis >= 0 else x.dim() + axis, idx)


def reduce_mean(x):
    return torch.mean(x)


def square(x):
    return x * x


def function(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


class _Variable(torch.Tensor):
    pass


def Variable(init, dtype=None, name=None, **kw):
    d = _dt(dtype)
    t = _t(init, d).detach().clone()
    t.requires_grad_(True)
    t._tf_name = name
    return t


class GradientTape:
    """Reverse-mode tape emulated with torch autograd (create_graph=True so the
    result is differentiable again, as TF's nested/persistent tapes are)."""

    def __init__(self, persistent=False):
        self.persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def watch(self, x):
        if not x.requires_grad:
            x.requires_grad_(True)

    def gradient(self, y, x):
        (g,) = torch.autograd.grad(y.sum(), x, create_graph=True, retain_graph=True)
        return g

    def batch_jacobian(self, y, x):
        # y: (B, m...) x: (B, n) -> (B, m..., n); rows are independent
        B = y.shape[0]
        yf = y.reshape(B, -1)
        cols = []
        for i in range(yf.shape[1]):
            (g,) = torch.autograd.grad(
                yf[:, i].sum(), x, create_graph=True, retain_graph=True
            )
            cols.append(g)
        j = torch.stack(cols, 1)
        return j.reshape(*y.shape, x.shape[-1])


# ---- tf.math ---------------------------------------------------------------
math = types.ModuleType("tensorflow.math")
math.sin = torch.sin
math.cos = torch.cos
math.sqrt = lambda x: torch.sqrt(_t(x, torch.float64) if not isinstance(x, torch.Tensor) else x)

# ---- tf.random -------------------------------------------------------------
random = types.ModuleType("tensorflow.random")


def _uniform(shape, minval=0.0, maxval=1.0, dtype="float32", **kw):
    d = _dt(dtype)
    lo = _t(minval, d)
    hi = _t(maxval, d)
    return torch.rand(tuple(shape), dtype=d) * (hi - lo) + lo


random.uniform = _uniform

from . import keras  # noqa: E402

initializers = keras.initializers

sys.modules[__name__ + ".math"] = math
sys.modules[__name__ + ".random"] = random
