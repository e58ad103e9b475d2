"""tf.keras subset (torch-backed). See ../__init__.py — test infrastructure only."""
import sys
import types

import torch

_DT = {
    "float64": torch.float64,
    "float32": torch.float32,
    "float16": torch.float16,
    "bfloat16": torch.bfloat16,
}


# ---- mixed precision --------------------------------------------------------
class Policy:
    def __init__(self, name):
        self.name = name
        if name.startswith("mixed_"):
            self.compute_dtype = name[len("mixed_") :]
            self.variable_dtype = "float32"
        else:
            self.compute_dtype = name
            self.variable_dtype = name


mixed_precision = types.ModuleType("tensorflow.keras.mixed_precision")
mixed_precision.Policy = Policy


def _policy(dtype):
    if isinstance(dtype, Policy):
        return dtype
    if dtype is None:
        return Policy("float32")
    return Policy(str(dtype))


# ---- activations -------------------------------------------------------------
def _swish(x):
    return x * torch.sigmoid(x)


_ACT = {
    None: lambda x: x,
    "linear": lambda x: x,
    "swish": _swish,
    "tanh": torch.tanh,
    "relu": torch.relu,
    "sigmoid": torch.sigmoid,
    "gelu": lambda x: torch.nn.functional.gelu(x),
    "elu": lambda x: torch.nn.functional.elu(x),
    "softplus": lambda x: torch.nn.functional.softplus(x),
}

activations = types.ModuleType("tensorflow.keras.activations")


def _get_act(a):
    if callable(a):
        return a
    return _ACT[a]


activations.get = _get_act


# ---- regularizers -------------------------------------------------------------
regularizers = types.ModuleType("tensorflow.keras.regularizers")


class L1:
    def __init__(self, l1=0.01):
        self.l1 = l1

    def __call__(self, w):
        return self.l1 * w.abs().sum()


class L2:
    def __init__(self, l2=0.01):
        self.l2 = l2

    def __call__(self, w):
        return self.l2 * (w * w).sum()


regularizers.L1 = L1
regularizers.L2 = L2
regularizers.Regularizer = object

# ---- initializers -------------------------------------------------------------
initializers = types.ModuleType("tensorflow.keras.initializers")


class TruncatedNormal:
    """Keras TruncatedNormal: N(mean, stddev) re-drawn outside +-2 stddev."""

    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev = mean, stddev

    def __call__(self, shape, dtype=None):
        d = _DT.get(str(dtype), torch.float32) if dtype is not None else torch.float32
        t = torch.empty(tuple(shape), dtype=d)
        torch.nn.init.trunc_normal_(
            t, self.mean, self.stddev, self.mean - 2 * self.stddev, self.mean + 2 * self.stddev
        )
        return t


initializers.TruncatedNormal = TruncatedNormal

# ---- layers -------------------------------------------------------------------
layers = types.ModuleType("tensorflow.keras.layers")


class Layer:
    def __init__(self, name=None, dtype=None, activity_regularizer=None, **kw):
        self.name = name or type(self).__name__
        self._policy = _policy(dtype)
        self.activity_regularizer = activity_regularizer
        self.losses = []

    def add_loss(self, v):
        self.losses.append(v)

    def __call__(self, *a, **kw):
        self.losses = []
        y = self.call(*a, **kw)
        if self.activity_regularizer is not None:
            # Keras divides activity regularisation by the batch size
            self.losses.append(self.activity_regularizer(y) / y.shape[0])
        return y

    def get_config(self):
        return {"name": self.name}


class Dense(Layer):
    def __init__(
        self,
        units,
        activation=None,
        kernel_initializer=None,
        bias_initializer=None,
        kernel_regularizer=None,
        bias_regularizer=None,
        activity_regularizer=None,
        dtype=None,
        name=None,
        **kw,
    ):
        super().__init__(name=name, dtype=dtype, activity_regularizer=activity_regularizer)
        self.units = units
        self.activation = _get_act(activation)
        self.kernel_initializer = kernel_initializer
        self.bias_initializer = bias_initializer
        self.kernel_regularizer = kernel_regularizer
        self.bias_regularizer = bias_regularizer
        self.kernel = None
        self.bias = None

    def build(self, in_dim):
        vd = self._policy.variable_dtype
        self.kernel = self.kernel_initializer((in_dim, self.units), dtype=vd).requires_grad_(True)
        self.bias = self.bias_initializer((self.units,), dtype=vd).requires_grad_(True)

    @property
    def weights(self):
        return [self.kernel, self.bias]

    @property
    def output_shape(self):
        return (None, self.units)

    def call(self, x):
        if self.kernel is None:
            self.build(x.shape[-1])
        cd = _DT[self._policy.compute_dtype]
        if self.kernel_regularizer is not None:
            self.add_loss(self.kernel_regularizer(self.kernel))
        if self.bias_regularizer is not None:
            self.add_loss(self.bias_regularizer(self.bias))
        y = torch.matmul(x.to(cd), self.kernel.to(cd)) + self.bias.to(cd)
        return self.activation(y)


class Dot(Layer):
    def __init__(self, axes, **kw):
        super().__init__(**kw)
        self.axes = axes

    def call(self, inputs):
        a, b = inputs
        assert tuple(self.axes) == (2, 1)
        return torch.einsum("bij,bj->bi", a, b)


def Input(shape=None, name=None, **kw):
    raise NotImplementedError("functional-API graph construction is not emulated")


layers.Layer = Layer
layers.Dense = Dense
layers.Dot = Dot
layers.Input = Input


class Model:
    def __init__(self, *a, **kw):
        raise NotImplementedError("functional-API graph construction is not emulated")


for _m in (mixed_precision, activations, regularizers, initializers, layers):
    sys.modules[_m.__name__] = _m
