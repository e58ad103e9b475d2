"""Derivative layers (mirror of nif/layers/gradient.py).

JacobianLayer(model, y_index, x_index)(x) -> (y, dy/dx[:, y_index, x_index]) exactly like the reference
(gradient.py:36-49, 207-231), but computed with forward-mode tangents carried through the fused kernel
instead of one reverse pass per output: one tangent direction per requested input column.  Directions on
ParameterNet inputs go through the trunk with torch's forward-mode AD to give the latent tangent.

HessianLayer(model, y_index, x_index)(x) -> (y, dy/dx, d2y/dx2) (gradient.py:130-180, 234-261): second-order forward
mode, one launch of the fused kernel per unordered pair of requested input columns."""
from __future__ import annotations

from typing import Sequence

import torch

from ._lib import NifError


class JacobianLayer:
    def __init__(self, model, y_index: Sequence[int], x_index: Sequence[int], **_kw):
        if getattr(model, "kind", None) != "full":
            raise NifError("JacobianLayer wraps the full model (as returned by NIF.build()/model())")
        self.model = model
        self.y_index = list(y_index)
        self.x_index = list(x_index)

    def as_model(self):
        """The Keras model tutorial 8 builds around this layer (tutorial/8_...ipynb:809-813):
        output = concat([y, reshape(dy/dx, (-1, |y|*|x|))], -1).  It shares the wrapped model's variables, predicts
        with the forward-mode kernel and trains with `SobolevMSE` through the reverse-over-forward kernels."""
        from .keras_like import Model
        m = Model(self.model.net, "jacobian")
        m.jac_y, m.jac_x = self.y_index, self.x_index
        return m

    @torch.no_grad()
    def __call__(self, x):
        m = self.model
        return m._jacobian_forward(m._dev(x), self.y_index, self.x_index)

    call = __call__


class HessianLayer:
    """compute_output_and_grad_and_hessian (nif/layers/gradient.py:234-261): returns (y, J, H) with
    J[b, a, c] = d y[b, y_index[a]] / d x[b, x_index[c]] and H[b, a, c, e] the matching second derivatives."""

    def __init__(self, model, y_index: Sequence[int], x_index: Sequence[int], **_kw):
        if getattr(model, "kind", None) != "full":
            raise NifError("HessianLayer wraps the full model (as returned by NIF.build()/model())")
        self.model = model
        self.y_index = list(y_index)
        self.x_index = list(x_index)

    @torch.no_grad()
    def __call__(self, x):
        m = self.model
        return m._hessian_forward(m._dev(x), self.y_index, self.x_index)

    call = __call__
