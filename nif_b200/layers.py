"""Derivative layers (mirror of nif/layers/gradient.py).

JacobianLayer(model, y_index, x_index)(x) -> (y, dy/dx[:, y_index, x_index]) exactly like the reference
(gradient.py:36-49, 207-231), but computed with forward-mode tangents carried through the fused kernel
instead of one reverse pass per output: one tangent direction per requested input column.  Directions on
ParameterNet inputs go through the trunk with torch's forward-mode AD to give the latent tangent."""
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

    @torch.no_grad()
    def __call__(self, x):
        m, n = self.model, self.model.net
        inp = m._dev(x)
        B = inp.shape[0]
        p_in = inp[:, : n.pi_dim].contiguous()
        xs = inp[:, n.pi_dim: n.pi_dim + n.si_dim].contiguous()
        nd = len(self.x_index)
        zdot = torch.zeros(nd, B, n.pi_hidden, device=inp.device)
        xdot = torch.zeros(nd, B, n.si_dim, device=inp.device)
        z = None
        for d, c in enumerate(self.x_index):
            if c < n.pi_dim:
                e = torch.zeros_like(p_in)
                e[:, c] = 1.0
                z, zd = torch.func.jvp(n._latent, (p_in,), (e,))
                zdot[d] = zd
            elif c < n.pi_dim + n.si_dim:
                xdot[d, :, c - n.pi_dim] = 1.0
            else:
                raise IndexError(f"x_index {c} outside the {n.pi_dim + n.si_dim} model inputs")
        if z is None:
            z = n._latent(p_in)
        packed = m._packed_weights()
        u, udot = n.engine.forward_tangent(z.contiguous(), xs, packed, zdot, xdot)  # udot [nd, B, so]
        J = udot.permute(1, 2, 0)[:, self.y_index, :]  # [B, |y|, |x|]
        return u, J.contiguous()

    call = __call__
