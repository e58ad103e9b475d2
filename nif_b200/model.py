"""Host-side mirror of `nif/model.py` (pswpswpsw/nif): NIF and NIFMultiScale.

Same constructor arguments, cfg-dict schema, method names, validation errors and variable names
as the reference (nif/model.py:48-480, 483-986), but nothing here builds a graph: the objects own
one flat fp32 parameter buffer in HBM and hand the hot path to libnif_b200.so.  The (B, po_dim)
tensor `pnet_output` of the reference is never formed unless the caller explicitly asks for it
through model_p_to_w() / model_lr_to_w().
"""
from __future__ import annotations

import json
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from ._lib import ACT, NifError
from .keras_like import Model
from .ops import FusedShapeNet

__all__ = ["NIF", "NIFMultiScale", "NIFMultiScaleLastLayerParameterized"]

_POLICIES = ("float32", "mixed_float16", "mixed_bfloat16")


def _trunc_normal(shape, std, gen):
    t = torch.empty(shape, dtype=torch.float64)
    torch.nn.init.trunc_normal_(t, 0.0, std, -2 * std, 2 * std, generator=gen)
    return t.float()


def _uniform(shape, bound, gen):
    b = torch.as_tensor(bound, dtype=torch.float64)
    return ((torch.rand(shape, dtype=torch.float64, generator=gen) * 2 - 1) * b).float()


class _TrunkLinear(torch.autograd.Function):
    """y = h @ W + b of a ParameterNet Dense layer for the paths that differentiate the trunk once.  Same arithmetic as
    the plain torch expression under the model's policy (bf16 operands and products under mixed_bfloat16, fp32 bias
    add), but the bias gradient is a GEMV with a row of ones instead of torch's column reduction (69 -> 12 microseconds
    per layer at 65 536 x 128) and the bf16 operand copies are made once and kept for the reverse pass."""

    @staticmethod
    def forward(ctx, h, W, b, bf16):
        if bf16:
            h, W = h.to(torch.bfloat16), W.to(torch.bfloat16)
        ctx.save_for_backward(h, W)
        with torch.autocast("cuda", enabled=False):
            return (h @ W).float() + b

    @staticmethod
    def backward(ctx, dy):
        h, W = ctx.saved_tensors
        g = dy.to(h.dtype)
        dh = (g @ W.t()).float() if ctx.needs_input_grad[0] else None
        dW = (h.t() @ g).float()
        db = (dy.new_ones(1, dy.shape[0]) @ dy.float()).reshape(-1)
        return dh, dW, db, None


class _WideTrunk(torch.autograd.Function):
    """The whole MLP trunk -- first Dense, l_st x MLP_SimpleShortCut, bottleneck Dense (nif/model.py:326-343,
    nif/layers/mlp.py:148-160) -- of a wide ParameterNet (units > 64) under mixed_bfloat16, for the paths that
    differentiate it once.  The Dense products are library GEMMs on bf16 operands; everything between two GEMMs of a layer
    (cast, bias, activation, shortcut, cast; the reverse pass's activation derivative, bias gradient, shortcut sum) is one
    kernel of the library (nif_trunk_ew_forward / nif_trunk_ew_backward) with the rounding points of _TrunkLinear: bf16 at
    the GEMM operands and products, fp32 elsewhere.  Keeps two bf16 tensors per layer (operand and product) for the reverse
    pass; the pre-activation is recomputed from the product."""

    @staticmethod
    def forward(ctx, p_in, act, *wb):  # wb = (W_first, b_first, W_hidden_0, b_hidden_0, ..., W_bottleneck, b_bottleneck)
        from . import _lib
        from .ops import _f32c, _stream
        L = _lib.lib()
        B = p_in.shape[0]
        bf = torch.bfloat16
        with torch.autocast("cuda", enabled=False):
            Ws = [w.to(bf) for w in wb[0::2]]
            bs = [_f32c(b, "bias") for b in wb[1::2]]
            n = Ws[0].shape[1]
            ins, ys, h32 = [p_in.to(bf)], [], None
            for i in range(len(Ws) - 1):
                y = ins[-1] @ Ws[i]
                h_out = torch.empty(B, n, dtype=torch.float32, device=p_in.device)
                h16 = torch.empty(B, n, dtype=bf, device=p_in.device)
                _lib.check(L.nif_trunk_ew_forward(B, n, act, y.data_ptr(), bs[i].data_ptr(),
                                                  None if h32 is None else h32.data_ptr(), h_out.data_ptr(), h16.data_ptr(),
                                                  _stream()), "nif_trunk_ew_forward")
                ys.append(y)
                ins.append(h16)
                h32 = h_out
            z = (ins[-1] @ Ws[-1]).float() + bs[-1]
        ctx.save_for_backward(*ins, *ys, *Ws, *bs[:-1])
        ctx.act, ctx.nl = act, len(Ws)
        return z

    @staticmethod
    def backward(ctx, dz):
        from . import _lib
        from .ops import _stream
        L = _lib.lib()
        nl, act = ctx.nl, ctx.act
        t = ctx.saved_tensors
        ins, ys, Ws, bs = t[:nl], t[nl: 2 * nl - 1], t[2 * nl - 1: 3 * nl - 1], t[3 * nl - 1:]
        B, n = ys[0].shape
        dev = dz.device
        grads = [None] * (2 * nl)
        with torch.autocast("cuda", enabled=False):
            dz = dz.float()
            g = dz.to(torch.bfloat16)
            grads[2 * nl - 2] = (ins[-1].t() @ g).float()
            grads[2 * nl - 1] = (dz.new_ones(1, B) @ dz).reshape(-1)
            pend = g @ Ws[-1].t()
            ws = torch.empty(int(L.nif_trunk_ew_ws_floats(n)), dtype=torch.float32, device=dev)
            dh = None
            for i in range(nl - 2, -1, -1):
                gi = torch.empty(B, n, dtype=torch.bfloat16, device=dev)
                db = torch.empty(n, dtype=torch.float32, device=dev)
                dh_out = None if i == 0 else (dh if dh is not None else torch.empty(B, n, dtype=torch.float32, device=dev))
                _lib.check(L.nif_trunk_ew_backward(B, n, act, ys[i].data_ptr(), bs[i].data_ptr(),
                                                   None if dh is None else dh.data_ptr(), pend.data_ptr(),
                                                   None if dh_out is None else dh_out.data_ptr(), gi.data_ptr(),
                                                   db.data_ptr(), ws.data_ptr(), _stream()), "nif_trunk_ew_backward")
                dh = dh_out
                grads[2 * i] = (ins[i].t() @ gi).float()
                grads[2 * i + 1] = db
                if i > 0:
                    pend = gi @ Ws[i].t()
        return (None, None, *grads)


def _wide_trunk_ok(net, input_p: torch.Tensor, units: int, act) -> bool:
    """_WideTrunk serves: the once-differentiated training step of a mixed_bfloat16 model on the GPU whose MLP trunk is
    wider than the tensor-core trunk kernels go (they stop at 64 units)."""
    return (net._first_order_only and getattr(net, "_use_wide_trunk", True) and input_p.is_cuda
            and torch.is_autocast_enabled() and units > 64 and units <= 256
            and units % 8 == 0 and 2048 % units == 0 and act in ACT and act != "sine")


class NIF(object):
    """Neural Implicit Flow with a swish/tanh/... ShapeNet with residual hidden layers
    (reference: class NIF, nif/model.py:48-480).

    Args follow the reference: cfg_shape_net {input_dim, output_dim, units, nlayers, activation},
    cfg_parameter_net {input_dim, latent_dim, units, nlayers, activation, [l1_reg|l2_reg]},
    mixed_policy (Keras policy name).  Extra keyword-only arguments are ours: `seed` for the
    initialisers and `device`.
    """

    _variant = "nif"

    def __init__(self, cfg_shape_net, cfg_parameter_net, mixed_policy="float32", *, seed: Optional[int] = None,
                 device=None, compute: str = "auto"):
        self.cfg_shape_net = cfg_shape_net
        self.cfg_parameter_net = cfg_parameter_net
        self._validate_cfg(cfg_shape_net, cfg_parameter_net)
        self.si_dim = cfg_shape_net["input_dim"]
        self.so_dim = cfg_shape_net["output_dim"]
        self.n_sx = cfg_shape_net["units"]
        self.l_sx = cfg_shape_net["nlayers"]
        self.pi_dim = cfg_parameter_net["input_dim"]
        self.pi_hidden = cfg_parameter_net["latent_dim"]
        self.n_st = cfg_parameter_net["units"]
        self.l_st = cfg_parameter_net["nlayers"]
        # regularisation knobs (nif/model.py:95-125)
        self.p_jac_reg = cfg_parameter_net.get("jac_reg", None)
        self.p_l1_reg = cfg_parameter_net.get("l1_reg", None)
        self.p_l2_reg = cfg_parameter_net.get("l2_reg", None)
        self.p_act_l1_reg = cfg_parameter_net.get("act_l1_reg", None)
        self.p_act_l2_reg = cfg_parameter_net.get("act_l2_reg", None)
        if mixed_policy not in _POLICIES:
            raise ValueError(f"mixed_policy must be one of {_POLICIES} (float64 has no GPU path)")
        if compute not in ("auto", "fp32", "fp16x3", "bf16"):
            raise ValueError("compute must be 'auto', 'fp32' (CUDA cores), 'fp16x3' (tensor cores, fp32-grade) or "
                             "'bf16' (tensor cores, bf16 operands)")
        self._compute_request = compute
        self.mixed_policy_name = mixed_policy
        self.variable_Dtype = "float32"
        # mixed_bfloat16 (nif/model.py:101-105): bf16 operands on the tensor cores, fp32 accumulation, fp32 variables.
        # mixed_float16 is served at fp32 grade (the FP16x3 path already runs on the fp16 tensor-core pipe).
        self.compute_Dtype = "bfloat16" if mixed_policy == "mixed_bfloat16" else "float32"
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        self._gen = torch.Generator().manual_seed(int(seed) if seed is not None else int(torch.seed() % (2**31)))
        self.po_dim = self._po_dim()
        self._engine = None
        self._init_parameters()

    # ---- configuration ---------------------------------------------------------------------------
    def _validate_cfg(self, cfg_shape_net, cfg_parameter_net):
        if not isinstance(cfg_parameter_net, dict):
            raise TypeError("cfg_parameter_net must be a dictionary")
        if not isinstance(cfg_shape_net, dict):
            raise TypeError("cfg_shape_net must be a dictionary")
        act = cfg_shape_net.get("activation")
        if act not in ACT:
            raise ValueError(f"cfg_shape_net['activation']={act!r} is not supported by the fused kernels {sorted(k for k in ACT if k)}")

    def _po_dim(self) -> int:
        # nif/model.py:169-173
        return self.l_sx * self.n_sx**2 + (self.si_dim + self.so_dim + 1 + self.l_sx) * self.n_sx + self.so_dim

    @property
    def _omega0(self) -> float:
        return 1.0

    @property
    def engine(self) -> FusedShapeNet:
        if self._engine is None:
            # 'auto': the tensor-core path (same parity gates) where the library builds it, else the CUDA-core path;
            # inside the library, shapes the tensor-core kernels do not cover fall through to the CUDA-core kernels.
            comp = self._compute_request
            if comp == "auto":
                if self.mixed_policy_name == "mixed_bfloat16" and self.n_sx > 32 and self._variant != "siren_res":
                    comp = "bf16"
                else:
                    comp = "fp16x3" if (32 < self.n_sx <= 64 and self.pi_hidden >= 1) else "fp32"
            self._engine = FusedShapeNet(self._variant, self.si_dim, self.so_dim, self.n_sx, self.l_sx, self.pi_hidden,
                                         self.cfg_shape_net.get("activation"), self._omega0, compute=comp)
            if self._engine.po_dim != self.po_dim:
                raise NifError("po_dim mismatch between host and library")
        return self._engine

    # ---- parameters --------------------------------------------------------------------------------
    def _trunk_layout(self) -> List[Tuple[str, Tuple[int, ...]]]:
        """[(variable name, shape)] in creation order; names follow nif/model.py:186-229."""
        L = [("first_dense_pnet/kernel", (self.pi_dim, self.n_st)), ("first_dense_pnet/bias", (self.n_st,))]
        for i in range(self.l_st):
            q = f"hidden_mlpshortcut_pnet_{i}"
            L += [(q + "/kernel", (self.n_st, self.n_st)), (q + "/bias", (self.n_st,))]
        L += [("bottleneck_pnet/kernel", (self.n_st, self.pi_hidden)), ("bottleneck_pnet/bias", (self.pi_hidden,))]
        return L

    _last_names = ("last_pnet/kernel", "last_pnet/bias")

    def _draw(self, name: str, shape) -> torch.Tensor:
        # every Dense of class NIF: TruncatedNormal(stddev=0.1) kernel and bias (nif/model.py:181-182, 222-223)
        return _trunc_normal(shape, 0.1, self._gen)

    def _fused_trunk_supported(self) -> bool:
        """Dense(act) -> l_st x MLP_SimpleShortCut -> Dense(latent) with a non-sine activation and <= 64 units is
        what libnif_b200's trunk kernels implement; other trunks run as torch ops."""
        p = self.cfg_parameter_net
        return (p.get("activation") in ("swish", "tanh", "relu", "sigmoid", "linear", None) and self.n_st <= 64
                and not bool(p.get("use_resblock", False)) and self.pi_dim <= 8
                and not isinstance(self.p_jac_reg, (float, int)))  # the Jacobian regulariser differentiates the trunk twice

    def _place(self, layout):
        """Offsets of every variable in the flat buffer.  With the fused trunk, the trunk variables are laid out
        kernels-first, unpadded, in the column order the trunk kernels read (include/nif_b200.h), so that the trunk
        weight vector and its gradient are plain views of the flat buffers."""
        names = [n for n, _ in layout]
        shapes = dict(layout)
        place, off = {}, 0
        n_extra = getattr(self, "_n_extra", 0)
        head = names[len(names) - n_extra - 2: len(names) - n_extra]
        extra_names = names[len(names) - n_extra:] if n_extra else []
        trunk = names[: len(names) - n_extra - 2]
        if self._fused_trunk_supported():
            order = [n for n in trunk if n.endswith("/kernel")] + [n for n in trunk if n.endswith("/bias")]
            for n in order:
                place[n] = off
                off += int(np.prod(shapes[n]))
            self._n_trunk = off
            off = (off + 3) // 4 * 4
        else:
            self._n_trunk = 0
            for n in trunk:
                place[n] = off
                off += (int(np.prod(shapes[n])) + 3) // 4 * 4
        for n in head + extra_names:  # the last linear layer (and what follows): 16-byte aligned for the kernels
            place[n] = off
            off += (int(np.prod(shapes[n])) + 3) // 4 * 4
        return place, off

    def _bind_views(self):
        self._views: Dict[str, torch.Tensor] = {}
        self._gviews: Dict[str, torch.Tensor] = {}
        for name, (o, shape) in self._layout.items():
            n = int(np.prod(shape))
            v = self.theta[o:o + n].view(shape)
            v.requires_grad_(True)
            g = self.grad[o:o + n].view(shape)
            v.grad = g
            self._views[name], self._gviews[name] = v, g
        self._trunk = None
        if self._n_trunk and self.device.type == "cuda":
            from .ops import FusedTrunk
            self._trunk = FusedTrunk(self.pi_dim, self.pi_hidden, self.n_st, self.l_st,
                                     self.cfg_parameter_net.get("activation") or "linear")
            if self._trunk.n_theta != self._n_trunk:
                raise NifError("trunk layout mismatch between host and library")

    def _rebind(self, theta: torch.Tensor, grad: torch.Tensor):
        """Adopt caller-provided flat buffers (data parallel: buffers in NVLink symmetric memory)."""
        self.theta, self.grad = theta, grad
        self._bind_views()

    @property
    def theta_trunk(self) -> torch.Tensor:
        return self.theta[: self._n_trunk]

    @property
    def grad_trunk(self) -> torch.Tensor:
        return self.grad[: self._n_trunk]

    def _extra_layout(self) -> List[Tuple[str, Tuple[int, ...]]]:
        """Variables created after the ParameterNet (none for the hyper-network classes)."""
        return []

    def _init_parameters(self):
        extra = self._extra_layout()
        layout = self._trunk_layout() + [(self._last_names[0], (self.pi_hidden, self.po_dim)),
                                         (self._last_names[1], (self.po_dim,))] + extra
        self._n_extra = len(extra)
        place, total = self._place(layout)
        self._layout: Dict[str, Tuple[int, Tuple[int, ...]]] = {name: (place[name], shape) for name, shape in layout}
        self.n_flat = total
        host = torch.zeros(total, dtype=torch.float32)
        for name, shape in layout:  # draws happen in creation order, independent of the placement
            o, _ = self._layout[name]
            host[o:o + int(np.prod(shape))] = self._draw(name, shape).reshape(-1)
        self.theta = host.to(self.device)
        self.grad = torch.zeros_like(self.theta)
        self._bind_views()

    @property
    def variables(self) -> Dict[str, torch.Tensor]:
        """name -> tensor view into the flat parameter buffer (shared by every derived model, like the
        shared layer objects of the reference, tutorial/1_simple_1d_wave.ipynb cell 22)."""
        return self._views

    @property
    def trainable_variables(self) -> List[torch.Tensor]:
        return list(self._views.values())

    def count_params(self) -> int:
        return sum(int(np.prod(s)) for _, s in self._layout.values())

    def to(self, device):
        device = torch.device(device)
        if device == self.device:
            return self
        data = {k: v.detach().cpu().clone() for k, v in self._views.items()}
        self.device = device
        self.theta = torch.zeros(self.n_flat, dtype=torch.float32, device=device)
        self.grad = torch.zeros_like(self.theta)
        for name, (o, shape) in self._layout.items():
            self.theta[o:o + int(np.prod(shape))] = data[name].reshape(-1).to(device)
        self._bind_views()
        self._engine = None
        return self

    def set_weights(self, named: Dict[str, "np.ndarray"]):
        with torch.no_grad():
            for k, a in named.items():
                if k not in self._views:
                    raise KeyError(f"unknown variable {k!r}")
                t = torch.as_tensor(np.asarray(a), dtype=torch.float32)
                if tuple(t.shape) != tuple(self._views[k].shape):
                    raise ValueError(f"{k}: shape {tuple(t.shape)} != {tuple(self._views[k].shape)}")
                self._views[k].copy_(t.to(self.device))

    def get_weights(self) -> Dict[str, "np.ndarray"]:
        return {k: v.detach().cpu().numpy().copy() for k, v in self._views.items()}

    # ---- ParameterNet trunk (everything before the last linear; plain torch ops) ----------------------
    # set around calls that differentiate the trunk once only (the plain training step): swish runs as the single silu
    # kernel (one forward, one backward instead of two and four); silu's backward has no forward-mode rule, so the paths
    # that differentiate the trunk twice (jac_reg, Hessians, du/dt in the loss) keep the composite form
    _first_order_only = False

    def _act(self, name):
        if name == "swish":
            if self._first_order_only:
                return torch.nn.functional.silu
            return lambda v: v * torch.sigmoid(v)
        if name == "tanh":
            return torch.tanh
        if name == "relu":
            return torch.relu
        if name == "sigmoid":
            return torch.sigmoid
        if name in (None, "linear"):
            return lambda v: v
        raise ValueError(f"ParameterNet activation {name!r} is not supported")

    def _lin(self, h, W, b):
        if self._first_order_only and h.is_cuda:
            return _TrunkLinear.apply(h, W, b, torch.is_autocast_enabled())
        return h @ W + b

    def _latent(self, input_p: torch.Tensor) -> torch.Tensor:
        """_call_parameter_net up to the bottleneck (nif/model.py:326-343; MLP_SimpleShortCut mlp.py:148-160)."""
        V = self._views
        if _wide_trunk_ok(self, input_p, int(self.cfg_parameter_net["units"]), self.cfg_parameter_net["activation"]):
            names = ["first_dense_pnet"] + [f"hidden_mlpshortcut_pnet_{i}" for i in range(self.l_st)] + ["bottleneck_pnet"]
            wb = [V[q + s] for q in names for s in ("/kernel", "/bias")]
            return _WideTrunk.apply(input_p, ACT[self.cfg_parameter_net["activation"]], *wb)
        f = self._act(self.cfg_parameter_net["activation"])
        h = f(self._lin(input_p, V["first_dense_pnet/kernel"], V["first_dense_pnet/bias"]))
        for i in range(self.l_st):
            q = f"hidden_mlpshortcut_pnet_{i}"
            h = h + f(self._lin(h, V[q + "/kernel"], V[q + "/bias"]))
        return self._lin(h, V["bottleneck_pnet/kernel"], V["bottleneck_pnet/bias"])

    @property
    def w_h(self) -> torch.Tensor:
        return self._views[self._last_names[0]]

    @property
    def b_h(self) -> torch.Tensor:
        return self._views[self._last_names[1]]

    def _jac_reg_loss(self, input_p: torch.Tensor) -> torch.Tensor:
        """JacRegLatentLayer (nif/layers/gradient.py:52-113) as build() wires it (nif/model.py:353-375, y_index = every
        latent unit, x_index = every ParameterNet input):  jac_reg * mean_{b,k,c} (d latent[b,k] / d input_p[b,c])^2.
        One forward-mode tangent per ParameterNet input; the result is differentiable w.r.t. the trunk variables
        (reverse over forward), which is how its gradient reaches the flat gradient buffer."""
        sq = input_p.new_zeros(())
        for c in range(self.pi_dim):
            e = torch.zeros_like(input_p)
            e[:, c] = 1.0
            _, zd = torch.func.jvp(self._latent, (input_p,), (e,))
            sq = sq + (zd * zd).sum()
        return float(self.p_jac_reg) * sq / (input_p.shape[0] * self.pi_hidden * self.pi_dim)

    def _reg_theta(self) -> torch.Tensor:
        """The part of the flat parameter buffer the ParameterNet kernel / bias regularisers act on (all of it: every
        variable of NIF / NIFMultiScale belongs to the ParameterNet)."""
        return self.theta

    def _kernel_regulariser(self) -> Tuple[float, float]:
        """(l1, l2) applied to every ParameterNet kernel and bias (nif/model.py:107-117); l2 wins."""
        if isinstance(self.p_l2_reg, (float, int)):
            return 0.0, float(self.p_l2_reg)
        if isinstance(self.p_l1_reg, (float, int)):
            return float(self.p_l1_reg), 0.0
        return 0.0, 0.0

    # ---- the reference's public surface ------------------------------------------------------------------
    def call(self, inputs, training=None, mask=None) -> torch.Tensor:
        """Forward pass on a (B, pi_dim + si_dim) batch (nif/model.py:130-154 / 510-539)."""
        return self.model()(inputs, training=bool(training))

    def build(self) -> Model:
        """nif/model.py:345-377: with `jac_reg` the reference wraps the model in JacRegLatentLayer, which only adds a loss
        term; here that term is added by the training step (Model._train_step), so this is `.model()`."""
        return self.model()

    def model(self) -> Model:
        return Model(self, "full")

    def model_p_to_w(self) -> Model:
        return Model(self, "p_to_w")

    def model_p_to_lr(self) -> Model:
        return Model(self, "p_to_lr")

    def model_lr_to_w(self) -> Model:
        return Model(self, "lr_to_w")

    def model_x_to_u_given_w(self) -> Model:
        return Model(self, "x_to_u_given_w")

    def save_config(self, filename="config.json"):
        """nif/model.py:466-480: the two cfg dicts and the policy name, as JSON."""
        config = {
            "cfg_shape_net": self.cfg_shape_net,
            "cfg_parameter_net": self.cfg_parameter_net,
            "mixed_policy": self.mixed_policy_name,
        }
        with open(filename, "w") as write_file:
            json.dump(config, write_file, indent=4)


class NIFMultiScale(NIF):
    """SIREN ShapeNet (omega_0-scaled sine, optional 2-layer res-blocks) with a swish-MLP or SIREN
    ParameterNet (reference: class NIFMultiScale, nif/model.py:483-986)."""

    def _validate_cfg(self, cfg_shape_net, cfg_parameter_net):
        # same checks and messages as nif/model.py:555-587
        if not isinstance(cfg_parameter_net, dict):
            raise TypeError("cfg_parameter_net must be a dictionary")
        if not isinstance(cfg_shape_net, dict):
            raise TypeError("cfg_shape_net must be a dictionary")
        assert "use_resblock" in cfg_shape_net.keys(), "`use_resblock` should be in cfg_shape_net"
        assert type(cfg_shape_net["use_resblock"]) == bool, "cfg_shape_net['use_resblock'] must be a bool"
        conn = cfg_shape_net.get("connectivity")
        if conn == "last_layer" and not getattr(self, "_last_layer_only", False):
            raise ValueError("connectivity='last_layer' is the class NIFMultiScaleLastLayerParameterized")
        if conn not in ("full", "last_layer"):
            raise ValueError("cfg_shape_net missing correct `connectivity`")
        self._variant = "siren_res" if cfg_shape_net["use_resblock"] else "siren"

    def _po_dim(self) -> int:
        # nif/model.py:572-582
        h = 2 * self.l_sx if self.cfg_shape_net["use_resblock"] else self.l_sx
        return h * self.n_sx**2 + (self.si_dim + self.so_dim + 1 + h) * self.n_sx + self.so_dim

    @property
    def _omega0(self) -> float:
        return float(self.cfg_shape_net["omega_0"])

    _last_names = ("HyperLinearForSIREN_w", "HyperLinearForSIREN_b")

    @property
    def _sine_trunk(self) -> bool:
        return self.cfg_parameter_net["activation"] == "sine"

    def _trunk_layout(self):
        p = self.cfg_parameter_net
        res = bool(p.get("use_resblock", False))
        L = []
        if self._sine_trunk:  # nif/model.py:591-648
            L += [("siren_first_pnet_w", (self.pi_dim, self.n_st)), ("siren_first_pnet_b", (self.n_st,))]
            for i in range(self.l_st):
                if res:
                    q = f"siren_hidden_resblock_pnet_{i}"
                    L += [(q + "_w", (self.n_st, self.n_st)), (q + "_b", (self.n_st,)),
                          (q + "_w2", (self.n_st, self.n_st)), (q + "_b2", (self.n_st,))]
                else:
                    q = f"siren_hidden_pnet_{i}"
                    L += [(q + "_w", (self.n_st, self.n_st)), (q + "_b", (self.n_st,))]
            L += [("siren_bottleneck_pnet_w", (self.n_st, self.pi_hidden)), ("siren_bottleneck_pnet_b", (self.pi_hidden,))]
        else:  # nif/model.py:665-720
            L += [("mlp_first_pnet/kernel", (self.pi_dim, self.n_st)), ("mlp_first_pnet/bias", (self.n_st,))]
            for i in range(self.l_st):
                if res:
                    q = f"mlp_hidden_resblock_pnet_{i}"
                    L += [(q + "_dense_1/kernel", (self.n_st, self.n_st)), (q + "_dense_1/bias", (self.n_st,)),
                          (q + "_dense_2/kernel", (self.n_st, self.n_st)), (q + "_dense_2/bias", (self.n_st,))]
                else:
                    q = f"mlp_hidden_pnet_{i}"
                    L += [(q + "/kernel", (self.n_st, self.n_st)), (q + "/bias", (self.n_st,))]
            L += [("bottleneck_pnet/kernel", (self.n_st, self.pi_hidden)), ("bottleneck_pnet/bias", (self.pi_hidden,))]
        return L

    def _draw(self, name: str, shape):
        s, p = self.cfg_shape_net, self.cfg_parameter_net
        if name == "HyperLinearForSIREN_w":
            # gen_hypernetwork_weights_bias_for_siren_shapenet (nif/layers/siren.py:36-40)
            return _uniform(shape, math.sqrt(6.0 / self.pi_hidden) * s["weight_init_factor"], self._gen)
        if name == "HyperLinearForSIREN_b":
            # per-column bounds (nif/layers/siren.py:42-62)
            n, si, so = self.n_sx, self.si_dim, self.so_dim
            H = 2 * self.l_sx if s["use_resblock"] else self.l_sx
            bound = np.ones(self.po_dim)
            n1, nh, nl = si * n, H * n * n, so * n
            bound[:n1] /= si
            bound[n1:n1 + nh] *= math.sqrt(6.0 / n) / s["omega_0"]
            bound[n1 + nh:n1 + nh + nl] *= math.sqrt(6.0 / (2 * n))
            bound[n1 + nh + nl:] /= n
            return _uniform(shape, bound, self._gen)
        if name.startswith("siren_"):
            # SIREN layer rules (nif/layers/siren.py:178-204); SIREN_ResNet copies w/b into w2/b2 (:370-379)
            w0 = float(p["omega_0"])
            if name.endswith("_w2") or name.endswith("_b2"):
                return self._drawn[name[:-1]].clone()
            first = name.startswith("siren_first")
            fan_in = self.pi_dim if first else self.n_st
            if name.endswith("_w"):
                t = _uniform(shape, 1.0 / fan_in if first else math.sqrt(6.0 / fan_in) / w0, self._gen)
            else:
                t = _uniform(shape, 1.0 / math.sqrt(fan_in), self._gen)
            self._drawn[name] = t
            return t
        return _trunc_normal(shape, 0.1, self._gen)

    def _init_parameters(self):
        self._drawn: Dict[str, torch.Tensor] = {}
        super()._init_parameters()
        self._drawn = {}

    def _latent(self, input_p: torch.Tensor) -> torch.Tensor:
        V = self._views
        p = self.cfg_parameter_net
        res = bool(p.get("use_resblock", False))
        if self._sine_trunk:  # SIREN / SIREN_ResNet (nif/layers/siren.py:256-281, 381-410)
            w0 = float(p["omega_0"])
            h = torch.sin(w0 * (input_p @ V["siren_first_pnet_w"]) + V["siren_first_pnet_b"])
            for i in range(self.l_st):
                if res:
                    q = f"siren_hidden_resblock_pnet_{i}"
                    g = torch.sin(w0 * (h @ V[q + "_w"]) + V[q + "_b"])
                    h = 0.5 * (h + torch.sin(w0 * (g @ V[q + "_w2"]) + V[q + "_b2"]))
                else:
                    q = f"siren_hidden_pnet_{i}"
                    h = torch.sin(w0 * (h @ V[q + "_w"]) + V[q + "_b"])
            return h @ V["siren_bottleneck_pnet_w"] + V["siren_bottleneck_pnet_b"]
        if not res and _wide_trunk_ok(self, input_p, int(p["units"]), p["activation"]):
            names = ["mlp_first_pnet"] + [f"mlp_hidden_pnet_{i}" for i in range(self.l_st)] + ["bottleneck_pnet"]
            wb = [V[q + s] for q in names for s in ("/kernel", "/bias")]
            return _WideTrunk.apply(input_p, ACT[p["activation"]], *wb)
        f = self._act(p["activation"])
        h = f(self._lin(input_p, V["mlp_first_pnet/kernel"], V["mlp_first_pnet/bias"]))
        for i in range(self.l_st):
            if res:  # MLP_ResNet (nif/layers/mlp.py:62-79)
                q = f"mlp_hidden_resblock_pnet_{i}"
                h1 = f(self._lin(h, V[q + "_dense_1/kernel"], V[q + "_dense_1/bias"]))
                h = f(h + self._lin(h1, V[q + "_dense_2/kernel"], V[q + "_dense_2/bias"]))
            else:  # MLP_SimpleShortCut (nif/layers/mlp.py:148-160)
                q = f"mlp_hidden_pnet_{i}"
                h = h + f(self._lin(h, V[q + "/kernel"], V[q + "/bias"]))
        return self._lin(h, V["bottleneck_pnet/kernel"], V["bottleneck_pnet/bias"])


class NIFMultiScaleLastLayerParameterized(NIFMultiScale):
    """NIFMultiScale whose ParameterNet parameterises only the LAST layer of the ShapeNet (reference: class
    NIFMultiScaleLastLayerParameterized, nif/model.py:989-1269): the ShapeNet is a shared-weight SIREN
    x -> phi(x) in R^{so x pi_hidden} (:1151-1242) and  u = phi(x) . pnet_output + last_layer_bias  (:1244-1269), a
    DeepONet-style product.  pnet_output = latent @ HyperLinearForSIREN_w + _b with po_dim = pi_hidden (:583-585).

    There are no per-sample weights here, so this class is not the fused hyper-network hot path: the ParameterNet trunk
    runs on the fused trunk kernels where they apply, the shared-weight ShapeNet and the final product are library GEMMs
    (torch) differentiated by autograd into the same flat gradient buffer."""

    _last_layer_only = True

    def __init__(self, cfg_shape_net, cfg_parameter_net, mixed_policy="float32", **kw):
        assert cfg_shape_net["connectivity"] == "last_layer", \
            "you should assign cfg_shape_net['connectivity'] == 'last_layer'"
        self.s_l1_reg = cfg_shape_net.get("l1_reg", None)
        self.s_l2_reg = cfg_shape_net.get("l2_reg", None)
        super().__init__(cfg_shape_net, cfg_parameter_net, mixed_policy, **kw)

    def _po_dim(self) -> int:
        return self.pi_hidden  # nif/model.py:583-585

    def _fused_trunk_supported(self) -> bool:
        return False  # the whole model is differentiated by autograd; keep one mechanism

    def _extra_layout(self):
        s = self.cfg_shape_net
        L = [("siren_first_snet_w", (self.si_dim, self.n_sx)), ("siren_first_snet_b", (self.n_sx,))]
        for i in range(self.l_sx):
            if s["use_resblock"]:
                q = f"siren_hidden_resblock_snet_{i}"
                L += [(q + "_w", (self.n_sx, self.n_sx)), (q + "_b", (self.n_sx,)),
                      (q + "_w2", (self.n_sx, self.n_sx)), (q + "_b2", (self.n_sx,))]
            else:
                q = f"siren_hidden_snet_{i}"
                L += [(q + "_w", (self.n_sx, self.n_sx)), (q + "_b", (self.n_sx,))]
        no = self.po_dim * self.so_dim
        L += [("siren_bottleneck_snet_w", (self.n_sx, no)), ("siren_bottleneck_snet_b", (no,)),
              ("last_layer_bias_snet", (self.so_dim,))]
        return L

    def _draw(self, name: str, shape):
        s = self.cfg_shape_net
        if name == "HyperLinearForSIREN_b":
            # connectivity 'last_layer' (nif/layers/siren.py:482-483): every output column counts as a last-layer weight
            return _uniform(shape, math.sqrt(6.0 / (2 * self.n_sx)), self._gen)
        if name == "last_layer_bias_snet":
            return _trunc_normal(shape, 0.1, self._gen)  # BiasAddLayer (nif/layers/mlp.py:245-250)
        if "_snet" in name:
            # SIREN layer rules with the ShapeNet's omega_0 (nif/layers/siren.py:178-204); SIREN_ResNet copies (:370-379)
            w0 = float(s["omega_0"])
            if name.endswith("_w2") or name.endswith("_b2"):
                return self._drawn[name[:-1]].clone()
            first = name.startswith("siren_first")
            fan_in = self.si_dim if first else self.n_sx
            if name.endswith("_w"):
                t = _uniform(shape, 1.0 / fan_in if first else math.sqrt(6.0 / fan_in) / w0, self._gen)
            else:
                t = _uniform(shape, 1.0 / math.sqrt(fan_in), self._gen)
            self._drawn[name] = t
            return t
        return super()._draw(name, shape)

    # ---- forward pieces (differentiable torch ops over the flat-buffer views) ----
    def _phi(self, x: torch.Tensor) -> torch.Tensor:
        """_call_shape_net_get_phi_x (nif/model.py:1222-1242): x -> [B, so, pi_hidden]."""
        V, s = self._views, self.cfg_shape_net
        w0 = float(s["omega_0"])
        h = torch.sin(w0 * (x @ V["siren_first_snet_w"]) + V["siren_first_snet_b"])
        for i in range(self.l_sx):
            if s["use_resblock"]:
                q = f"siren_hidden_resblock_snet_{i}"
                g = torch.sin(w0 * (h @ V[q + "_w"]) + V[q + "_b"])
                h = 0.5 * (h + torch.sin(w0 * (g @ V[q + "_w2"]) + V[q + "_b2"]))
            else:
                q = f"siren_hidden_snet_{i}"
                h = torch.sin(w0 * (h @ V[q + "_w"]) + V[q + "_b"])
        phi = h @ V["siren_bottleneck_snet_w"] + V["siren_bottleneck_snet_b"]
        return phi.reshape(-1, self.so_dim, self.pi_hidden)

    def _pnet(self, p_in: torch.Tensor) -> torch.Tensor:
        return self._latent(p_in) @ self.w_h + self.b_h

    def _u_given(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        """_call_shape_net_mres_only_para_last_layer (nif/model.py:1244-1269): Dot(axes=(2,1)) + BiasAddLayer."""
        return torch.einsum("bok,bk->bo", self._phi(x), w) + self._views["last_layer_bias_snet"]

    def _forward_train(self, inp: torch.Tensor) -> torch.Tensor:
        return self._u_given(inp[:, self.pi_dim: self.pi_dim + self.si_dim], self._pnet(inp[:, : self.pi_dim]))

    def _forward_kind(self, kind: str, x: torch.Tensor, w: Optional[torch.Tensor] = None) -> torch.Tensor:
        if kind == "full":
            return self._forward_train(x)
        if kind == "p_to_lr":
            return self._pnet(x)  # nif/model.py:1070-1083: in this class the "hidden LR" is the ParameterNet output
        if kind == "x_to_phi":
            return self._phi(x)
        if kind == "x_to_u_given_w":
            return self._u_given(x, w)
        raise ValueError(kind)

    def _count_params_kind(self, kind: str) -> int:
        snet = sum(int(np.prod(sh)) for nm, (_, sh) in self._layout.items() if "_snet" in nm)
        return {"full": self.count_params(), "p_to_lr": self.count_params() - snet,
                "x_to_phi": snet - self.so_dim, "x_to_u_given_w": snet}[kind]

    def _reg_theta(self) -> torch.Tensor:
        """ParameterNet regularisers act on the ParameterNet variables only (the ShapeNet has its own l1_reg / l2_reg keys,
        nif/model.py:1028-1040, which are not built here)."""
        first_snet = self._layout["siren_first_snet_w"][0]
        return self.theta[:first_snet]

    def _kernel_regulariser(self):
        l1, l2 = super()._kernel_regulariser()
        if l1 or l2:
            raise NifError("ParameterNet l1_reg / l2_reg with NIFMultiScaleLastLayerParameterized is not built "
                           "(the optimiser-side regulariser acts on the whole flat buffer)")
        if isinstance(self.s_l1_reg, (float, int)) or isinstance(self.s_l2_reg, (float, int)):
            raise NifError("cfg_shape_net l1_reg / l2_reg is not built")
        return 0.0, 0.0

    # ---- the reference's public surface ----
    def model_p_to_w(self) -> Model:
        raise ValueError("In this class: NIFMultiScaleLastLayerParameterization, `w` is the same as `lr`")

    def model_lr_to_w(self) -> Model:
        raise ValueError("In this class: NIFMultiScaleLastLayerParameterization, `w` is the same as `lr`")

    def model_x_to_phi(self) -> Model:
        return Model(self, "x_to_phi")
