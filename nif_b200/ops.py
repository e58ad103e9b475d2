"""Host-side operators over the C ABI: torch tensors in, torch tensors out.

torch is used for device memory, streams and autograd plumbing only; every
number on the hot path is produced by libnif_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import ACT, VARIANT, Desc, NifError, Sizes, TrunkDesc, check


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise NifError(f"{name} must be a CUDA tensor (nif_b200 has no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:  # e.g. a row slice a[s:e] of a cached array: the C ABI wants 16-byte aligned pointers
        t = t.clone()
    return t


class FusedShapeNet:
    """The hyper-network head + ShapeNet of one model, bound to a descriptor.

    Mirrors the static kernels of the reference:
      NIF._call_shape_net(input_s, pnet_output, si_dim, so_dim, n_sx, l_sx, activation, variable_dtype)
        (nif/model.py:233-236)
      NIFMultiScale._call_shape_net_mres(input_s, pnet_output, flag_resblock, omega_0, si_dim, so_dim,
        n_sx, l_sx, variable_dtype)   (nif/model.py:738-749)
    except that `pnet_output` is never formed: the operator takes the latent z and the last
    linear layer (w_h, b_h) instead.
    """

    # fp32 CUDA cores | tensor cores, bf16 operands (mixed_bfloat16) | tensor cores, 3-product fp16 split (fp32-grade)
    COMPUTE = {"fp32": 0, "bf16": 1, "fp16x3": 2}

    def __init__(self, variant: str, si: int, so: int, n: int, l: int, K: int,
                 activation: Optional[str] = "swish", omega0: float = 1.0, compute: str = "fp32", acc_rows: int = 0):
        if variant not in VARIANT:
            raise ValueError(f"variant must be one of {list(VARIANT)}")
        if variant == "nif" and activation not in ACT:
            raise ValueError(f"activation {activation!r} is not supported by the fused kernels {list(ACT)}")
        self.variant, self.si, self.so, self.n, self.l, self.K = variant, si, so, n, l, K
        self.activation, self.omega0 = activation, float(omega0)
        if compute not in self.COMPUTE:
            raise ValueError(f"compute must be one of {list(self.COMPUTE)}")
        self.compute = compute
        self.desc = Desc(VARIANT[variant], ACT[activation] if variant == "nif" else ACT["sine"], si, so, n, l, K,
                         float(omega0) if variant != "nif" else 1.0, self.COMPUTE[compute], int(acc_rows))
        self.acc_rows = int(acc_rows)
        s = Sizes()
        check(_lib.lib().nif_query_sizes(C.byref(self.desc), 0, C.byref(s)), "nif_query_sizes")
        self.po_dim, self.np, self.packed_floats = int(s.po_dim), int(s.np), int(s.packed_floats)
        self.save_floats_per_row, self.tile_rows = int(s.save_floats_per_row), int(s.tile_rows)
        # which kernels the library will run for this descriptor (shapes a tensor-core path lacks fall through to
        # the CUDA-core kernels): "fp32" | "fp16x3" | "bf16"
        self.kernel_path = {v: k for k, v in self.COMPUTE.items()}[int(s.kernel_path)]
        # forward tangents along ShapeNet inputs (and the Sobolev step) on the tensor-core kernels: sine, no residual layers
        # (nif_plan_tc_sobolev in csrc/nif_tc_fwd.cu)
        self.tc_sobolev = self.kernel_path == "fp16x3" and variant == "siren"
        self._ws = None

    def with_latent(self, K: int) -> "FusedShapeNet":
        # the FP16x3 tensor-core kernels need a latent contraction (K >= 1) and a single group; a K == 0 engine
        # (explicit per-group weight vectors) runs on the kernels that serve grouped launches
        comp = self.compute
        if K == 0 and comp == "fp16x3":
            comp = "fp32"
        return FusedShapeNet(self.variant, self.si, self.so, self.n, self.l, K, self.activation, self.omega0, comp,
                             self.acc_rows)

    # ------------------------------------------------------------------------------------------
    def grad_ws_floats(self, B: int) -> int:
        s = Sizes()
        check(_lib.lib().nif_query_sizes(C.byref(self.desc), B, C.byref(s)), "nif_query_sizes")
        return int(s.grad_ws_floats)

    def _workspace(self, B: int, device) -> torch.Tensor:
        need = self.grad_ws_floats(B)
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.float32, device=device)
        return self._ws

    def pack(self, w_h: Optional[torch.Tensor], b_h: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Re-lay (w_h [K,P], b_h [P] or [G,P]) for the kernels."""
        b_h = _f32c(b_h, "b_h")
        G = 1 if b_h.dim() == 1 else b_h.shape[0]
        if b_h.shape[-1] != self.po_dim:
            raise NifError(f"b_h has {b_h.shape[-1]} columns, po_dim is {self.po_dim}")
        if self.K > 0:
            w_h = _f32c(w_h, "w_h")
            if tuple(w_h.shape) != (self.K, self.po_dim):
                raise NifError(f"w_h must be [{self.K},{self.po_dim}], got {tuple(w_h.shape)}")
        if out is None:
            out = torch.empty(G * self.packed_floats, dtype=torch.float32, device=b_h.device)
        check(_lib.lib().nif_pack(C.byref(self.desc), G, _ptr(w_h) if self.K > 0 else None, _ptr(b_h), _ptr(out),
                                  _stream()), "nif_pack")
        return out

    def forward(self, z: Optional[torch.Tensor], x: torch.Tensor, packed: torch.Tensor, save: bool = False,
                groups: int = 1, x_shared: bool = False):
        """u = ShapeNet(x; z @ w_h + b_h).  Returns u, or (u, stash) when save=True."""
        x = _f32c(x, "x")
        if x_shared:
            B = x.shape[0]
        else:
            B = x.shape[0] // groups
        if self.K > 0:
            z = _f32c(z, "z")
            if z.shape[0] != groups * B or z.shape[1] != self.K:
                raise NifError(f"z must be [{groups * B},{self.K}], got {tuple(z.shape)}")
        if x.shape[-1] != self.si:
            raise NifError(f"x must have {self.si} columns")
        u = torch.empty(groups * B, self.so, dtype=torch.float32, device=x.device)
        # (rows rounded up to 64: the tensor-core path tiles the stash in row groups)
        stash = (torch.empty(self.save_floats_per_row * ((B + 63) // 64 * 64), dtype=torch.float32, device=x.device)
                 if save else None)
        check(_lib.lib().nif_forward(C.byref(self.desc), groups, B, _ptr(z) if self.K > 0 else None, _ptr(x),
                                     1 if x_shared else 0, _ptr(packed), _ptr(u), _ptr(stash), _stream()),
              "nif_forward")
        return (u, stash) if save else u

    def forward_tangent(self, z, x, packed, zdot: Optional[torch.Tensor], xdot: Optional[torch.Tensor],
                        save: bool = False):
        """(u, udot[n_dir,B,so]) for tangent directions zdot [n_dir,B,K] / xdot [n_dir,B,si].
        save=True also returns the stash of every direction for sobolev_backward."""
        x = _f32c(x, "x")
        B = x.shape[0]
        z = _f32c(z, "z") if self.K > 0 else None
        n_dir = (zdot if zdot is not None else xdot).shape[0]
        zdot = _f32c(zdot, "zdot") if zdot is not None else None
        xdot = _f32c(xdot, "xdot") if xdot is not None else None
        u = torch.empty(B, self.so, dtype=torch.float32, device=x.device)
        udot = torch.empty(n_dir, B, self.so, dtype=torch.float32, device=x.device)
        if not save and zdot is None and xdot is not None and self.tc_sobolev and B > 0:
            # ShapeNet-input directions of a SIREN ShapeNet the tensor-core kernels serve: the library runs the stashing
            # pair (forward kernel + its tangent mode) on them; here the stash is scratch, so the rows go in blocks
            per_row = C.c_int64(0)
            check(_lib.lib().nif_sobolev_query_dirs(C.byref(self.desc), B, n_dir, C.byref(per_row), None), "nif_sobolev_query_dirs")
            blk = min(B, 65536)
            scratch = torch.empty(int(per_row.value) * ((blk + 63) // 64 * 64), dtype=torch.float32, device=x.device)
            for s in range(0, B, blk):
                e = min(B, s + blk)
                ub = u[s:e]
                ud = torch.empty(n_dir, e - s, self.so, dtype=torch.float32, device=x.device)
                xd = xdot[:, s:e].contiguous()
                check(_lib.lib().nif_forward_tangent_save(C.byref(self.desc), e - s, _ptr(z[s:e]) if z is not None else None,
                                                          _ptr(x[s:e]), _ptr(packed), n_dir, None, _ptr(xd), _ptr(ub),
                                                          _ptr(ud), _ptr(scratch), _stream()), "nif_forward_tangent_save")
                udot[:, s:e] = ud
            return u, udot
        if not save:
            check(_lib.lib().nif_forward_tangent(C.byref(self.desc), B, _ptr(z), _ptr(x), _ptr(packed), n_dir,
                                                 _ptr(zdot), _ptr(xdot), _ptr(u), _ptr(udot), _stream()),
                  "nif_forward_tangent")
            return u, udot
        if xdot is None:
            xdot = torch.zeros(n_dir, B, self.si, dtype=torch.float32, device=x.device)
        per_row = C.c_int64(0)
        check(_lib.lib().nif_sobolev_query_dirs(C.byref(self.desc), B, n_dir, C.byref(per_row), None), "nif_sobolev_query_dirs")
        # (rows rounded up to 64: the tensor-core path tiles the stash in row groups)
        stash = torch.empty(int(per_row.value) * ((B + 63) // 64 * 64), dtype=torch.float32, device=x.device)
        check(_lib.lib().nif_forward_tangent_save(C.byref(self.desc), B, _ptr(z), _ptr(x), _ptr(packed), n_dir,
                                                  _ptr(zdot), _ptr(xdot), _ptr(u), _ptr(udot), _ptr(stash), _stream()),
              "nif_forward_tangent_save")
        return u, udot, stash

    def forward_tangent2(self, z, x, packed, zdot: Optional[torch.Tensor], xdot: Optional[torch.Tensor],
                         zddot: Optional[torch.Tensor]):
        """Second-order forward mode for one pair of directions (a, b): zdot [2,B,K] / xdot [2,B,si] (either may be
        None), zddot [B,K] = second derivative of the latent code along (a, b) (None = 0).  Returns
        (u [B,so], udot [2,B,so] = (du/da, du/db), uddot [B,so] = d2u/da db)."""
        x = _f32c(x, "x")
        B = x.shape[0]
        z = _f32c(z, "z") if self.K > 0 else None
        zdot = _f32c(zdot, "zdot") if zdot is not None else None
        xdot = _f32c(xdot, "xdot") if xdot is not None else None
        zddot = _f32c(zddot, "zddot") if zddot is not None else None
        for name, t, shape in (("zdot", zdot, (2, B, self.K)), ("xdot", xdot, (2, B, self.si)), ("zddot", zddot, (B, self.K))):
            if t is not None and tuple(t.shape) != shape:
                raise NifError(f"{name} must be {list(shape)}, got {list(t.shape)}")
        u = torch.empty(B, self.so, dtype=torch.float32, device=x.device)
        udot = torch.empty(2, B, self.so, dtype=torch.float32, device=x.device)
        uddot = torch.empty(B, self.so, dtype=torch.float32, device=x.device)
        check(_lib.lib().nif_forward_tangent2(C.byref(self.desc), B, _ptr(z), _ptr(x), _ptr(packed), _ptr(zdot), _ptr(xdot),
                                              _ptr(zddot), _ptr(u), _ptr(udot), _ptr(uddot), _stream()),
              "nif_forward_tangent2")
        return u, udot, uddot

    def sobolev_backward(self, z, x, xdot, packed, stash, du, dudot, dw_h, db_h, beta: float = 0.0, zdot=None,
                         zdot_dirs=None):
        """Reverse-over-forward pass over the directions of the forward_tangent(save=True) call: seeds du = dL/du [B,so]
        and dudot = dL/d(udot) [n_dir,B,so]; xdot [n_dir,B,si] and zdot [n_dir,B,K] (None: no direction moves the latent
        code) are the directions of that call ([B,*] is read as one direction).  Fills dw_h / db_h; returns dz, or
        (dz, dzdot [n_dir,B,K]) when zdot is given.  zdot_dirs: indices of the directions whose zdot is not identically zero
        (default: all of them)"""
        x, xdot = _f32c(x, "x"), _f32c(xdot, "xdot")
        B = x.shape[0]
        du, dudot = _f32c(du, "du"), _f32c(dudot, "dudot")
        if xdot.dim() == 2:
            xdot = xdot.unsqueeze(0)
        if dudot.dim() == 2:
            dudot = dudot.unsqueeze(0)
        n_dir = xdot.shape[0]
        if tuple(xdot.shape) != (n_dir, B, self.si) or tuple(dudot.shape) != (n_dir, B, self.so):
            raise NifError(f"xdot / dudot must be [{n_dir},{B},{self.si}] / [{n_dir},{B},{self.so}], got "
                           f"{tuple(xdot.shape)} / {tuple(dudot.shape)}")
        dzdot = None
        if zdot is not None and self.K > 0:
            zdot = _f32c(zdot, "zdot")
            if tuple(zdot.shape) != (n_dir, B, self.K):
                raise NifError(f"zdot must be [{n_dir},{B},{self.K}], got {tuple(zdot.shape)}")
            dzdot = torch.empty(n_dir, B, self.K, dtype=torch.float32, device=x.device)
        else:
            zdot = None
        per_row = C.c_int64(0)
        check(_lib.lib().nif_sobolev_query_dirs(C.byref(self.desc), B, n_dir, C.byref(per_row), None), "nif_sobolev_query_dirs")
        if stash.numel() < per_row.value * B:
            raise NifError(f"the stash does not hold {n_dir} directions")
        dz = torch.empty(B, self.K, dtype=torch.float32, device=x.device) if self.K > 0 else None
        wsn = C.c_int64(0)
        check(_lib.lib().nif_sobolev_query_dirs(C.byref(self.desc), B, n_dir, None, C.byref(wsn)), "nif_sobolev_query_dirs")
        if self._ws is None or self._ws.numel() < wsn.value or self._ws.device != x.device:
            self._ws = torch.empty(int(wsn.value), dtype=torch.float32, device=x.device)
        mask = 0
        if zdot is not None:
            for d in (range(n_dir) if zdot_dirs is None else zdot_dirs):
                mask |= 1 << int(d)
        check(_lib.lib().nif_sobolev_backward_dirs(C.byref(self.desc), B, _ptr(z), _ptr(x), n_dir, _ptr(zdot), mask, _ptr(xdot),
                                                   _ptr(packed), _ptr(stash), _ptr(du), _ptr(dudot), _ptr(dw_h), _ptr(db_h),
                                                   float(beta), _ptr(dz), _ptr(dzdot), _ptr(self._ws), _stream()),
              "nif_sobolev_backward_dirs")
        return (dz, dzdot) if zdot is not None else dz

    def given_w(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        """model_x_to_u_given_w: every row of `w` is a full weight vector (nif/model.py:435-464, 956-986)."""
        x, w = _f32c(x, "x"), _f32c(w, "w")
        if w.shape != (x.shape[0], self.po_dim):
            raise NifError(f"w must be [{x.shape[0]},{self.po_dim}], got {tuple(w.shape)}")
        u = torch.empty(x.shape[0], self.so, dtype=torch.float32, device=x.device)
        check(_lib.lib().nif_forward_given_w(C.byref(self.desc), x.shape[0], _ptr(x), _ptr(w), _ptr(u), _stream()),
              "nif_forward_given_w")
        return u

    def backward(self, z, x, packed, stash, du, dw_h, db_h, beta: float = 0.0):
        """Reverse pass for a caller-supplied seed du [B,so]; fills dw_h/db_h, returns dz."""
        B = x.shape[0]
        du = _f32c(du, "du")
        dz = torch.empty(B, self.K, dtype=torch.float32, device=x.device) if self.K > 0 else None
        ws = self._workspace(B, x.device)
        check(_lib.lib().nif_backward(C.byref(self.desc), B, _ptr(z), _ptr(x), _ptr(packed), _ptr(stash), _ptr(du),
                                      _ptr(dw_h), _ptr(db_h), float(beta), _ptr(dz), _ptr(ws), _stream()),
              "nif_backward")
        return dz

    def mse_backward(self, z, x, packed, u, stash, target, sample_weight, inv_global_batch: float, loss, dw_h, db_h,
                     beta: float = 0.0, dz_event: Optional["torch.cuda.Event"] = None):
        """Keras 'mse' + reverse pass.  `loss` is a 1-element tensor that is accumulated into.  dz_event (already created:
        recorded at least once) is recorded on the current stream as soon as dz is final."""
        B = x.shape[0]
        target = _f32c(target, "target")
        sw = _f32c(sample_weight, "sample_weight") if sample_weight is not None else None
        dz = torch.empty(B, self.K, dtype=torch.float32, device=x.device) if self.K > 0 else None
        ws = self._workspace(B, x.device)
        ev = C.c_void_p(dz_event.cuda_event) if dz_event is not None else None
        check(_lib.lib().nif_mse_backward_ev(C.byref(self.desc), B, _ptr(z), _ptr(x), _ptr(packed), _ptr(u), _ptr(stash),
                                             _ptr(target), _ptr(sw), float(inv_global_batch), _ptr(loss), _ptr(dw_h),
                                             _ptr(db_h), float(beta), _ptr(dz), _ptr(ws), ev, _stream()),
              "nif_mse_backward")
        return dz


class FusedTrunk:
    """ParameterNet trunk up to the bottleneck (nif/model.py:326-343) as fused kernels: Dense(act) ->
    nlayers x MLP_SimpleShortCut -> Dense(latent).  `theta` is the flat trunk weight vector in the column
    order documented in include/nif_b200.h."""

    def __init__(self, pi: int, latent: int, units: int, nlayers: int, activation: str):
        if activation not in ACT or activation == "sine":
            raise NifError(f"trunk activation {activation!r} is outside the fused trunk kernels")
        self.pi, self.latent, self.units, self.nlayers = pi, latent, units, nlayers
        self.desc = TrunkDesc(pi, latent, units, nlayers, ACT[activation])
        nt, sv, pk = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(_lib.lib().nif_trunk_query(C.byref(self.desc), 0, C.byref(nt), C.byref(sv), C.byref(pk), None),
              "nif_trunk_query")
        self.n_theta, self.save_floats_per_row, self.packed_floats = int(nt.value), int(sv.value), int(pk.value)
        # "bf16x3": tcgen05 tensor cores, fp32-grade three-way bf16 split | "fp32": CUDA-core tile kernels
        self.kernel_path = "bf16x3" if _lib.lib().nif_trunk_kernel_path(C.byref(self.desc)) == 3 else "fp32"
        self._ws = None
        self._packed = None

    def _packed_buf(self, device) -> torch.Tensor:
        if self._packed is None or self._packed.device != device:
            self._packed = torch.empty(self.packed_floats, dtype=torch.float32, device=device)
        return self._packed

    def forward(self, p_in: torch.Tensor, theta: torch.Tensor, save: bool = False):
        p_in = _f32c(p_in, "p_in")
        B = p_in.shape[0]
        z = torch.empty(B, self.latent, dtype=torch.float32, device=p_in.device)
        stash = (torch.empty(self.save_floats_per_row * ((B + 63) // 64 * 64), dtype=torch.float32, device=p_in.device)
                 if save else None)
        check(_lib.lib().nif_trunk_forward(C.byref(self.desc), B, _ptr(p_in), _ptr(theta), _ptr(z), _ptr(stash),
                                           _ptr(self._packed_buf(p_in.device)), _stream()), "nif_trunk_forward")
        return (z, stash) if save else z

    def backward(self, p_in, theta, stash, dz, g_theta, beta: float = 0.0):
        """Reverse pass of the forward(save=True) call that preceded it (it reuses that call's re-laid weights)."""
        p_in = _f32c(p_in, "p_in")
        B = p_in.shape[0]
        wsn = C.c_int64(0)
        check(_lib.lib().nif_trunk_query(C.byref(self.desc), B, None, None, None, C.byref(wsn)), "nif_trunk_query")
        if self._ws is None or self._ws.numel() < wsn.value or self._ws.device != p_in.device:
            self._ws = torch.empty(int(wsn.value), dtype=torch.float32, device=p_in.device)
        check(_lib.lib().nif_trunk_backward(C.byref(self.desc), B, _ptr(p_in), _ptr(theta), _ptr(stash),
                                            _ptr(_f32c(dz, "dz")), _ptr(g_theta), float(beta),
                                            _ptr(self._packed_buf(p_in.device)), _ptr(self._ws), _stream()),
              "nif_trunk_backward")


class _FusedFn(torch.autograd.Function):
    """autograd bridge: (z, x, w_h, b_h) -> u with the fused forward / reverse kernels."""

    @staticmethod
    def forward(ctx, z, x, w_h, b_h, engine: FusedShapeNet):
        packed = engine.pack(w_h, b_h)
        # the reverse pass hands these tensors' raw pointers to the library: normalise them once, here
        x = _f32c(x, "x")
        z = _f32c(z, "z") if engine.K > 0 else z
        need = any(ctx.needs_input_grad[:4])
        if need:
            u, stash = engine.forward(z, x, packed, save=True)
            ctx.save_for_backward(z, x, packed, stash)
            ctx.engine = engine
            ctx.shapes = (w_h.shape, b_h.shape)
        else:
            u = engine.forward(z, x, packed)
        return u

    @staticmethod
    def backward(ctx, du):
        z, x, packed, stash = ctx.saved_tensors
        e = ctx.engine
        dw = torch.empty(ctx.shapes[0], dtype=torch.float32, device=x.device)
        db = torch.empty(ctx.shapes[1], dtype=torch.float32, device=x.device)
        dz = e.backward(z, x, packed, stash, du.contiguous(), dw, db, 0.0)
        return dz, None, dw, db, None


def fused_shapenet(z, x, w_h, b_h, engine: FusedShapeNet):
    return _FusedFn.apply(z, x, w_h, b_h, engine)


def adam_step(p, g, m, v, lr, t, b1=0.9, b2=0.999, eps=1e-7, l1=0.0, l2=0.0, g_scale=1.0):
    """tf.keras Adam on flat fp32 buffers, in place."""
    check(_lib.lib().nif_adam_step(p.numel(), _ptr(p), _ptr(g), _ptr(m), _ptr(v), float(lr), float(b1), float(b2),
                                   float(eps), int(t), float(l1), float(l2), float(g_scale), _stream()),
          "nif_adam_step")


def adam_step_dev(p, g, m, v, alpha_dev, b1=0.9, b2=0.999, eps=1e-7, l1=0.0, l2=0.0, g_scale=1.0):
    """The same update with the bias-corrected step size read from the 1-element device tensor `alpha_dev`
    (graph-replayed steps: no launch argument changes from one step to the next)."""
    check(_lib.lib().nif_adam_step_dev(p.numel(), _ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(alpha_dev), float(b1),
                                       float(b2), float(eps), float(l1), float(l2), float(g_scale), _stream()),
          "nif_adam_step_dev")


def measure_fp32_peak() -> float:
    v = C.c_double(0.0)
    check(_lib.lib().nif_measure_fp32_peak(C.byref(v)), "nif_measure_fp32_peak")
    return float(v.value)


class kernel_profile:
    """Context manager around nif_profile_begin / nif_profile_end: per-kernel CUDA-event times of the library's launches
    (eager launches only).  After exit, `.table` is [(kernel name, launches, total ms)] in order of first launch."""

    def __enter__(self):
        check(_lib.lib().nif_profile_begin(), "nif_profile_begin")
        self.table = []
        return self

    def __exit__(self, *exc):
        buf = C.create_string_buffer(1 << 16)
        check(_lib.lib().nif_profile_end(buf, len(buf)), "nif_profile_end")
        for ln in buf.value.decode().splitlines():
            name, cnt, ms = ln.rsplit(" ", 2)
            self.table.append((name, int(cnt), float(ms)))
        return False
