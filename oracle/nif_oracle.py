"""CPU oracle for the NIF hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.  The product (`nif_b200/`)
never does; it fails loudly when its CUDA library is missing.

What it is: a plain torch-CPU restatement (fp64 = truth, fp32 = what a TF2 CPU
run computes up to summation order) of the reference algorithm in
`/root/reference/nif/model.py` and `nif/layers/*.py`.  Every function cites the
reference lines it restates.  Gradients come from torch autograd over this
restatement (the reference likewise relies on TF's tape), Jacobians from
autograd per output index exactly like `compute_output_and_grad`.

Parity pinning: TensorFlow 2.11 cannot run in this image, so the reference
cannot be executed as shipped.  Instead `tests/golden/make_golden.py` executes
the reference's *unmodified source files* from /root/reference on top of a
torch-backed TF shim (tests/golden/tf_shim) and commits the resulting
input/weight/output/gradient vectors under tests/golden/*.npz;
`tests/test_oracle_golden.py` checks this oracle against all of them, plus the
structural known answers stored in the reference notebooks (po_dim = 1951,
1 951 / 3 902 parameters, Jacobian shapes).  The TF *kernels* themselves
(MatMul, Einsum, Sin, sigmoid, Adam) are third-party and un-vendored
(tensorflow==2.11.1, requirements.txt:3): their arithmetic is restated from
their published definitions, so Adam / 'mse' semantics stay "unpinned".
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------
@dataclass
class Spec:
    """Everything the two cfg dicts determine (nif/model.py:84-91, 559-587)."""

    variant: str  # "nif" (act + residual), "siren", "siren_res"
    pi: int
    si: int
    so: int
    n: int  # ShapeNet units
    l: int  # ShapeNet hidden layers (res-blocks for siren_res)
    K: int  # latent_dim
    n_st: int
    l_st: int
    p_act: str  # ParameterNet activation ("sine" selects the SIREN trunk)
    p_resblock: bool = False
    p_omega0: float = 1.0
    s_act: str = "swish"  # ShapeNet activation for variant "nif"
    omega0: float = 1.0
    weight_init_factor: float = 1.0

    @property
    def po_dim(self) -> int:
        return po_dim(self.si, self.so, self.n, self.l, self.variant == "siren_res")


def spec_from_cfg(cls_name: str, cfg_shape_net: dict, cfg_parameter_net: dict) -> Spec:
    """cls_name is 'NIF' or 'NIFMultiScale' (nif/model.py:48, 483)."""
    s, p = cfg_shape_net, cfg_parameter_net
    common = dict(
        pi=p["input_dim"],
        si=s["input_dim"],
        so=s["output_dim"],
        n=s["units"],
        l=s["nlayers"],
        K=p["latent_dim"],
        n_st=p["units"],
        l_st=p["nlayers"],
        p_act=p["activation"],
    )
    if cls_name == "NIF":
        return Spec(variant="nif", s_act=s["activation"], **common)
    if cls_name == "NIFMultiScale":
        if s.get("connectivity") != "full":
            raise ValueError("oracle covers connectivity='full' only (model.py:569-587)")
        return Spec(
            variant="siren_res" if s["use_resblock"] else "siren",
            omega0=float(s["omega_0"]),
            weight_init_factor=float(s["weight_init_factor"]),
            p_resblock=bool(p.get("use_resblock", False)),
            p_omega0=float(p.get("omega_0", 1.0)),
            **common,
        )
    raise ValueError(cls_name)


def po_dim(si: int, so: int, n: int, l: int, resblock: bool) -> int:
    """nif/model.py:169-173 (NIF), 572-582 (NIFMultiScale)."""
    h = 2 * l if resblock else l
    return h * n * n + (si + so + 1 + h) * n + so


@dataclass
class Layout:
    """Column ranges of pnet_output (model.py:253-300, 769-846, 883-933).

    w[m] = (offset, n_in, n_out) for matrix m (row-major [n_in, n_out]);
    b[m] = (offset, n_out).  m = 0 first, 1..H hidden, H+1 last."""

    w: List[Tuple[int, int, int]] = field(default_factory=list)
    b: List[Tuple[int, int]] = field(default_factory=list)
    P: int = 0


def layout(si: int, so: int, n: int, l: int, resblock: bool) -> Layout:
    H = 2 * l if resblock else l
    L = Layout()
    off = 0
    L.w.append((off, si, n))
    off += si * n
    for _ in range(H):
        L.w.append((off, n, n))
        off += n * n
    L.w.append((off, n, so))
    off += n * so
    L.b.append((off, n))
    off += n
    for _ in range(H):
        L.b.append((off, n))
        off += n
    L.b.append((off, so))
    off += so
    L.P = off
    assert off == po_dim(si, so, n, l, resblock)
    return L


# ----------------------------------------------------------------------------
# activations (tf.keras.activations.get, model.py:303)
# ----------------------------------------------------------------------------
def activation(name: Optional[str]):
    if name in (None, "linear"):
        return lambda v: v
    if name == "swish":
        return lambda v: v * torch.sigmoid(v)
    if name == "tanh":
        return torch.tanh
    if name == "relu":
        return torch.relu
    if name == "sigmoid":
        return torch.sigmoid
    if name == "sine":
        return torch.sin
    raise ValueError(f"activation {name!r} not restated")


# ----------------------------------------------------------------------------
# hot path
# ----------------------------------------------------------------------------
def hyper_linear(z: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """HyperLinearForSIREN.call (nif/layers/siren.py:514-522) and the Keras
    Dense(po_dim) of class NIF (nif/model.py:220-230): y = z @ w + b."""
    return z @ w + b


def _split(p: Tensor, L: Layout):
    ws = [p[:, o : o + a * c].reshape(-1, a, c) for (o, a, c) in L.w]
    bs = [p[:, o : o + c] for (o, c) in L.b]
    return ws, bs


def _bvm(u: Tensor, W: Tensor) -> Tensor:
    """EinsumLayer('ai,aij->aj') (nif/layers/mlp.py:209-219)."""
    return torch.einsum("ai,aij->aj", u, W)


def shape_net_nif(x: Tensor, p: Tensor, si: int, so: int, n: int, l: int, act: str) -> Tensor:
    """NIF._call_shape_net (nif/model.py:233-324): u0 = act(x W1 + b1);
    u <- act(u Wk + bk) + u; y = u WL + bL."""
    L = layout(si, so, n, l, False)
    ws, bs = _split(p, L)
    f = activation(act)
    u = f(_bvm(x, ws[0]) + bs[0])
    for k in range(1, l + 1):
        u = f(_bvm(u, ws[k]) + bs[k]) + u
    return _bvm(u, ws[l + 1]) + bs[l + 1]


def shape_net_mres(
    x: Tensor, p: Tensor, resblock: bool, omega0: float, si: int, so: int, n: int, l: int
) -> Tensor:
    """NIFMultiScale._call_shape_net_mres (nif/model.py:738-954).  omega0
    multiplies the einsum only; the bias is added afterwards (:936-949).
    Res-block branch (:849-877): h = sin(w0 u Wa + ba); u <- 0.5 (u + sin(w0 h Wb + bb))."""
    L = layout(si, so, n, l, resblock)
    ws, bs = _split(p, L)
    u = torch.sin(omega0 * _bvm(x, ws[0]) + bs[0])
    if resblock:
        for k in range(l):
            h = torch.sin(omega0 * _bvm(u, ws[1 + 2 * k]) + bs[1 + 2 * k])
            u = 0.5 * (u + torch.sin(omega0 * _bvm(h, ws[2 + 2 * k]) + bs[2 + 2 * k]))
    else:
        for k in range(1, l + 1):
            u = torch.sin(omega0 * _bvm(u, ws[k]) + bs[k])
    return _bvm(u, ws[-1]) + bs[-1]


def _bf16(t: Tensor) -> Tensor:
    """Round to bfloat16 (round-to-nearest-even) and come back in the working precision."""
    return t.to(torch.float32).to(torch.bfloat16).to(t.dtype)


def shape_net_factored(spec: Spec, x: Tensor, z: Tensor, w_h: Tensor, b_h: Tensor, quant: Optional[str] = None) -> Tensor:
    """The same ShapeNet (nif/model.py:233-324, 738-954) in the re-associated form the kernels evaluate, with the
    latent code kept separate from the last linear layer:
        pre_m[b,j] = omega * sum_{kappa,i} zt[b,kappa] h[b,i] M_m[kappa,i,j] + sum_kappa zt[b,kappa] C_m[kappa,j],
    zt = [z, 1], M_m / C_m = the column slices of [w_h; b_h].  quant=None is exact arithmetic in the working precision
    (equal to shape_net(hyper_linear(z, w_h, b_h)) up to summation order).

    quant='bf16_main' is 'bf16' with layer 0 and every bias left unrounded: the rounding points of the K = 0 latent-sweep
    kernel, which takes those thin terms from the fp32 image on the CUDA cores.

    quant='bf16' restates what the reference's `mixed_bfloat16` policy (nif/model.py:101-105, 146, 530-533, 954;
    nif/layers/siren.py:519-521: variables cast to the compute dtype, einsum in bf16) permits, at the rounding points of
    the bf16 tensor-core kernels: every matmul operand -- the entries of [w_h; b_h], the activations h, and the latent
    code where it is a matmul operand (layer 0 and the bias sums) -- is rounded to bfloat16; products are accumulated
    in the working precision; the latent contraction, omega_0, the activation and the last layer's bias stay fp32."""
    q = _bf16 if quant in ("bf16", "bf16_main") else (lambda t: t)
    qt = (lambda t: t) if quant == "bf16_main" else q  # "bf16_main": layer 0 and every bias stay fp32 (the K = 0 sweep kernel)
    L = layout(spec.si, spec.so, spec.n, spec.l, spec.variant == "siren_res")
    K1 = z.shape[1] + 1
    W = torch.cat([w_h, b_h[None, :]], 0)  # [K+1, P]
    zt = torch.cat([z, torch.ones(z.shape[0], 1, dtype=z.dtype)], 1)
    Wq, ztq = q(W), qt(zt)
    Wt = qt(W)
    mats = [Wq[:, o:o + a * c].reshape(K1, a, c) for (o, a, c) in L.w]
    mats[0] = Wt[:, L.w[0][0]:L.w[0][0] + L.w[0][1] * L.w[0][2]].reshape(K1, L.w[0][1], L.w[0][2])
    bias = [Wt[:, o:o + c] for (o, c) in L.b]
    sine = spec.variant != "nif"
    f = torch.sin if sine else activation(spec.s_act)
    om = spec.omega0 if sine else 1.0
    H = len(mats) - 2
    # layer 0: the coordinates stay fp32 and multiply the tensor-core result
    d0 = torch.einsum("bk,kij->bij", ztq, mats[0])
    u = f(om * torch.einsum("bi,bij->bj", x, d0) + ztq @ bias[0])
    carry = None
    for m in range(1, H + 1):
        pre = om * torch.einsum("bk,bi,kij->bj", zt, q(u), mats[m]) + ztq @ bias[m]
        if spec.variant == "nif":
            u = f(pre) + u
        elif spec.variant == "siren_res":
            if m % 2 == 1:
                carry, u = u, f(pre)
            else:
                u = 0.5 * (carry + f(pre))
        else:
            u = f(pre)
    bL = W[:, L.b[-1][0]:L.b[-1][0] + L.b[-1][1]]  # the last bias is added on the CUDA cores: not rounded
    return torch.einsum("bk,bkc->bc", zt, torch.einsum("bi,kic->bkc", q(u), mats[-1]) + bL[None])


def shape_net_factored_backward_bf16(spec: Spec, x: Tensor, z: Tensor, w_h: Tensor, b_h: Tensor, du: Tensor):
    """Reverse pass of shape_net_factored(quant='bf16') at the rounding points of the bf16 tensor-core kernels
    (what GradientTape computes for nif/model.py:130-154 / 510-539 under `mixed_bfloat16`, with fp32
    accumulation): the operands of every tensor-core product are rounded to bfloat16 -- da_m, the weight slices, zt
    where it is a matmul operand, the generated operand zt (x) h_m of the weight-gradient GEMM (the PRODUCT is
    rounded) -- while the stashed activations h_m / act'(pre_m), the latent contraction and the thin batch reductions
    (bias rows, first and last matrix: fp32 CUDA-core kernel) are not.  Variants 'nif' and 'siren'.
    Returns (u, dw_h [K,P], db_h [P], dz [B,K])."""
    assert spec.variant in ("nif", "siren")
    q = _bf16
    L = layout(spec.si, spec.so, spec.n, spec.l, False)
    K = z.shape[1]
    W = torch.cat([w_h, b_h[None, :]], 0)
    zt = torch.cat([z, torch.ones(z.shape[0], 1, dtype=z.dtype)], 1)
    Wq, ztq = q(W), q(zt)
    mats = [Wq[:, o:o + a * c].reshape(K + 1, a, c) for (o, a, c) in L.w]
    bias = [Wq[:, o:o + c] for (o, c) in L.b]
    sine = spec.variant != "nif"
    om = spec.omega0 if sine else 1.0
    H = len(mats) - 2

    def act_fd(v):
        if sine:
            return torch.sin(v), torch.cos(v)
        vv = v.detach().clone().requires_grad_(True)
        f = activation(spec.s_act)(vv)
        (d,) = torch.autograd.grad(f.sum(), vv)
        return f.detach(), d

    hs, ds = [], []  # hs[m] = input of matrix m (m = 1..H+1 -> index m-1), ds[m] = act'(pre_m)
    d0 = torch.einsum("bk,kij->bij", ztq, mats[0])
    f, d = act_fd(om * torch.einsum("bi,bij->bj", x, d0) + ztq @ bias[0])
    u = f
    ds.append(d)
    hs.append(u)
    for m in range(1, H + 1):
        pre = om * torch.einsum("bk,bi,kij->bj", zt, q(u), mats[m]) + ztq @ bias[m]
        f, d = act_fd(pre)
        u = f + u if spec.variant == "nif" else f
        ds.append(d)
        hs.append(u)
    bL = W[:, L.b[-1][0]:L.b[-1][0] + L.b[-1][1]]
    y = torch.einsum("bk,bkc->bc", zt, torch.einsum("bi,kic->bkc", q(u), mats[-1]) + bL[None])

    dW = torch.zeros_like(W)
    dzt = torch.zeros_like(zt)
    # last matrix
    dh = torch.einsum("bc,bk,kic->bi", du, ztq, mats[-1])
    dzt += torch.einsum("bc,bkc->bk", du, torch.einsum("bi,kic->bkc", q(hs[H]), mats[-1]) + bL[None])
    o, a_, c_ = L.w[-1]
    dW[:, o:o + a_ * c_] = torch.einsum("bk,bi,bc->kic", zt, hs[H], du).reshape(K + 1, -1)
    o, c_ = L.b[-1]
    dW[:, o:o + c_] = zt.T @ du
    for m in range(H, -1, -1):
        da = dh * ds[m]
        daq = q(da)
        dzt += daq @ bias[m].T
        o, c_ = L.b[m]
        dW[:, o:o + c_] = zt.T @ da
        o, a_, c_ = L.w[m]
        if m >= 1:
            T = torch.einsum("bj,kij->bki", daq, mats[m])
            dh = (dh if spec.variant == "nif" else 0) + om * torch.einsum("bk,bki->bi", zt, T)
            dzt += om * torch.einsum("bki,bi->bk", T, hs[m - 1])
            A = q(zt[:, :, None] * hs[m - 1][:, None, :])  # the generated operand: the product is rounded
            dW[:, o:o + a_ * c_] = om * torch.einsum("bki,bj->kij", A, daq).reshape(K + 1, -1)
        else:
            dzt += om * torch.einsum("bi,bj,kij->bk", x, daq, mats[0])
            dW[:, o:o + a_ * c_] = om * torch.einsum("bk,bi,bj->kij", zt, x, da).reshape(K + 1, -1)
    return y, dW[:K], dW[K], dzt[:, :K]


def shape_net(spec: Spec, x: Tensor, p: Tensor) -> Tensor:
    if spec.variant == "nif":
        return shape_net_nif(x, p, spec.si, spec.so, spec.n, spec.l, spec.s_act)
    return shape_net_mres(
        x, p, spec.variant == "siren_res", spec.omega0, spec.si, spec.so, spec.n, spec.l
    )


# ----------------------------------------------------------------------------
# ParameterNet trunk (everything before the last linear)
# ----------------------------------------------------------------------------
def trunk_param_names(spec: Spec) -> List[str]:
    """Variable names follow the reference layer names (model.py:186-229,
    604-660, 676-733; siren.py:249-254, 370-379, 503-512)."""
    names: List[str] = []
    if spec.variant == "nif":
        names += ["first_dense_pnet/kernel", "first_dense_pnet/bias"]
        for i in range(spec.l_st):
            names += [f"hidden_mlpshortcut_pnet_{i}/kernel", f"hidden_mlpshortcut_pnet_{i}/bias"]
        names += ["bottleneck_pnet/kernel", "bottleneck_pnet/bias"]
        names += ["last_pnet/kernel", "last_pnet/bias"]
        return names
    if spec.p_act == "sine":
        names += ["siren_first_pnet_w", "siren_first_pnet_b"]
        for i in range(spec.l_st):
            if spec.p_resblock:
                q = f"siren_hidden_resblock_pnet_{i}"
                names += [q + "_w", q + "_b", q + "_w2", q + "_b2"]
            else:
                names += [f"siren_hidden_pnet_{i}_w", f"siren_hidden_pnet_{i}_b"]
        names += ["siren_bottleneck_pnet_w", "siren_bottleneck_pnet_b"]
    else:
        names += ["mlp_first_pnet/kernel", "mlp_first_pnet/bias"]
        for i in range(spec.l_st):
            if spec.p_resblock:
                q = f"mlp_hidden_resblock_pnet_{i}"
                names += [q + "_dense_1/kernel", q + "_dense_1/bias", q + "_dense_2/kernel", q + "_dense_2/bias"]
            else:
                names += [f"mlp_hidden_pnet_{i}/kernel", f"mlp_hidden_pnet_{i}/bias"]
        names += ["bottleneck_pnet/kernel", "bottleneck_pnet/bias"]
    names += ["HyperLinearForSIREN_w", "HyperLinearForSIREN_b"]
    return names


def last_layer_names(spec: Spec) -> Tuple[str, str]:
    if spec.variant == "nif":
        return "last_pnet/kernel", "last_pnet/bias"
    return "HyperLinearForSIREN_w", "HyperLinearForSIREN_b"


def latent(spec: Spec, prm: Dict[str, Tensor], p_in: Tensor) -> Tensor:
    """_call_parameter_net up to the bottleneck (nif/model.py:326-343).
    swish trunk: Dense -> l_st x MLP_SimpleShortCut (mlp.py:148-160) or
    MLP_ResNet (mlp.py:62-79) -> linear Dense(latent).
    sine trunk : SIREN first -> SIREN hidden / SIREN_ResNet (siren.py:256-281,
    381-410) -> linear SIREN bottleneck."""
    if spec.variant != "nif" and spec.p_act == "sine":
        w0 = spec.p_omega0
        h = torch.sin(w0 * (p_in @ prm["siren_first_pnet_w"]) + prm["siren_first_pnet_b"])
        for i in range(spec.l_st):
            if spec.p_resblock:
                q = f"siren_hidden_resblock_pnet_{i}"
                g = torch.sin(w0 * (h @ prm[q + "_w"]) + prm[q + "_b"])
                h = 0.5 * (h + torch.sin(w0 * (g @ prm[q + "_w2"]) + prm[q + "_b2"]))
            else:
                q = f"siren_hidden_pnet_{i}"
                h = torch.sin(w0 * (h @ prm[q + "_w"]) + prm[q + "_b"])
        return h @ prm["siren_bottleneck_pnet_w"] + prm["siren_bottleneck_pnet_b"]
    f = activation(spec.p_act)
    first = "first_dense_pnet" if spec.variant == "nif" else "mlp_first_pnet"
    h = f(p_in @ prm[first + "/kernel"] + prm[first + "/bias"])
    for i in range(spec.l_st):
        if spec.variant == "nif":
            q = f"hidden_mlpshortcut_pnet_{i}"
            h = h + f(h @ prm[q + "/kernel"] + prm[q + "/bias"])
        elif spec.p_resblock:
            q = f"mlp_hidden_resblock_pnet_{i}"
            h1 = f(h @ prm[q + "_dense_1/kernel"] + prm[q + "_dense_1/bias"])
            h2 = h1 @ prm[q + "_dense_2/kernel"] + prm[q + "_dense_2/bias"]
            h = f(h + h2)
        else:
            q = f"mlp_hidden_pnet_{i}"
            h = h + f(h @ prm[q + "/kernel"] + prm[q + "/bias"])
    return h @ prm["bottleneck_pnet/kernel"] + prm["bottleneck_pnet/bias"]


def forward(spec: Spec, prm: Dict[str, Tensor], inputs: Tensor) -> Tensor:
    """NIF.call / NIFMultiScale.call (nif/model.py:130-154, 510-539): split the
    (B, pi+si) input by columns, trunk -> last linear -> ShapeNet.  This is the
    *materialised* dataflow of the reference: the (B, po_dim) tensor exists."""
    p_in = inputs[:, : spec.pi]
    x = inputs[:, spec.pi : spec.pi + spec.si]
    wn, bn = last_layer_names(spec)
    z = latent(spec, prm, p_in)
    p = hyper_linear(z, prm[wn], prm[bn])
    return shape_net(spec, x, p)


# ----------------------------------------------------------------------------
# NIFMultiScaleLastLayerParameterized (nif/model.py:989-1269)
# ----------------------------------------------------------------------------
def last_layer_forward(cfg_s: dict, cfg_p: dict, prm: Dict[str, Tensor], inputs: Tensor):
    """call() of the last-layer-parameterised class (nif/model.py:1045-1066): ParameterNet -> pnet_output (po_dim =
    pi_hidden, :583-585); shared-weight SIREN ShapeNet x -> phi [B, so, pi_hidden] (:1222-1242, layers :1151-1220 with
    SIREN / SIREN_ResNet semantics of nif/layers/siren.py:256-281, 381-410); u = Dot(axes=(2,1))(phi, pnet_output) +
    last_layer_bias (:1264-1269).  Returns (u, phi, pnet_output)."""
    pi, si, so = cfg_p["input_dim"], cfg_s["input_dim"], cfg_s["output_dim"]
    K = cfg_p["latent_dim"]
    spec = Spec(variant="siren_res" if cfg_s["use_resblock"] else "siren", pi=pi, si=si, so=so, n=cfg_s["units"],
                l=cfg_s["nlayers"], K=K, n_st=cfg_p["units"], l_st=cfg_p["nlayers"], p_act=cfg_p["activation"],
                omega0=float(cfg_s["omega_0"]), weight_init_factor=cfg_s["weight_init_factor"],
                p_resblock=bool(cfg_p.get("use_resblock", False)), p_omega0=float(cfg_p.get("omega_0", 1.0)))
    z = latent(spec, prm, inputs[:, :pi])
    pout = z @ prm["HyperLinearForSIREN_w"] + prm["HyperLinearForSIREN_b"]
    w0 = float(cfg_s["omega_0"])
    x = inputs[:, pi:pi + si]
    h = torch.sin(w0 * (x @ prm["siren_first_snet_w"]) + prm["siren_first_snet_b"])
    for i in range(cfg_s["nlayers"]):
        if cfg_s["use_resblock"]:
            q = f"siren_hidden_resblock_snet_{i}"
            g = torch.sin(w0 * (h @ prm[q + "_w"]) + prm[q + "_b"])
            h = 0.5 * (h + torch.sin(w0 * (g @ prm[q + "_w2"]) + prm[q + "_b2"]))
        else:
            q = f"siren_hidden_snet_{i}"
            h = torch.sin(w0 * (h @ prm[q + "_w"]) + prm[q + "_b"])
    phi = (h @ prm["siren_bottleneck_snet_w"] + prm["siren_bottleneck_snet_b"]).reshape(-1, so, K)
    u = torch.einsum("bok,bk->bo", phi, pout) + prm["last_layer_bias_snet"]
    return u, phi, pout


# ----------------------------------------------------------------------------
# optimisers of nif/optimizers (restated; the update rules are in the reference tree, the Keras base class is not)
# ----------------------------------------------------------------------------
def adabelief_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float = 1e-3, beta_1: float = 0.9,
                   beta_2: float = 0.999, epsilon: float = 1e-14, weight_decay: float = 0.0, rectify: bool = True,
                   amsgrad: bool = False, vhat: Optional[Tensor] = None, sma_threshold: float = 5.0, total_steps: int = 0,
                   warmup_proportion: float = 0.1, min_lr: float = 0.0) -> None:
    """AdaBeliefOptimizer._resource_apply_dense (nif/optimizers/external_optimizers.py:458-528), in place; `step` is
    iterations + 1 (:467)."""
    lr_t = lr
    if total_steps > 0:  # :470-479
        warmup_steps = total_steps * warmup_proportion
        decay_steps = max(total_steps - warmup_steps, 1)
        decay_rate = (min_lr - lr_t) / decay_steps
        lr_t = lr_t * (step / warmup_steps) if step <= warmup_steps else lr_t + decay_rate * min(step - warmup_steps, decay_steps)
    b1p, b2p = beta_1 ** step, beta_2 ** step
    sma_inf = 2.0 / (1.0 - beta_2) - 1.0
    sma_t = sma_inf - 2.0 * step * b2p / (1.0 - b2p)
    m.mul_(beta_1).add_(g, alpha=1.0 - beta_1)                       # :484-486
    m_corr = m / (1.0 - b1p)
    v.mul_(beta_2).add_((g - m) ** 2, alpha=1.0 - beta_2).add_(epsilon)  # :489-492
    if amsgrad:
        torch.maximum(vhat, v, out=vhat)
        v_corr = torch.sqrt(vhat / (1.0 - b2p))
    else:
        v_corr = torch.sqrt(v / (1.0 - b2p))
    if rectify:
        if sma_t >= sma_threshold:
            r_t = math.sqrt((sma_t - 4.0) / (sma_inf - 4.0) * (sma_t - 2.0) / (sma_inf - 2.0) * sma_inf / sma_t)  # :502-509
            upd = r_t * m_corr / (v_corr + epsilon)
        else:
            upd = m_corr
    else:
        upd = m_corr / (v_corr + epsilon)
    if weight_decay != 0.0:
        upd = upd + weight_decay * p
    p.sub_(lr_t * upd)


def lion_step(p: Tensor, g: Tensor, m: Tensor, lr: float = 1e-4, beta_1: float = 0.9, beta_2: float = 0.99, wd: float = 0.0):
    """Lion._resource_apply_dense (nif/optimizers/external_optimizers.py:681-702), in place."""
    p.sub_(lr * (torch.sign(m * beta_1 + g * (1.0 - beta_1)) + p * wd))
    m.mul_(beta_2).add_(g, alpha=1.0 - beta_2)


def centralize_gradient(g: Tensor) -> Tensor:
    """get_centralized_gradients (nif/optimizers/gtcf.py:27-32): rank >= 2 gradients lose their mean over every axis but
    the last."""
    if g.dim() > 1:
        return g - g.mean(dim=tuple(range(g.dim() - 1)), keepdim=True)
    return g


# ----------------------------------------------------------------------------
# initialisers
# ----------------------------------------------------------------------------
def _trunc_normal(shape, std, gen, dtype):
    t = torch.empty(shape, dtype=torch.float64)
    torch.nn.init.trunc_normal_(t, 0.0, std, -2 * std, 2 * std, generator=gen)
    return t.to(dtype)


def _uniform(shape, bound, gen, dtype):
    b = torch.as_tensor(bound, dtype=torch.float64)
    return ((torch.rand(shape, dtype=torch.float64, generator=gen) * 2 - 1) * b).to(dtype)


def hyper_init_bounds(spec: Spec) -> Tuple[float, np.ndarray]:
    """gen_hypernetwork_weights_bias_for_siren_shapenet (nif/layers/siren.py:6-63):
    returns (bound of w ~ U(+-sqrt(6/K) * factor), per-column bound of b)."""
    L = layout(spec.si, spec.so, spec.n, spec.l, spec.variant == "siren_res")
    n_first = spec.si * spec.n
    n_hidden = (L.w[-1][0]) - n_first
    n_last = spec.so * spec.n
    s = np.ones(L.P, dtype=np.float64)
    s[:n_first] /= spec.si
    s[n_first : n_first + n_hidden] *= math.sqrt(6.0 / spec.n) / spec.omega0
    s[n_first + n_hidden : n_first + n_hidden + n_last] *= math.sqrt(6.0 / (2 * spec.n))
    s[n_first + n_hidden + n_last :] /= spec.n
    return math.sqrt(6.0 / spec.K) * spec.weight_init_factor, s


def init_params(spec: Spec, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Reference initialisers: TruncatedNormal(stddev=0.1) for every Dense
    (model.py:181-182, 222-223, 671-672), SIREN uniform rules (siren.py:178-204),
    SIREN-aware hyper-network init (siren.py:6-63).  The random stream is ours
    (torch Generator); only the distributions follow the reference."""
    g = torch.Generator().manual_seed(seed)
    prm: Dict[str, Tensor] = {}
    P = spec.po_dim

    def dense(name, a, c):
        prm[name + "/kernel"] = _trunc_normal((a, c), 0.1, g, dtype)
        prm[name + "/bias"] = _trunc_normal((c,), 0.1, g, dtype)

    if spec.variant == "nif":
        dense("first_dense_pnet", spec.pi, spec.n_st)
        for i in range(spec.l_st):
            dense(f"hidden_mlpshortcut_pnet_{i}", spec.n_st, spec.n_st)
        dense("bottleneck_pnet", spec.n_st, spec.K)
        dense("last_pnet", spec.K, P)
        return prm

    if spec.p_act == "sine":
        w0 = spec.p_omega0

        def siren(name, a, c, first):
            wb = 1.0 / a if first else math.sqrt(6.0 / a) / w0
            prm[name + "_w"] = _uniform((a, c), wb, g, dtype)
            prm[name + "_b"] = _uniform((c,), 1.0 / math.sqrt(a), g, dtype)

        siren("siren_first_pnet", spec.pi, spec.n_st, True)
        for i in range(spec.l_st):
            if spec.p_resblock:
                q = f"siren_hidden_resblock_pnet_{i}"
                siren(q, spec.n_st, spec.n_st, False)
                # SIREN_ResNet copies w_init/b_init into w2/b2 (siren.py:370-379)
                prm[q + "_w2"] = prm[q + "_w"].clone()
                prm[q + "_b2"] = prm[q + "_b"].clone()
            else:
                siren(f"siren_hidden_pnet_{i}", spec.n_st, spec.n_st, False)
        siren("siren_bottleneck_pnet", spec.n_st, spec.K, False)
    else:
        dense("mlp_first_pnet", spec.pi, spec.n_st)
        for i in range(spec.l_st):
            if spec.p_resblock:
                q = f"mlp_hidden_resblock_pnet_{i}"
                dense(q + "_dense_1", spec.n_st, spec.n_st)
                dense(q + "_dense_2", spec.n_st, spec.n_st)
            else:
                dense(f"mlp_hidden_pnet_{i}", spec.n_st, spec.n_st)
        dense("bottleneck_pnet", spec.n_st, spec.K)
    wb, bb = hyper_init_bounds(spec)
    prm["HyperLinearForSIREN_w"] = _uniform((spec.K, P), wb, g, dtype)
    prm["HyperLinearForSIREN_b"] = _uniform((P,), bb, g, dtype)
    return prm


# ----------------------------------------------------------------------------
# loss / optimiser semantics (third-party in the reference: Keras 2.11)
# ----------------------------------------------------------------------------
def mse(y: Tensor, t: Tensor, sample_weight: Optional[Tensor] = None) -> Tensor:
    """Keras 'mse': mean over the last axis, (weighted) mean over the batch."""
    per_row = ((y - t) ** 2).mean(dim=-1)
    if sample_weight is not None:
        per_row = per_row * sample_weight.reshape(-1)
    return per_row.mean()


def adam_tf(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
            b1: float = 0.9, b2: float = 0.999, eps: float = 1e-7) -> None:
    """tf.keras.optimizers.Adam update, in place (epsilon outside the bias
    correction, default 1e-7).  `step` is 1-based."""
    m.add_((g - m) * (1.0 - b1))
    v.add_((g * g - v) * (1.0 - b2))
    alpha = lr * math.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    p.sub_(alpha * m / (v.sqrt() + eps))


def loss_and_grads(spec: Spec, prm: Dict[str, Tensor], inputs: Tensor, target: Tensor,
                   sample_weight: Optional[Tensor] = None):
    """One reverse pass of the materialised graph; returns (loss, {name: grad}, dL/dz)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in prm.items()}
    p_in = inputs[:, : spec.pi]
    x = inputs[:, spec.pi : spec.pi + spec.si]
    wn, bn = last_layer_names(spec)
    z = latent(spec, leaves, p_in)
    z.retain_grad()
    y = shape_net(spec, x, hyper_linear(z, leaves[wn], leaves[bn]))
    loss = mse(y, target, sample_weight)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return loss.detach(), grads, z.grad.detach(), y.detach()


# ----------------------------------------------------------------------------
# JacobianLayer semantics (nif/layers/gradient.py:36-49, 207-231)
# ----------------------------------------------------------------------------
def jacobian(spec: Spec, prm: Dict[str, Tensor], inputs: Tensor,
             y_index: Sequence[int], x_index: Sequence[int], create_graph: bool = False):
    """(y, J) with J[b,a,c] = d y[b, y_index[a]] / d inputs[b, x_index[c]];
    one reverse pass per output index, then a gather on the input axis."""
    inp = inputs.detach().clone().requires_grad_(True)
    y = forward(spec, prm, inp)
    rows = []
    for i in y_index:
        (g,) = torch.autograd.grad(y[:, i].sum(), inp, create_graph=create_graph, retain_graph=True)
        rows.append(g)
    J = torch.stack(rows, 1)[:, :, list(x_index)]
    return y, J


def hessian(spec: Spec, prm: Dict[str, Tensor], inputs: Tensor,
            y_index: Sequence[int], x_index: Sequence[int]):
    """compute_output_and_grad_and_hessian (gradient.py:234-261): (y, J, H),
    H[b,a,c,d] = d^2 y[b,y_a] / d x_c d x_d."""
    inp = inputs.detach().clone().requires_grad_(True)
    y = forward(spec, prm, inp)
    xi = list(x_index)
    J = []
    H = []
    for i in y_index:
        (g,) = torch.autograd.grad(y[:, i].sum(), inp, create_graph=True, retain_graph=True)
        g = g[:, xi]
        J.append(g)
        hr = []
        for c in range(len(xi)):
            (h,) = torch.autograd.grad(g[:, c].sum(), inp, retain_graph=True)
            hr.append(h[:, xi])
        H.append(torch.stack(hr, 1))
    return y.detach(), torch.stack(J, 1).detach(), torch.stack(H, 1).detach()


def sobolev_loss(y3: Tensor, t3: Tensor, coef_grad: float) -> Tensor:
    """Sobolov_MSE of tutorial 8 (tutorial/8_...ipynb:815-820): columns are
    [u, du/dt, du/dx]; only u and du/dx enter the loss."""
    sd = (t3[:, 0] - y3[:, 0]) ** 2
    sg = (t3[:, 2] - y3[:, 2]) ** 2
    return sd.mean() + coef_grad * sg.mean()


def sobolev_loss_and_grads(spec: Spec, prm: Dict[str, Tensor], inputs: Tensor, target_u: Tensor, target_g: Tensor,
                           x_col: int, coef_grad: float):
    """Gradient of  mean_b mean_c (u - t_u)^2 + coef_grad * mean_b mean_c (du/dinputs[:, x_col] - t_g)^2  w.r.t. every
    parameter: what Keras' tape computes for the tutorial-8 model (JacobianLayer output concatenated into the model
    output, Sobolov_MSE on it, tutorial/8_...ipynb:809-820), i.e. reverse mode through compute_output_and_grad
    (nif/layers/gradient.py:207-231).  Returns (loss, {name: grad}, dL/dz, u, du/dx_col)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in prm.items()}
    inp = inputs.detach().clone().requires_grad_(True)
    p_in = inp[:, : spec.pi]
    x = inp[:, spec.pi : spec.pi + spec.si]
    wn, bn = last_layer_names(spec)
    z = latent(spec, leaves, p_in)
    y = shape_net(spec, x, hyper_linear(z, leaves[wn], leaves[bn]))
    cols = []
    for c in range(y.shape[1]):  # one reverse pass per output index, like the reference
        (g,) = torch.autograd.grad(y[:, c].sum(), inp, create_graph=True, retain_graph=True)
        cols.append(g[:, x_col])
    dy = torch.stack(cols, 1)
    loss = ((y - target_u) ** 2).mean(-1).mean() + coef_grad * ((dy - target_g) ** 2).mean(-1).mean()
    # (torch.autograd.grad, not .backward(): a retain_grad() hook on z would also collect the inner reverse passes)
    names = list(leaves)
    got = torch.autograd.grad(loss, [z] + [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, got[1:])}
    return loss.detach(), grads, got[0].detach(), y.detach(), dy.detach()


def sobolev_loss_and_grads_pairs(spec: Spec, prm: Dict[str, Tensor], inputs: Tensor, target_u: Tensor, target_g: Tensor,
                                 pairs, coef_grad: float):
    """The same tape with any set of Jacobian entries in the loss: pairs = [(output index, input column)], target_g
    [B, len(pairs)]; input columns < spec.pi are ParameterNet inputs (du/dt of tutorial 8's JacobianLayer output, which
    the tutorial only monitors but Keras would differentiate just the same).
        loss = mean_b mean_c (u - t_u)^2 + coef_grad * mean_b mean_pairs (du_y/dinput_c - t_g)^2
    Returns (loss, {name: grad}, u, [B, len(pairs)] derivatives)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in prm.items()}
    inp = inputs.detach().clone().requires_grad_(True)
    p_in = inp[:, : spec.pi]
    x = inp[:, spec.pi : spec.pi + spec.si]
    wn, bn = last_layer_names(spec)
    z = latent(spec, leaves, p_in)
    y = shape_net(spec, x, hyper_linear(z, leaves[wn], leaves[bn]))
    rows = {}
    for yc in sorted({a for a, _ in pairs}):  # one reverse pass per output index, like the reference
        (rows[yc],) = torch.autograd.grad(y[:, yc].sum(), inp, create_graph=True, retain_graph=True)
    dy = torch.stack([rows[a][:, c] for a, c in pairs], 1)
    loss = ((y - target_u) ** 2).mean(-1).mean() + coef_grad * ((dy - target_g) ** 2).mean(-1).mean()
    names = list(leaves)
    got = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, got)}
    return loss.detach(), grads, y.detach(), dy.detach()


def jacobian_model_loss_and_grads(spec: Spec, prm: Dict[str, Tensor], inputs: Tensor, y_true: Tensor, loss_fn, y_index,
                                  x_index):
    """Any loss on the output of a JacobianLayer-wrapped model (README.md:119-131: Model([x], [y, dydx]); a PDE residual
    is such a loss): y_pred = [y | dy[y_index]/dx[x_index] flattened], loss_fn(y_true, y_pred) -> scalar, differentiated by
    the outer tape with respect to every parameter.  Returns (loss, {name: grad}, y_pred)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in prm.items()}
    inp = inputs.detach().clone().requires_grad_(True)
    y = forward(spec, leaves, inp)
    cols = []
    for a in y_index:  # one reverse pass per output index (nif/layers/gradient.py:207-231)
        (g,) = torch.autograd.grad(y[:, a].sum(), inp, create_graph=True, retain_graph=True)
        cols.append(g[:, list(x_index)])
    y_pred = torch.cat([y] + cols, -1)
    loss = loss_fn(y_true, y_pred)
    names = list(leaves)
    got = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, got)}
    return loss.detach(), grads, y_pred.detach()


# ----------------------------------------------------------------------------
# a whole training step in the reference's materialised dataflow (CPU baseline)
# ----------------------------------------------------------------------------
def jac_reg_loss(spec: Spec, prm: Dict[str, Tensor], inputs: Tensor, l1: float) -> Tensor:
    """The add_loss term of JacRegLatentLayer (nif/layers/gradient.py:52-113) as NIF.build() wires it
    (nif/model.py:353-375): y_index = range(latent_dim), x_index = range(pi_dim) on the augmented model whose second
    output is the latent code, so  l1 * reduce_mean(square(d latent / d input_p))  over (batch, latent, pi).
    compute_output_and_augment_grad (gradient.py:183-204) takes the batch Jacobian w.r.t. the whole input and gathers
    the ParameterNet columns; the latent code does not depend on the other columns."""
    p_in = inputs[:, : spec.pi].detach().clone().requires_grad_(True)
    z = latent(spec, prm, p_in)
    rows = []
    for k in range(z.shape[1]):
        (g,) = torch.autograd.grad(z[:, k].sum(), p_in, create_graph=True, retain_graph=True)
        rows.append(g)
    J = torch.stack(rows, 1)  # [B, K, pi]
    return l1 * (J * J).mean()


class MaterialisedTrainer:
    """fit()-equivalent inner loop on CPU: forward with the (B, po_dim) tensor
    materialised, reverse-mode gradient, TF-semantics Adam.  Used as the
    'TF2-CPU proxy' baseline."""

    def __init__(self, spec: Spec, prm: Dict[str, Tensor], lr: float = 1e-3, jac_reg: Optional[float] = None):
        self.jac_reg = jac_reg
        self.spec = spec
        self.prm = {k: v.detach().clone().requires_grad_(True) for k, v in prm.items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.prm.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.prm.items()}
        self.t = 0
        self.lr = lr

    def step(self, inputs: Tensor, target: Tensor, sample_weight: Optional[Tensor] = None) -> float:
        for p in self.prm.values():
            p.grad = None
        y = forward(self.spec, self.prm, inputs)
        loss = mse(y, target, sample_weight)
        if self.jac_reg is not None:
            loss = loss + jac_reg_loss(self.spec, self.prm, inputs, self.jac_reg)
        loss.backward()
        self.t += 1
        with torch.no_grad():
            for k, p in self.prm.items():
                adam_tf(p, p.grad, self.m[k], self.v[k], self.t, self.lr)
        return float(loss.detach())


# ----------------------------------------------------------------------------
# bundled datasets, regenerated analytically (tutorial/1 cell 3, tutorial/2 cell 3)
# ----------------------------------------------------------------------------
def traveling_wave_raw(omega: float) -> np.ndarray:
    """(2000, 3) float32 [t, x, u]: 10 t x 200 x, u = exp(-1000 s^2) sin(omega s),
    s = x - 0.2 - 0.006 t.  omega = 4 -> traveling_wave.npz, 400 -> ..._high_freq.npz."""
    x = np.linspace(0, 1, 200, endpoint=False)
    t = np.linspace(0, 100, 10, endpoint=False)
    xx, tt = np.meshgrid(x, t)
    s = xx - 0.2 - (0.12 / 20) * tt
    u = np.exp(-1000 * s**2) * np.sin(omega * s)
    return np.stack([tt.ravel(), xx.ravel(), u.ravel()], 1).astype(np.float32)


def standard_normalize(raw: np.ndarray):
    """PointWiseData.standard_normalize (nif/data/point_wise_data.py:51-78)."""
    mean, std = raw.mean(0), raw.std(0)
    return (raw - mean) / std, mean, std


def minmax_normalize(raw: np.ndarray, n_para: int, n_x: int, n_target: int):
    """PointWiseData.minmax_normalize (nif/data/point_wise_data.py:81-114)."""
    mean, std = raw.mean(0), raw.std(0)
    for i in range(n_para + n_x):
        lo, hi = raw[:, i].min(), raw[:, i].max()
        mean[i], std[i] = 0.5 * (lo + hi), 0.5 * (hi - lo)
    for j in range(n_para + n_x, n_para + n_x + n_target):
        std[j] = np.abs(raw[:, j]).max()
    return (raw - mean) / std, mean, std
