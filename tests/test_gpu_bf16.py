"""bf16 tensor-core path (dtype_compute = 1, the arithmetic `mixed_bfloat16` permits: nif/model.py:101-105, 146,
530-533, 954) against the oracle, through the C ABI.

Two references, two gates:
  * the oracle's bf16 restatement (oracle.shape_net_factored(quant='bf16') / shape_net_factored_backward_bf16), which
    rounds the same matmul operands to bfloat16 and accumulates in fp64: gate 1e-2 on max|a-b|/max|b|.  What is left is
    fp32 accumulation order plus the occasional operand whose rounding flips (one flip = 2^-8 relative on one element);
    measured 1e-7 .. 5e-3 on these shapes.
  * the exact fp64 oracle: gate 0.25 for SIREN (omega_0 = 30 amplifies every bf16 rounding by the pre-activation scale;
    measured 2e-2 .. 1.8e-1 at 6 x 128), 2e-2 for the swish NIF.
"""
import numpy as np
import pytest
import torch

from oracle import nif_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

GATE_EMU = 1e-2   # forward
GATE_EMU_GRAD = 3e-2  # reverse pass: the forward deviations pass through omega_0-scaled layers once more


def _setup(variant, si, so, n, l, K, B, seed=5):
    spec = O.Spec(variant=variant, pi=1, si=si, so=so, n=n, l=l, K=max(K, 1), n_st=16, l_st=1, p_act="swish", omega0=30.0,
                  weight_init_factor=0.01, s_act="swish")
    prm = {k: v.double() for k, v in O.init_params(spec, 1).items()}
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(B, si, generator=g, dtype=torch.float64) * 2 - 1).float().double()
    z = (torch.rand(B, max(K, 1), generator=g, dtype=torch.float64) - 0.5).float().double()
    tgt = (torch.rand(B, so, generator=g, dtype=torch.float64) * 2 - 1).float().double()
    wn, bn = O.last_layer_names(spec)
    return spec, prm[wn].float().double(), prm[bn].float().double(), x, z, tgt


def _engine(variant, si, so, n, l, K):
    from nif_b200.ops import FusedShapeNet
    eng = FusedShapeNet(variant, si, so, n, l, K, "swish", 30.0, compute="bf16")
    assert eng.kernel_path == "bf16", "this descriptor must be served by the bf16 tcgen05 kernels, not a fall-through"
    return eng


CASES = [
    ("siren", 2, 1, 64, 2, 3, 300),     # NP = 64: two latent coordinates per chunk, even K + 1
    ("siren", 3, 3, 128, 2, 5, 300),    # NP = 128, ragged batch
    ("siren", 2, 1, 64, 4, 32, 1000),   # C2 shape
    ("siren", 3, 3, 128, 6, 64, 700),   # C3 shape (BASELINE.json configs[2])
    ("nif", 2, 2, 48, 3, 7, 90),        # swish + residual, padded width 64
    ("nif", 2, 2, 100, 2, 6, 257),      # swish + residual, padded width 128
    ("siren", 1, 1, 128, 1, 1, 20000),  # many tiles per CTA (persistent loop, barrier phases)
]


@pytest.mark.parametrize("variant,si,so,n,l,K,B", CASES)
def test_bf16_forward_and_reverse(variant, si, so, n, l, K, B):
    dev = torch.device("cuda:0")
    spec, w_h, b_h, x, z, tgt = _setup(variant, si, so, n, l, K, B)
    eng = _engine(variant, si, so, n, l, K)
    zd, xd = z.float().to(dev), x.float().to(dev)
    packed = eng.pack(w_h.float().to(dev), b_h.float().to(dev))
    u = eng.forward(zd, xd, packed)
    u2, stash = eng.forward(zd, xd, packed, save=True)
    assert torch.equal(u, u2)
    ref = O.shape_net_factored(spec, x, z, w_h, b_h)
    gate64 = 0.25 if variant != "nif" else 2e-2
    # reverse pass: Keras 'mse' seed from the kernel's own outputs, so that both sides differentiate the same point
    du = 2.0 * (u.double().cpu() - tgt) / (B * so)
    emu, dw_e, db_e, dz_e = O.shape_net_factored_backward_bf16(spec, x, z, w_h, b_h, du)
    assert rel_err(u.cpu(), emu) < GATE_EMU and rel_err(u.cpu(), ref) < gate64
    loss = torch.zeros(1, device=dev)
    dw = torch.empty(K, eng.po_dim, device=dev)
    db = torch.empty(eng.po_dim, device=dev)
    dz = eng.mse_backward(zd, xd, packed, u2, stash, tgt.float().to(dev), None, 1.0 / B, loss, dw, db)
    loss_ref = float(((u.double().cpu() - tgt) ** 2).mean(-1).mean())
    assert abs(float(loss) - loss_ref) <= 1e-5 * abs(loss_ref)
    for name, got, want in (("dw_h", dw, dw_e), ("db_h", db, db_e), ("dz", dz, dz_e)):
        e = rel_err(got.cpu(), want)
        assert e < GATE_EMU_GRAD, f"{name}: {e:.3e} against the bf16 restatement"
    # and against exact fp64 differentiation (loose: bf16 operands)
    zq, wq, bq = z.clone().requires_grad_(True), w_h.clone().requires_grad_(True), b_h.clone().requires_grad_(True)
    ((O.shape_net_factored(spec, x, zq, wq, bq) - tgt) ** 2).mean(-1).mean().backward()
    for name, got, want in (("dw_h", dw, wq.grad), ("db_h", db, bq.grad), ("dz", dz, zq.grad)):
        assert rel_err(got.cpu(), want) < 1.5 * gate64, name
    # accumulate semantics
    eng.mse_backward(zd, xd, packed, u2, stash, tgt.float().to(dev), None, 1.0 / B, loss, dw, db, 1.0)
    assert rel_err(dw.cpu(), 2 * dw_e) < GATE_EMU_GRAD and rel_err(db.cpu(), 2 * db_e) < GATE_EMU_GRAD


def test_bf16_reverse_is_deterministic_and_row_local():
    """size-independent properties at a batch of many tiles: bitwise repeatability, and rows are independent (the result
    for a row does not depend on which tile / CTA it lands in)."""
    dev = torch.device("cuda:0")
    variant, si, so, n, l, K, B = "siren", 3, 3, 128, 3, 16, 40000
    spec, w_h, b_h, x, z, tgt = _setup(variant, si, so, n, l, K, B)
    eng = _engine(variant, si, so, n, l, K)
    zd, xd, td = z.float().to(dev), x.float().to(dev), tgt.float().to(dev)
    packed = eng.pack(w_h.float().to(dev), b_h.float().to(dev))

    def run(zz, xx, tt):
        u, stash = eng.forward(zz, xx, packed, save=True)
        loss = torch.zeros(1, device=dev)
        dw, db = torch.empty(K, eng.po_dim, device=dev), torch.empty(eng.po_dim, device=dev)
        dz = eng.mse_backward(zz, xx, packed, u, stash, tt, None, 1.0 / B, loss, dw, db)
        return u, dz, dw, db, loss

    a, b = run(zd, xd, td), run(zd, xd, td)
    for p, q in zip(a, b):
        assert torch.equal(p, q)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).to(dev)
    c = run(zd[perm].contiguous(), xd[perm].contiguous(), td[perm].contiguous())
    assert torch.equal(c[0], a[0][perm]) and torch.equal(c[1], a[1][perm])
    assert rel_err(c[2].cpu(), a[2].cpu()) < 1e-5 and rel_err(c[3].cpu(), a[3].cpu()) < 1e-5


def test_bf16_grouped_latent_by_grid():
    """C5 in miniature (model_x_to_u_given_w in factored form): G explicit weight vectors x one shared grid."""
    dev = torch.device("cuda:0")
    variant, si, so, n, l, B, G = "siren", 3, 1, 128, 3, 500, 3
    spec, w_h, b_h, x, z, _ = _setup(variant, si, so, n, l, 1, B)
    zg = torch.rand(G, 1, generator=torch.Generator().manual_seed(2), dtype=torch.float64) - 0.5
    wv = (zg @ w_h + b_h).float().double()
    eng = _engine(variant, si, so, n, l, 0)
    packed = eng.pack(None, wv.float().to(dev))
    u = eng.forward(None, x.float().to(dev), packed, groups=G, x_shared=True).view(G, B, so)
    xs = x.float().to(dev).repeat(G, 1)
    u_own = eng.forward(None, xs, packed, groups=G, x_shared=False).view(G, B, so)
    assert torch.equal(u, u_own)
    z0, w0 = torch.zeros(B, 0, dtype=torch.float64), torch.zeros(0, wv.shape[1], dtype=torch.float64)
    for gi in range(G):
        emu = O.shape_net_factored(spec, x, z0, w0, wv[gi], quant="bf16_main")  # the sweep kernel's rounding points
        ref = O.shape_net_factored(spec, x, z0, w0, wv[gi])
        assert rel_err(u[gi].cpu(), emu) < GATE_EMU and rel_err(u[gi].cpu(), ref) < 0.25


def test_mixed_bfloat16_policy_selects_the_bf16_kernels_and_trains():
    import nif_b200
    dev = torch.device("cuda:0")
    cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 2, "output_dim": 1, "units": 128, "nlayers": 2,
             "weight_init_factor": 0.01, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 8, "units": 32, "nlayers": 2, "activation": "swish"}
    net = nif_b200.NIFMultiScale(cfg_s, cfg_p, "mixed_bfloat16", seed=0, device=dev)
    assert net.engine.kernel_path == "bf16" and net.compute_Dtype == "bfloat16" and net.variable_Dtype == "float32"
    model = net.build()
    model.compile(nif_b200.Adam(1e-3), loss="mse")
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (2048, 3)).astype(np.float32)
    Y = np.sin(3 * X[:, :1]) * np.cos(2 * X[:, 1:2] + X[:, 2:3])
    first = model.train_on_batch(X, Y)
    for _ in range(60):
        last = model.train_on_batch(X, Y)
    assert np.isfinite(last) and last < 0.5 * first
    # the float32 policy on the same architecture runs elsewhere (fp32-grade kernels)
    net32 = nif_b200.NIFMultiScale(cfg_s, cfg_p, "float32", seed=0, device=dev)
    assert net32.engine.kernel_path != "bf16"


@pytest.mark.parametrize("policy", ["mixed_bfloat16", "float32"])
def test_torch_trunk_fast_path_equals_the_composite_ops(policy):
    """A 128-wide ParameterNet trunk runs as torch ops.  Where it is differentiated once, swish is the single silu kernel
    and the Dense layers go through model._TrunkLinear (bias gradient as a GEMV, bf16 operand copies kept): same latent code
    and the same gradients as the composite expression under the same policy."""
    import nif_b200
    dev = torch.device("cuda:0")
    cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 2, "output_dim": 1, "units": 64, "nlayers": 2,
             "weight_init_factor": 0.01, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 16, "units": 128, "nlayers": 3, "activation": "swish"}
    net = nif_b200.NIFMultiScale(cfg_s, cfg_p, policy, seed=0, device=dev)
    assert net._trunk is None  # not one of the fused trunk shapes
    g = torch.Generator().manual_seed(3)
    p_in = (torch.rand(3000, 1, generator=g) * 2 - 1).to(dev)
    dz = torch.randn(3000, 16, generator=g).to(dev)
    out = {}
    for fast in (True, False):
        net.grad.zero_()
        net._first_order_only = fast
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(policy == "mixed_bfloat16")):
                z = net._latent(p_in)
        finally:
            net._first_order_only = False
        z.float().backward(dz)
        out[fast] = (z.detach().float().clone(), net.grad.clone())
    # bf16: an fp32 difference of one ulp between silu(x) and x * sigmoid(x) can flip the bf16 rounding of an activation (one
    # bf16 ulp = 4e-3 of that element) ahead of the next matmul: the two paths agree to bf16 rounding noise, not bit for bit
    tol = GATE_EMU if policy == "mixed_bfloat16" else 1e-5
    assert rel_err(out[True][0].cpu(), out[False][0].cpu()) < tol
    assert rel_err(out[True][1].cpu(), out[False][1].cpu()) < tol
