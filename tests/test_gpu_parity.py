"""Parity of the CUDA hot path against the oracle, through the C ABI (ctypes).

Tolerance (BASELINE.md section 5 / north_star "within 1e-5 rel-fp32"):
    err = max|a - b| / max|b|  per tensor, against the fp64 oracle;
    gate  err_gpu <= max(1e-5, 2 * err_cpu_fp32)
where err_cpu_fp32 is the error of the fp32 CPU oracle against the same fp64 truth.
"""
import numpy as np
import pytest
import torch

from oracle import nif_oracle as O
from tests.helpers import golden_cases, load_golden, rel_err

pytestmark = pytest.mark.gpu


def _engine(spec):
    from nif_b200.ops import FusedShapeNet
    return FusedShapeNet(spec.variant, spec.si, spec.so, spec.n, spec.l, spec.K, spec.s_act, spec.omega0)


def _gate(err_gpu, err_cpu32, floor=1e-5):
    return err_gpu <= max(floor, 2.0 * err_cpu32)


def _run_case(spec, prm64, inputs64, target64, sw64, floor=1e-5):
    """forward + mse backward on the GPU vs fp64 oracle (and fp32 oracle for the gate)."""
    dev = torch.device("cuda:0")
    wn, bn = O.last_layer_names(spec)
    loss64, g64, gz64, y64 = O.loss_and_grads(spec, prm64, inputs64, target64, sw64)
    prm32 = {k: v.float() for k, v in prm64.items()}
    loss32, g32, gz32, y32 = O.loss_and_grads(spec, prm32, inputs64.float(), target64.float(),
                                              None if sw64 is None else sw64.float())
    z64 = O.latent(spec, prm64, inputs64[:, : spec.pi])
    eng = _engine(spec)
    z = z64.float().to(dev)
    x = inputs64[:, spec.pi: spec.pi + spec.si].float().contiguous().to(dev)
    w_h, b_h = prm64[wn].float().to(dev), prm64[bn].float().to(dev)
    packed = eng.pack(w_h, b_h)
    u, stash = eng.forward(z, x, packed, save=True)
    u_inf = eng.forward(z, x, packed)
    assert torch.equal(u, u_inf), "stash on/off must not change the output"
    e_u = rel_err(u.cpu(), y64)
    assert _gate(e_u, rel_err(y32, y64), floor), f"forward err {e_u:.3e} (cpu32 {rel_err(y32, y64):.3e})"
    B = inputs64.shape[0]
    loss = torch.zeros(1, device=dev)
    dw = torch.full_like(w_h, float("nan"))
    db = torch.full_like(b_h, float("nan"))
    dz = eng.mse_backward(z, x, packed, u, stash, target64.float().to(dev),
                          None if sw64 is None else sw64.float().to(dev), 1.0 / B, loss, dw, db, 0.0)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(loss64)) <= 1e-5 * max(1.0, abs(float(loss64)))
    for name, got, ref64, ref32 in (("dz", dz, gz64, gz32), ("dw_h", dw, g64[wn], g32[wn]), ("db_h", db, g64[bn], g32[bn])):
        e = rel_err(got.cpu(), ref64)
        assert _gate(e, rel_err(ref32, ref64), floor), f"{name} err {e:.3e} (cpu32 {rel_err(ref32, ref64):.3e})"
    # accumulate semantics: beta = 1 doubles the gradient
    dz2 = eng.mse_backward(z, x, packed, u, stash, target64.float().to(dev),
                           None if sw64 is None else sw64.float().to(dev), 1.0 / B, loss, dw, db, 1.0)
    assert rel_err(dw.cpu(), 2 * g64[wn]) < 1e-4 and rel_err(db.cpu(), 2 * g64[bn]) < 1e-4
    return eng, packed


@pytest.mark.parametrize("case", golden_cases())
def test_golden_forward_backward(case):
    d, cls, cfg_s, cfg_p, spec, prm, grads = load_golden(case)
    if spec.variant == "nif" and spec.s_act not in ("swish", "tanh", "relu", "sigmoid", "sine", "linear"):
        pytest.skip("activation not in the fused set")
    _run_case(spec, prm, torch.as_tensor(d["inputs"]), torch.as_tensor(d["target"]), torch.as_tensor(d["sample_weight"]))
    # and against the committed reference outputs themselves
    dev = torch.device("cuda:0")
    eng = _engine(spec)
    wn, bn = O.last_layer_names(spec)
    packed = eng.pack(prm[wn].float().to(dev), prm[bn].float().to(dev))
    x = torch.as_tensor(d["inputs"])[:, spec.pi:].float().contiguous().to(dev)
    u = eng.forward(torch.as_tensor(d["latent"]).float().to(dev), x, packed)
    assert rel_err(u.cpu(), d["y"]) <= max(1e-5, 2 * rel_err(d["y32"], d["y"]))


def _random_problem(variant, si, so, n, l, K, B, seed, act="swish", omega0=30.0, wif=0.05):
    if variant == "nif":
        spec = O.Spec(variant="nif", pi=1, si=si, so=so, n=n, l=l, K=K, n_st=16, l_st=1, p_act="swish", s_act=act)
    else:
        spec = O.Spec(variant=variant, pi=1, si=si, so=so, n=n, l=l, K=K, n_st=16, l_st=1, p_act="swish",
                      omega0=omega0, weight_init_factor=wif)
    prm = {k: v.double() for k, v in O.init_params(spec, seed).items()}
    if variant == "nif":
        # TruncatedNormal(0.1) on a wide last layer saturates swish; scale to keep the test well conditioned
        prm["last_pnet/kernel"] *= 0.5
    rng = np.random.default_rng(seed)
    inputs = torch.as_tensor(rng.uniform(-1, 1, (B, 1 + si)))
    target = torch.as_tensor(rng.uniform(-1, 1, (B, so)))
    sw = torch.as_tensor(rng.uniform(0.5, 1.5, (B,)))
    return spec, prm, inputs, target, sw


@pytest.mark.parametrize(
    "variant,si,so,n,l,K,B",
    [
        ("siren", 2, 1, 64, 4, 32, 300),      # C2 shape, ragged batch
        ("siren", 2, 1, 64, 4, 32, 128),      # exactly one tile
        ("siren", 1, 1, 64, 4, 32, 257),      # C4 shape
        ("siren", 3, 3, 128, 2, 8, 200),      # width 128 path
        ("siren", 3, 1, 100, 1, 5, 77),       # padded to 128
        ("siren_res", 2, 2, 64, 2, 6, 150),   # res-blocks
        ("siren_res", 1, 1, 20, 3, 2, 130),
        ("nif", 1, 1, 30, 2, 1, 512),         # C1
        ("nif", 2, 2, 48, 3, 7, 90),
        ("siren", 1, 1, 30, 0, 3, 64),        # no hidden layer
        ("siren", 2, 1, 64, 1, 70, 140),      # latent > 68 (two edge passes)
        ("siren", 1, 1, 8, 2, 1, 1),          # single row
    ],
)
def test_random_forward_backward(variant, si, so, n, l, K, B):
    spec, prm, inputs, target, sw = _random_problem(variant, si, so, n, l, K, B, seed=si * 100 + n + K)
    _run_case(spec, prm, inputs, target, sw)


def test_no_sample_weight_and_large_batch_splits():
    spec, prm, inputs, target, _ = _random_problem("siren", 2, 1, 64, 2, 4, 5000, seed=5)
    _run_case(spec, prm, inputs, target, None)


def test_autograd_bridge_matches_oracle():
    from nif_b200.ops import fused_shapenet
    spec, prm, inputs, target, sw = _random_problem("siren", 2, 1, 32, 2, 3, 100, seed=9)
    dev = torch.device("cuda:0")
    wn, bn = O.last_layer_names(spec)
    eng = _engine(spec)
    z = O.latent(spec, prm, inputs[:, :1]).float().to(dev).requires_grad_(True)
    x = inputs[:, 1:].float().contiguous().to(dev)
    w = prm[wn].float().to(dev).requires_grad_(True)
    b = prm[bn].float().to(dev).requires_grad_(True)
    u = fused_shapenet(z, x, w, b, eng)
    loss = ((u - target.float().to(dev)) ** 2).mean()
    loss.backward()
    l64, g64, gz64, _ = O.loss_and_grads(spec, prm, inputs, target, None)
    assert abs(float(loss) - float(l64)) < 1e-5
    assert rel_err(w.grad.cpu(), g64[wn]) < 1e-4 and rel_err(b.grad.cpu(), g64[bn]) < 1e-4
    assert rel_err(z.grad.cpu(), gz64) < 1e-4


@pytest.mark.parametrize("variant,n,l", [("siren", 64, 4), ("siren_res", 30, 2), ("nif", 30, 2), ("siren", 128, 1)])
def test_given_w_matches_oracle(variant, n, l):
    spec, prm, inputs, _, _ = _random_problem(variant, 2, 2, n, l, 3, 333, seed=n + l)
    dev = torch.device("cuda:0")
    wn, bn = O.last_layer_names(spec)
    z = O.latent(spec, prm, inputs[:, :1])
    w = O.hyper_linear(z, prm[wn], prm[bn])  # (B, P) per-row weights, the reference's materialised tensor
    y64 = O.shape_net(spec, inputs[:, 1:], w)
    y32 = O.shape_net(spec, inputs[:, 1:].float(), w.float())
    eng = _engine(spec)
    u = eng.given_w(inputs[:, 1:].float().contiguous().to(dev), w.float().to(dev))
    assert _gate(rel_err(u.cpu(), y64), rel_err(y32, y64))


def test_grouped_inference_latent_by_grid():
    """C5 shape in miniature: G latents x N shared grid points, weights packed per latent (K=0 heads)."""
    spec, prm, inputs, _, _ = _random_problem("siren", 3, 1, 64, 2, 6, 50, seed=1)
    dev = torch.device("cuda:0")
    wn, bn = O.last_layer_names(spec)
    G, N = 5, 333
    rng = np.random.default_rng(0)
    zg = torch.as_tensor(rng.normal(size=(G, spec.K)))
    grid = torch.as_tensor(rng.uniform(-1, 1, (N, spec.si)))
    wg = O.hyper_linear(zg, prm[wn], prm[bn])  # (G, P)
    ref = torch.stack([O.shape_net(spec, grid, wg[g: g + 1].expand(N, -1)) for g in range(G)])  # (G, N, so)
    ref32 = torch.stack([O.shape_net(spec, grid.float(), wg[g: g + 1].float().expand(N, -1)) for g in range(G)])
    eng0 = _engine(spec).with_latent(0)
    packed = eng0.pack(None, wg.float().to(dev))
    u = eng0.forward(None, grid.float().to(dev), packed, groups=G, x_shared=True).reshape(G, N, spec.so)
    assert _gate(rel_err(u.cpu(), ref), rel_err(ref32, ref))


def test_adam_matches_tf_semantics():
    from nif_b200.ops import adam_step
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    n = 100003
    p = torch.as_tensor(rng.normal(size=n)).float()
    m = torch.zeros(n)
    v = torch.zeros(n)
    pg, mg, vg = p.to(dev), m.to(dev), v.to(dev)
    p64, m64, v64 = p.double(), m.double(), v.double()
    for t in range(1, 4):
        g = torch.as_tensor(rng.normal(size=n)).float()
        adam_step(pg, g.to(dev), mg, vg, 1e-3, t)
        O.adam_tf(p64, g.double(), m64, v64, t, 1e-3)
    assert rel_err(pg.cpu(), p64) < 1e-6 and rel_err(mg.cpu(), m64) < 1e-6 and rel_err(vg.cpu(), v64) < 1e-6


def test_empty_batch_and_bad_arguments():
    from nif_b200 import _lib
    from nif_b200.ops import FusedShapeNet
    dev = torch.device("cuda:0")
    eng = FusedShapeNet("siren", 2, 1, 16, 1, 2, omega0=30.0)
    packed = eng.pack(torch.zeros(2, eng.po_dim, device=dev), torch.zeros(eng.po_dim, device=dev))
    u = eng.forward(torch.zeros(0, 2, device=dev), torch.zeros(0, 2, device=dev), packed)
    assert u.shape == (0, 1)
    with pytest.raises(_lib.NifError):
        eng.pack(torch.zeros(3, eng.po_dim, device=dev), torch.zeros(eng.po_dim, device=dev))
    with pytest.raises(_lib.NifError):
        FusedShapeNet("siren", 2, 1, 500, 1, 2)  # units > 128: rejected by the library
    with pytest.raises(_lib.NifError):
        eng.forward(torch.zeros(4, 2), torch.zeros(4, 2), packed)  # CPU tensors: no fallback


def test_full_size_properties_c2():
    """BASELINE config 2 at full batch: size-independent checks (determinism, row independence,
    gradient linearity in the seed)."""
    spec, prm, _, _, _ = _random_problem("siren", 2, 1, 64, 4, 32, 8, seed=2)
    dev = torch.device("cuda:0")
    wn, bn = O.last_layer_names(spec)
    eng = _engine(spec)
    B = 65536
    g = torch.Generator(device="cpu").manual_seed(0)
    z = (torch.rand(B, 32, generator=g) - 0.5).to(dev)
    x = (torch.rand(B, 2, generator=g) * 2 - 1).to(dev)
    w_h, b_h = prm[wn].float().to(dev), prm[bn].float().to(dev)
    packed = eng.pack(w_h, b_h)
    u1, stash = eng.forward(z, x, packed, save=True)
    u2 = eng.forward(z, x, packed)
    assert torch.equal(u1, u2)
    # row independence: a permuted batch gives the permuted output, bit for bit
    perm = torch.randperm(B, generator=g).to(dev)
    assert torch.equal(eng.forward(z[perm], x[perm], packed), u1[perm])
    # oracle on a sample of rows
    idx = torch.arange(0, B, 997)
    y64 = O.shape_net(spec, x[idx].cpu().double(), O.hyper_linear(z[idx].cpu().double(), prm[wn], prm[bn]))
    y32 = O.shape_net(spec, x[idx].cpu(), O.hyper_linear(z[idx].cpu(), prm[wn].float(), prm[bn].float()))
    assert _gate(rel_err(u1[idx].cpu(), y64), rel_err(y32, y64))
    # reverse pass: deterministic, and linear in the seed
    du = torch.randn(B, 1, generator=g).to(dev)
    dw1, db1 = torch.empty_like(w_h), torch.empty_like(b_h)
    dw2, db2 = torch.empty_like(w_h), torch.empty_like(b_h)
    dz1 = eng.backward(z, x, packed, stash, du, dw1, db1)
    dz2 = eng.backward(z, x, packed, stash, du, dw2, db2)
    assert torch.equal(dw1, dw2) and torch.equal(db1, db2) and torch.equal(dz1, dz2)
    dz3 = eng.backward(z, x, packed, stash, 2 * du, dw2, db2)
    assert rel_err(dw2.cpu(), 2 * dw1.cpu()) < 1e-6 and rel_err(dz3.cpu(), 2 * dz1.cpu()) < 1e-6
    # sum of two half batches == whole batch (what data parallel relies on)
    h = B // 2
    _, st_a = eng.forward(z[:h], x[:h], packed, save=True)
    _, st_b = eng.forward(z[h:], x[h:], packed, save=True)
    dwa, dba = torch.empty_like(w_h), torch.empty_like(b_h)
    eng.backward(z[:h].contiguous(), x[:h].contiguous(), packed, st_a, du[:h].contiguous(), dwa, dba)
    eng.backward(z[h:].contiguous(), x[h:].contiguous(), packed, st_b, du[h:].contiguous(), dwa, dba, beta=1.0)
    assert rel_err(dwa.cpu(), dw1.cpu()) < 1e-5 and rel_err(dba.cpu(), db1.cpu()) < 1e-5


@pytest.mark.parametrize("case", ["siren_2x30", "nif_tanh_si3_so2", "siren_res_so3_n12_K3_sine"])
def test_jacobian_layer_matches_reference_golden(case):
    """JacobianLayer via forward tangents vs the reference's compute_output_and_grad (golden `jac`)."""
    import nif_b200
    from nif_b200.layers import JacobianLayer
    d, cls, cfg_s, cfg_p, spec, prm, _ = load_golden(case)
    net = getattr(nif_b200, cls)(cfg_s, cfg_p, seed=0, device="cuda:0")
    net.set_weights({k: v.float().numpy() for k, v in prm.items()})
    yi, xi = list(d["jac_y_index"]), list(d["jac_x_index"])
    y, J = JacobianLayer(net.build(), yi, xi)(d["inputs"].astype(np.float32))
    assert tuple(J.shape) == d["jac"].shape
    prm32 = {k: v.float() for k, v in prm.items()}
    _, J32 = O.jacobian(spec, prm32, torch.as_tensor(d["inputs"]).float(), yi, xi)
    assert _gate(rel_err(y.cpu(), d["y"]), rel_err(d["y32"], d["y"]))
    assert _gate(rel_err(J.cpu(), d["jac"]), rel_err(J32, d["jac"]), floor=2e-5)


@pytest.mark.parametrize("variant,si,so,n,l,K,B,dirs", [("siren", 1, 1, 64, 4, 32, 300, [0, 1]),   # C4: d/dt, d/dx
                                                         ("siren", 3, 3, 128, 1, 4, 100, [0, 1, 2, 3]),
                                                         ("siren_res", 2, 1, 64, 1, 3, 70, [2]),
                                                         ("nif", 2, 2, 30, 2, 2, 150, [1, 0, 2])])
def test_tangents_random(variant, si, so, n, l, K, B, dirs):
    spec, prm, inputs, _, _ = _random_problem(variant, si, so, n, l, K, B, seed=K + n)
    dev = torch.device("cuda:0")
    wn, bn = O.last_layer_names(spec)
    yi = list(range(so))
    y64, J64 = O.jacobian(spec, prm, inputs, yi, dirs)
    y32, J32 = O.jacobian(spec, {k: v.float() for k, v in prm.items()}, inputs.float(), yi, dirs)
    # latent tangents from the oracle trunk (forward-mode through the trunk is plain torch in the product)
    p_in = inputs[:, :1].clone().requires_grad_(True)
    z = O.latent(spec, prm, p_in)
    zdot = torch.zeros(len(dirs), B, K, dtype=torch.float64)
    xdot = torch.zeros(len(dirs), B, si, dtype=torch.float64)
    for d, c in enumerate(dirs):
        if c == 0:
            for k in range(K):
                (g,) = torch.autograd.grad(z[:, k].sum(), p_in, retain_graph=True)
                zdot[d, :, k] = g[:, 0]
        else:
            xdot[d, :, c - 1] = 1.0
    eng = _engine(spec)
    packed = eng.pack(prm[wn].float().to(dev), prm[bn].float().to(dev))
    u, udot = eng.forward_tangent(z.detach().float().to(dev), inputs[:, 1:].float().contiguous().to(dev), packed,
                                  zdot.float().to(dev), xdot.float().to(dev))
    J = udot.permute(1, 2, 0)
    assert _gate(rel_err(u.cpu(), y64.detach()), rel_err(y32.detach(), y64.detach()))
    assert _gate(rel_err(J.cpu(), J64), rel_err(J32, J64), floor=2e-5)


@pytest.mark.parametrize("variant,si,so,n,l,K,B,xc,act", [
    ("siren", 1, 1, 64, 4, 32, 300, 0, None),      # C4 shape: d/dx of the 1-D travelling wave
    ("siren", 2, 2, 30, 2, 3, 257, 1, None),       # padded width, two outputs, second coordinate
    ("siren_res", 2, 1, 64, 1, 3, 70, 0, None),    # res-blocks
    ("nif", 2, 2, 30, 2, 2, 150, 1, "swish"),      # swish + residual
    ("nif", 3, 1, 20, 1, 4, 65, 2, "tanh"),
    ("siren", 3, 3, 128, 1, 4, 100, 1, None),      # width 128
])
def test_sobolev_reverse_over_forward(variant, si, so, n, l, K, B, xc, act):
    """JacobianLayer inside the loss (tutorial 8): gradients of  mse(u) + coef * mse(du/dx_c)  from the
    reverse-over-forward kernels vs autograd-of-autograd over the oracle."""
    spec, prm, inputs, target, _ = _random_problem(variant, si, so, n, l, K, B, seed=7 * K + n, act=act or "swish")
    dev = torch.device("cuda:0")
    wn, bn = O.last_layer_names(spec)
    g = torch.Generator().manual_seed(5)
    tgt_g = torch.randn(B, so, generator=g, dtype=torch.float64)
    coef = 0.37
    l64, g64, gz64, y64, dy64 = O.sobolev_loss_and_grads(spec, prm, inputs, target, tgt_g, spec.pi + xc, coef)
    prm32 = {k: v.float() for k, v in prm.items()}
    l32, g32, gz32, y32, dy32 = O.sobolev_loss_and_grads(spec, prm32, inputs.float(), target.float(), tgt_g.float(),
                                                         spec.pi + xc, coef)
    eng = _engine(spec)
    z = O.latent(spec, prm, inputs[:, : spec.pi]).float().to(dev)
    x = inputs[:, spec.pi:].float().contiguous().to(dev)
    w_h, b_h = prm[wn].float().to(dev), prm[bn].float().to(dev)
    packed = eng.pack(w_h, b_h)
    xdot = torch.zeros(1, B, si, device=dev)
    xdot[0, :, xc] = 1.0
    u, udot, stash = eng.forward_tangent(z, x, packed, None, xdot, save=True)
    u_ref, udot_ref = eng.forward_tangent(z, x, packed, None, xdot)
    assert torch.equal(u, u_ref) and torch.equal(udot, udot_ref), "stash on/off must not change the outputs"
    assert _gate(rel_err(u.cpu(), y64), rel_err(y32, y64))
    assert _gate(rel_err(udot[0].cpu(), dy64), rel_err(dy32, dy64), floor=2e-5)
    du = (2.0 / (B * so)) * (u - target.float().to(dev))
    dud = (2.0 * coef / (B * so)) * (udot[0] - tgt_g.float().to(dev))
    dw = torch.full_like(w_h, float("nan"))
    db = torch.full_like(b_h, float("nan"))
    dz = eng.sobolev_backward(z, x, xdot[0], packed, stash, du, dud, dw, db, 0.0)
    torch.cuda.synchronize()
    for name, got, r64, r32 in (("dz", dz, gz64, gz32), ("dw_h", dw, g64[wn], g32[wn]), ("db_h", db, g64[bn], g32[bn])):
        e = rel_err(got.cpu(), r64)
        assert _gate(e, rel_err(r32, r64), floor=2e-5), f"{name} err {e:.3e} (cpu32 {rel_err(r32, r64):.3e})"
    # accumulate semantics
    eng.sobolev_backward(z, x, xdot[0], packed, stash, du, dud, dw, db, 1.0)
    assert rel_err(dw.cpu(), 2 * g64[wn]) < 1e-4 and rel_err(db.cpu(), 2 * g64[bn]) < 1e-4


@pytest.mark.parametrize("variant,si,so,n,l,K,B,pairs,act", [
    ("siren", 1, 1, 64, 4, 32, 300, [(0, 0)], None),                  # C4 shape, du/dt alone in the loss
    ("siren", 1, 1, 64, 4, 32, 300, [(0, 0), (0, 1)], None),          # du/dt and du/dx (tutorial 8's full Jacobian)
    ("siren", 2, 2, 30, 2, 3, 257, [(1, 0), (0, 2), (1, 1)], None),   # padded width, two outputs, three directions
    ("siren_res", 2, 1, 64, 1, 3, 70, [(0, 0), (0, 1)], None),        # res-blocks
    ("nif", 2, 2, 30, 2, 2, 150, [(0, 0), (1, 0), (1, 2)], "swish"),  # swish + residual, two outputs along d/dt
    ("siren", 3, 3, 128, 1, 4, 100, [(2, 0)], None),                  # width 128
])
def test_sobolev_with_parameter_net_directions(variant, si, so, n, l, K, B, pairs, act):
    """Jacobian entries w.r.t. ParameterNet inputs inside the loss: the direction moves the latent code (zdot = trunk
    tangent), pre_m' gains sum_k zdot_k (w h_m M_k + C_k), and the reverse-over-forward pass returns dL/dzdot as well.
    Gradients of every variable -- the trunk's through (dz, dzdot) and reverse-over-forward autograd of the oracle's own
    trunk -- vs autograd-of-autograd over the oracle."""
    spec, prm, inputs, target, _ = _random_problem(variant, si, so, n, l, K, B, seed=11 * K + n, act=act or "swish")
    assert spec.pi == 1
    dev = torch.device("cuda:0")
    wn, bn = O.last_layer_names(spec)
    g = torch.Generator().manual_seed(6)
    tgt_g = torch.randn(B, len(pairs), generator=g, dtype=torch.float64)
    coef = 0.37
    l64, g64, y64, dy64 = O.sobolev_loss_and_grads_pairs(spec, prm, inputs, target, tgt_g, pairs, coef)
    prm32 = {k: v.float() for k, v in prm.items()}
    l32, g32, y32, dy32 = O.sobolev_loss_and_grads_pairs(spec, prm32, inputs.float(), target.float(), tgt_g.float(),
                                                         pairs, coef)
    eng = _engine(spec)
    cols = []
    for _, c in pairs:
        if c not in cols:
            cols.append(c)
    D = len(cols)
    # latent code and its tangent along the ParameterNet input, with a tape for the way back (float64, CPU)
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in prm.items()}
    p_in = inputs[:, : spec.pi].detach()
    z64, zd64 = torch.func.jvp(lambda p: O.latent(spec, leaves, p), (p_in,), (torch.ones_like(p_in),))
    z = z64.detach().float().to(dev)
    x = inputs[:, spec.pi:].float().contiguous().to(dev)
    zdot = torch.zeros(D, B, K, device=dev)
    xdot = torch.zeros(D, B, si, device=dev)
    for d, c in enumerate(cols):
        if c < spec.pi:
            zdot[d] = zd64.detach().float().to(dev)
        else:
            xdot[d, :, c - spec.pi] = 1.0
    w_h, b_h = prm[wn].float().to(dev), prm[bn].float().to(dev)
    packed = eng.pack(w_h, b_h)
    u, udot, stash = eng.forward_tangent(z, x, packed, zdot, xdot, save=True)
    u_ref, udot_ref = eng.forward_tangent(z, x, packed, zdot, xdot)
    assert torch.equal(u, u_ref) and torch.equal(udot, udot_ref), "stash on/off must not change the outputs"
    got_dy = torch.stack([udot[cols.index(c)][:, a] for a, c in pairs], 1)
    assert _gate(rel_err(u.cpu(), y64), rel_err(y32, y64))
    assert _gate(rel_err(got_dy.cpu(), dy64), rel_err(dy32, dy64), floor=2e-5)
    du = (2.0 / (B * so)) * (u - target.float().to(dev))
    dud = torch.zeros_like(udot)
    for i, (a, c) in enumerate(pairs):
        dud[cols.index(c)][:, a] += (2.0 * coef / (B * len(pairs))) * (got_dy[:, i] - tgt_g[:, i].float().to(dev))
    dw = torch.full_like(w_h, float("nan"))
    db = torch.full_like(b_h, float("nan"))
    tdir = [d for d, c in enumerate(cols) if c < spec.pi]
    dz, dzdot = eng.sobolev_backward(z, x, xdot, packed, stash, du, dud, dw, db, 0.0, zdot=zdot, zdot_dirs=tdir)
    torch.cuda.synchronize()
    assert all(float(dzdot[d].abs().max()) == 0.0 for d in range(D) if d not in tdir)  # skipped directions: zero
    torch.autograd.backward([z64, zd64], [dz.cpu().double(), sum(dzdot[d] for d in tdir).cpu().double()])
    for name in prm:
        got = dw.cpu() if name == wn else db.cpu() if name == bn else leaves[name].grad
        if got is None:
            got = torch.zeros_like(prm[name])
        e = rel_err(got, g64[name])
        assert _gate(e, rel_err(g32[name], g64[name]), floor=2e-5), f"{name} err {e:.3e} (cpu32 {rel_err(g32[name], g64[name]):.3e})"


def test_sobolev_training_with_dudt_in_the_loss():
    """Tutorial 8's model with BOTH Jacobian columns in the loss (du/dt differentiates through the ParameterNet): three
    Adam steps against autograd-of-autograd over the oracle."""
    import nif_b200
    cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2,
             "weight_init_factor": 0.01, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 2, "units": 30, "nlayers": 2, "activation": "swish"}
    spec = O.spec_from_cfg("NIFMultiScale", cfg_s, cfg_p)
    prm0 = O.init_params(spec, 4)
    net = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=0, device="cuda:0")
    net.set_weights({k: v.numpy() for k, v in prm0.items()})
    coef = 1e-2
    model = nif_b200.JacobianLayer(net.build(), y_index=[0], x_index=[0, 1]).as_model()
    model.compile(nif_b200.Adam(1e-3), loss=nif_b200.SobolevMSE(coef, value_cols=[0], grad_cols=[1, 2]))
    prm = {k: v.double().clone() for k, v in prm0.items()}
    m_ = {k: torch.zeros_like(v) for k, v in prm.items()}
    v_ = {k: torch.zeros_like(v) for k, v in prm.items()}
    rng = np.random.default_rng(10)
    B = 300
    for step in range(1, 4):
        X = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
        Y3 = rng.uniform(-1, 1, (B, 3)).astype(np.float32)
        Y3[:, 1:] *= 5.0
        l_gpu = model.train_on_batch(X, Y3)
        Xd, Yd = torch.as_tensor(X).double(), torch.as_tensor(Y3).double()
        l_ref, g, _, _ = O.sobolev_loss_and_grads_pairs(spec, prm, Xd, Yd[:, :1], Yd[:, 1:3], [(0, 0), (0, 1)], coef)
        for k in prm:
            O.adam_tf(prm[k], g[k], m_[k], v_[k], step, 1e-3)
        assert abs(l_gpu - float(l_ref)) <= 1e-4 * max(1.0, abs(float(l_ref))), (step, l_gpu, float(l_ref))
    got = net.get_weights()
    diffs = np.concatenate([np.abs(got[k] - v.numpy()).ravel() for k, v in prm.items()])
    assert float(np.quantile(diffs, 0.999)) < 1e-4, float(np.quantile(diffs, 0.999))
    assert float(diffs.max()) < 4 * 2e-3


def test_pde_residual_loss_on_jacobian_model():
    """A callable loss over the whole output of a JacobianLayer model -- here the advection residual u_t + c u_x next to
    a data term, the physics-informed usage the reference's README points at -- and plain 'mse' over [u, u_t, u_x]: three
    Adam steps each against autograd-of-autograd over the oracle."""
    import nif_b200
    cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2,
             "weight_init_factor": 0.01, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 2, "units": 30, "nlayers": 2, "activation": "swish"}
    spec = O.spec_from_cfg("NIFMultiScale", cfg_s, cfg_p)

    def residual(y_true, y_pred):
        r = y_pred[:, 1] + 0.7 * y_pred[:, 2]
        return ((y_true[:, 0] - y_pred[:, 0]) ** 2).mean() + 1e-2 * (r * r).mean()

    def mse(y_true, y_pred):
        return ((y_true - y_pred) ** 2).mean(-1).mean()

    for loss_gpu, loss_ref in ((residual, residual), ("mse", mse)):
        prm0 = O.init_params(spec, 5)
        net = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=0, device="cuda:0")
        net.set_weights({k: v.numpy() for k, v in prm0.items()})
        model = nif_b200.JacobianLayer(net.build(), y_index=[0], x_index=[0, 1]).as_model()
        model.compile(nif_b200.Adam(1e-3), loss=loss_gpu)
        prm = {k: v.double().clone() for k, v in prm0.items()}
        m_ = {k: torch.zeros_like(v) for k, v in prm.items()}
        v_ = {k: torch.zeros_like(v) for k, v in prm.items()}
        rng = np.random.default_rng(12)
        B = 257
        for step in range(1, 4):
            X = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
            Y3 = rng.uniform(-1, 1, (B, 3)).astype(np.float32)
            l_gpu = model.train_on_batch(X, Y3)
            l_ref, g, _ = O.jacobian_model_loss_and_grads(spec, prm, torch.as_tensor(X).double(), torch.as_tensor(Y3).double(),
                                                          loss_ref, [0], [0, 1])
            for k in prm:
                O.adam_tf(prm[k], g[k], m_[k], v_[k], step, 1e-3)
            assert abs(l_gpu - float(l_ref)) <= 1e-4 * max(1.0, abs(float(l_ref))), (loss_gpu, step, l_gpu, float(l_ref))
        got = net.get_weights()
        diffs = np.concatenate([np.abs(got[k] - v.numpy()).ravel() for k, v in prm.items()])
        assert float(np.quantile(diffs, 0.999)) < 1e-4, (loss_gpu, float(np.quantile(diffs, 0.999)))
        assert float(diffs.max()) < 4 * 2e-3


# --------------------------------------------------------------------------------------------------
# tensor-core path (tcgen05, FP16x3): same gates as the fp32 CUDA-core path
# --------------------------------------------------------------------------------------------------
def _engine_tc(spec):
    from nif_b200.ops import FusedShapeNet
    return FusedShapeNet(spec.variant, spec.si, spec.so, spec.n, spec.l, spec.K, spec.s_act, spec.omega0,
                         compute="fp16x3")


@pytest.mark.parametrize("si,so,n,l,K,B,pairs", [
    (1, 1, 64, 4, 32, 300, [(0, 1)]),                    # C4 shape: du/dx of the 1-D travelling wave
    (2, 2, 48, 2, 5, 1000, [(0, 1), (1, 2), (1, 1)]),    # two directions, two outputs, padded width, several tiles
    (3, 1, 64, 1, 3, 77, [(0, 3), (0, 1)]),              # one hidden matrix, ragged tile, directions out of order
    (2, 1, 64, 2, 8, 128, [(0, 2)]),                     # exactly one tile
])
def test_sobolev_on_the_tensor_cores(si, so, n, l, K, B, pairs):
    """Sobolev step of a SIREN ShapeNet with ShapeNet-input directions on the tcgen05 kernels: forward tangents as a
    mode of the forward kernel (nif_tc_fwd_kernel<tangent>), both adjoint passes through nif_tc_bwd_data_kernel<ext> and
    the FP16x3 batch reductions.  Every variable's gradient vs autograd-of-autograd over the fp64 oracle, and the
    kernel table must name the tensor-core kernels (no fall-through to the CUDA-core pass)."""
    from nif_b200.ops import kernel_profile
    spec, prm, inputs, target, _ = _random_problem("siren", si, so, n, l, K, B, seed=13 * K + n)
    dev = torch.device("cuda:0")
    wn, bn = O.last_layer_names(spec)
    g = torch.Generator().manual_seed(8)
    tgt_g = torch.randn(B, len(pairs), generator=g, dtype=torch.float64)
    coef = 0.37
    l64, g64, y64, dy64 = O.sobolev_loss_and_grads_pairs(spec, prm, inputs, target, tgt_g, pairs, coef)
    prm32 = {k: v.float() for k, v in prm.items()}
    l32, g32, y32, dy32 = O.sobolev_loss_and_grads_pairs(spec, prm32, inputs.float(), target.float(), tgt_g.float(),
                                                         pairs, coef)
    eng = _engine_tc(spec)
    assert eng.kernel_path == "fp16x3"
    cols = []
    for _, c in pairs:
        assert c >= spec.pi
        if c not in cols:
            cols.append(c)
    D = len(cols)
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in prm.items()}
    z64 = O.latent(spec, leaves, inputs[:, : spec.pi].detach())
    z = z64.detach().float().to(dev)
    x = inputs[:, spec.pi:].float().contiguous().to(dev)
    xdot = torch.zeros(D, B, si, device=dev)
    for d, c in enumerate(cols):
        xdot[d, :, c - spec.pi] = 1.0
    w_h, b_h = prm[wn].float().to(dev), prm[bn].float().to(dev)
    packed = eng.pack(w_h, b_h)
    dw = torch.full_like(w_h, float("nan"))
    db = torch.full_like(b_h, float("nan"))
    with kernel_profile() as prof:
        u, udot, stash = eng.forward_tangent(z, x, packed, None, xdot, save=True)
        got_dy = torch.stack([udot[cols.index(c)][:, a] for a, c in pairs], 1)
        du = (2.0 / (B * so)) * (u - target.float().to(dev))
        dud = torch.zeros_like(udot)
        for i, (a, c) in enumerate(pairs):
            dud[cols.index(c)][:, a] += (2.0 * coef / (B * len(pairs))) * (got_dy[:, i] - tgt_g[:, i].float().to(dev))
        dz = eng.sobolev_backward(z, x, xdot, packed, stash, du, dud, dw, db, 0.0)
        torch.cuda.synchronize()
    names = {k: c for k, c, _ in prof.table}
    assert names.get("nif_tc_fwd_kernel<tangent>") == D and names.get("nif_tc_bwd_data_kernel<ext>") == D + 1, names
    assert names.get("nif_tc_bwd_weight_kernel") == D + 1 and not any(k.startswith("nif_bwd_") or k.startswith("nif_tangent") for k in names), names
    # the CUDA-core tangent kernel (an explicit zdot selects it) is an independent implementation of the same outputs
    with kernel_profile() as prof2:
        u_ref, udot_ref = eng.forward_tangent(z, x, packed, torch.zeros(D, B, K, device=dev), xdot)
        u_tc, udot_tc = eng.forward_tangent(z, x, packed, None, xdot)  # no stash asked for: still the tensor-core pair
    names2 = {k: c for k, c, _ in prof2.table}
    assert names2.get("nif_tc_fwd_kernel<tangent>") == D and sum(c for k, c in names2.items() if k.startswith("nif_tangent")) >= 1, names2
    assert torch.equal(u_tc, u) and torch.equal(udot_tc, udot)
    assert rel_err(u.cpu(), u_ref.cpu()) < 1e-5 and rel_err(udot.cpu(), udot_ref.cpu()) < 2e-5
    assert _gate(rel_err(u.cpu(), y64), rel_err(y32, y64))
    assert _gate(rel_err(got_dy.cpu(), dy64), rel_err(dy32, dy64), floor=2e-5)
    z64.backward(dz.cpu().double())
    for name in prm:
        got = dw.cpu() if name == wn else db.cpu() if name == bn else leaves[name].grad
        if got is None:
            got = torch.zeros_like(prm[name])
        e = rel_err(got, g64[name])
        assert _gate(e, rel_err(g32[name], g64[name]), floor=2e-5), f"{name} err {e:.3e} (cpu32 {rel_err(g32[name], g64[name]):.3e})"
    # accumulate semantics
    eng.sobolev_backward(z, x, xdot, packed, stash, du, dud, dw, db, 1.0)
    assert rel_err(dw.cpu(), 2 * g64[wn]) < 1e-4 and rel_err(db.cpu(), 2 * g64[bn]) < 1e-4


@pytest.mark.parametrize(
    "variant,si,so,n,l,K,B",
    [
        ("siren", 2, 1, 64, 4, 32, 300),     # C2 shape, ragged batch, odd number of latent rows (K+1 = 33)
        ("siren", 1, 1, 64, 4, 32, 128),     # C4 shape, exactly one tile
        ("siren", 2, 1, 64, 1, 3, 1000),     # even K+1, several tiles
        ("siren", 3, 2, 48, 2, 5, 77),       # padded width
        ("siren", 2, 2, 64, 2, 6, 150),      # two outputs share one last-layer chunk
        ("nif", 2, 2, 48, 3, 7, 90),         # swish + residual
        ("siren", 2, 1, 64, 2, 1, 20000),    # many tiles per CTA (persistent loop, barrier phases)
    ],
)
def test_tc_forward_and_stash(variant, si, so, n, l, K, B):
    spec, prm, inputs, target, sw = _random_problem(variant, si, so, n, l, K, B, seed=si * 100 + n + K)
    dev = torch.device("cuda:0")
    wn, bn = O.last_layer_names(spec)
    loss64, g64, gz64, y64 = O.loss_and_grads(spec, prm, inputs, target, sw)
    prm32 = {k: v.float() for k, v in prm.items()}
    loss32, g32, gz32, y32 = O.loss_and_grads(spec, prm32, inputs.float(), target.float(), sw.float())
    eng = _engine_tc(spec)
    assert eng.kernel_path == "fp16x3", "this descriptor must be served by the tcgen05 kernels, not a fall-through"
    z = O.latent(spec, prm, inputs[:, :1]).float().to(dev)
    x = inputs[:, 1:].float().contiguous().to(dev)
    w_h, b_h = prm[wn].float().to(dev), prm[bn].float().to(dev)
    packed = eng.pack(w_h, b_h)
    u = eng.forward(z, x, packed)
    u2, stash = eng.forward(z, x, packed, save=True)
    torch.cuda.synchronize()
    assert torch.equal(u, u2)
    assert _gate(rel_err(u.cpu(), y64), rel_err(y32, y64)), (rel_err(u.cpu(), y64), rel_err(y32, y64))
    # the stash written by the tensor-core forward drives the reverse pass
    loss = torch.zeros(1, device=dev)
    dw, db = torch.empty_like(w_h), torch.empty_like(b_h)
    dz = eng.mse_backward(z, x, packed, u, stash, target.float().to(dev), sw.float().to(dev), 1.0 / B, loss, dw, db)
    for name, got, r64, r32 in (("dz", dz, gz64, gz32), ("dw_h", dw, g64[wn], g32[wn]), ("db_h", db, g64[bn], g32[bn])):
        e = rel_err(got.cpu(), r64)
        assert _gate(e, rel_err(r32, r64)), f"{name} err {e:.3e} (cpu32 {rel_err(r32, r64):.3e})"


# --------------------------------------------------------------------------------------------------
# fused ParameterNet trunk and the whole optimisation step
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pi,K,n_st,l_st,act,B", [(1, 32, 64, 4, "swish", 700), (1, 1, 30, 2, "swish", 300),
                                                  (2, 5, 17, 1, "tanh", 129), (3, 64, 64, 0, "relu", 50),
                                                  (8, 64, 64, 4, "sigmoid", 1000),      # widest shapes of the tensor-core trunk
                                                  (1, 32, 64, 4, "swish", 65536),       # the benchmark batch
                                                  (1, 16, 48, 2, "swish", 700000)])     # > 32 tiles per CTA: accumulators flushed
def test_fused_trunk_matches_oracle(pi, K, n_st, l_st, act, B):
    from nif_b200.ops import FusedTrunk
    dev = torch.device("cuda:0")
    spec = O.Spec(variant="siren", pi=pi, si=1, so=1, n=8, l=1, K=K, n_st=n_st, l_st=l_st, p_act=act, omega0=30.0,
                  weight_init_factor=0.1)
    prm = {k: v.double() for k, v in O.init_params(spec, 7).items()}
    rng = np.random.default_rng(3)
    p_in = torch.as_tensor(rng.uniform(-1, 1, (B, pi)))
    dz = torch.as_tensor(rng.normal(size=(B, K)))
    names = [k for k in O.trunk_param_names(spec)[:-2]]
    order = [k for k in names if k.endswith("/kernel")] + [k for k in names if k.endswith("/bias")]
    leaves = {k: prm[k].clone().requires_grad_(True) for k in names}
    z64 = O.latent(spec, {**prm, **leaves}, p_in)
    (z64 * dz).sum().backward()
    z32 = O.latent(spec, {k: v.float() for k, v in prm.items()}, p_in.float())
    g64 = torch.cat([leaves[k].grad.reshape(-1) for k in order])
    theta = torch.cat([prm[k].reshape(-1) for k in order]).float().to(dev)
    tr = FusedTrunk(pi, K, n_st, l_st, act)
    assert tr.n_theta == theta.numel()
    assert tr.kernel_path == ("bf16x3" if 1 <= l_st <= 4 else "fp32")  # tcgen05 trunk kernels wherever they are built
    z, stash = tr.forward(p_in.float().to(dev), theta, save=True)
    assert torch.equal(z, tr.forward(p_in.float().to(dev), theta))
    assert _gate(rel_err(z.cpu(), z64.detach()), rel_err(z32, z64.detach()))
    g = torch.full_like(theta, float("nan"))
    tr.backward(p_in.float().to(dev), theta, stash, dz.float().to(dev), g, 0.0)
    assert rel_err(g.cpu(), g64) < 2e-5
    tr.backward(p_in.float().to(dev), theta, stash, dz.float().to(dev), g, 1.0)
    assert rel_err(g.cpu(), 2 * g64) < 2e-5


@pytest.mark.parametrize("cls,cfg_s,cfg_p", [
    ("NIFMultiScale",
     {"use_resblock": False, "connectivity": "full", "input_dim": 2, "output_dim": 1, "units": 64, "nlayers": 4,
      "weight_init_factor": 0.01, "omega_0": 30.0},
     {"use_resblock": False, "input_dim": 1, "latent_dim": 32, "units": 64, "nlayers": 4, "activation": "swish"}),
    ("NIF",
     {"connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"},
     {"input_dim": 1, "latent_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}),
    ("NIFMultiScale",
     {"use_resblock": True, "connectivity": "full", "input_dim": 2, "output_dim": 2, "units": 24, "nlayers": 1,
      "weight_init_factor": 0.1, "omega_0": 10.0},
     {"use_resblock": True, "input_dim": 1, "latent_dim": 3, "units": 10, "nlayers": 2, "activation": "swish"}),
])
def test_training_steps_match_oracle_trainer(cls, cfg_s, cfg_p):
    """Model.fit-style steps (trunk + head + ShapeNet + 'mse' + reverse pass + Adam) against the materialised
    fp64 restatement of the reference's dataflow with TF-semantics Adam."""
    import nif_b200
    spec = O.spec_from_cfg(cls, cfg_s, cfg_p)
    prm = O.init_params(spec, 11)
    net = getattr(nif_b200, cls)(cfg_s, cfg_p, seed=0, device="cuda:0")
    net.set_weights({k: v.numpy() for k, v in prm.items()})
    model = net.build()
    model.compile(nif_b200.Adam(1e-3), loss="mse")
    ref = O.MaterialisedTrainer(spec, {k: v.double() for k, v in prm.items()}, lr=1e-3)
    rng = np.random.default_rng(5)
    B = 400
    for step in range(4):
        X = rng.uniform(-1, 1, (B, spec.pi + spec.si)).astype(np.float32)
        Y = rng.uniform(-1, 1, (B, spec.so)).astype(np.float32)
        sw = rng.uniform(0.5, 1.5, (B,)).astype(np.float32) if step % 2 else None
        l_gpu = model.train_on_batch(X, Y, sample_weight=sw)
        l_ref = ref.step(torch.as_tensor(X).double(), torch.as_tensor(Y).double(),
                         None if sw is None else torch.as_tensor(sw).double())
        assert abs(l_gpu - l_ref) <= 1e-4 * max(1.0, abs(l_ref)), (step, l_gpu, l_ref)
    got = net.get_weights()
    # Adam's first steps move every weight by ~lr whatever the gradient scale, and a weight whose gradient is at the
    # rounding level may step the other way: compare absolutely, and allow a vanishing fraction of such weights.
    diffs = np.concatenate([np.abs(got[k] - v.detach().numpy()).ravel() for k, v in ref.prm.items()])
    assert float(np.quantile(diffs, 0.999)) < 1e-4, float(np.quantile(diffs, 0.999))
    assert float(diffs.max()) < 4 * 2e-3


def test_sobolev_training_matches_oracle_tutorial8():
    """Tutorial 8 end to end: JacobianLayer(model, [0], [0, 1]) -> concat [u, du/dt, du/dx] -> Sobolov_MSE (u and du/dx)
    -> Adam, against autograd-of-autograd over the oracle with TF-semantics Adam."""
    import nif_b200
    cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2,
             "weight_init_factor": 0.01, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
    spec = O.spec_from_cfg("NIFMultiScale", cfg_s, cfg_p)
    prm0 = O.init_params(spec, 3)
    net = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=0, device="cuda:0")
    net.set_weights({k: v.numpy() for k, v in prm0.items()})
    coef = 1e-3
    model = nif_b200.JacobianLayer(net.build(), y_index=[0], x_index=[0, 1]).as_model()
    model.compile(nif_b200.Adam(1e-3), loss=nif_b200.SobolevMSE(coef, value_cols=[0], grad_cols=[2]))
    prm = {k: v.double().clone() for k, v in prm0.items()}
    m_ = {k: torch.zeros_like(v) for k, v in prm.items()}
    v_ = {k: torch.zeros_like(v) for k, v in prm.items()}
    rng = np.random.default_rng(9)
    B = 300
    for step in range(1, 4):
        X = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
        Y3 = rng.uniform(-1, 1, (B, 3)).astype(np.float32)
        Y3[:, 2] *= 20.0
        if step == 1:  # predict() of the wrapped model = [u, du/dt, du/dx] of the oracle
            y64, J64 = O.jacobian(spec, prm, torch.as_tensor(X).double(), [0], [0, 1])
            y32, J32 = O.jacobian(spec, {k: v.float() for k, v in prm.items()}, torch.as_tensor(X), [0], [0, 1])
            ref3 = torch.cat([y64.detach(), J64.reshape(B, 2)], -1)
            ref3_32 = torch.cat([y32.detach(), J32.reshape(B, 2)], -1).double()
            got3 = model.predict(X)
            assert got3.shape == (B, 3)
            assert _gate(rel_err(torch.as_tensor(got3), ref3), rel_err(ref3_32, ref3), floor=2e-5)
        l_gpu = model.train_on_batch(X, Y3)
        Xd, Yd = torch.as_tensor(X).double(), torch.as_tensor(Y3).double()
        l_ref, g, _, _, _ = O.sobolev_loss_and_grads(spec, prm, Xd, Yd[:, :1], Yd[:, 2:3], 1, coef)
        for k in prm:
            O.adam_tf(prm[k], g[k], m_[k], v_[k], step, 1e-3)
        assert abs(l_gpu - float(l_ref)) <= 1e-4 * max(1.0, abs(float(l_ref))), (step, l_gpu, float(l_ref))
    got = net.get_weights()
    diffs = np.concatenate([np.abs(got[k] - v.numpy()).ravel() for k, v in prm.items()])
    assert float(np.quantile(diffs, 0.999)) < 1e-4, float(np.quantile(diffs, 0.999))
    assert float(diffs.max()) < 4 * 2e-3


def test_graph_replayed_steps_equal_eager_steps_bitwise():
    """The CUDA-graph replay of the fused step (Model._train_step_graph) launches the same kernels on the same
    buffers as the eager step: after a run that mixes two batch sizes, sample weights and a learning-rate change the
    parameters, the Adam moments and every reported loss are bit-identical."""
    import nif_b200
    cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 2, "output_dim": 1, "units": 64, "nlayers": 4,
             "weight_init_factor": 0.01, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 32, "units": 64, "nlayers": 4, "activation": "swish"}
    rng = np.random.default_rng(21)
    batches = []
    for step in range(9):
        B = 700 if step % 3 != 2 else 333
        X = rng.uniform(-1, 1, (B, 3)).astype(np.float32)
        Y = rng.uniform(-1, 1, (B, 1)).astype(np.float32)
        sw = rng.uniform(0.5, 1.5, (B,)).astype(np.float32) if step >= 6 and B == 700 else None
        batches.append((X, Y, sw))
    runs = {}
    for graph in (False, True):
        net = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=4, device="cuda:0")
        model = net.build()
        opt = nif_b200.Adam(1e-3)
        model.compile(opt, loss="mse", graph=graph)
        losses = []
        for step, (X, Y, sw) in enumerate(batches):
            if step == 5:
                opt.learning_rate = 2.5e-4  # what LearningRateScheduler does at an epoch boundary
            losses.append(model.train_on_batch(X, Y, sample_weight=sw))
        runs[graph] = (net.theta.cpu().clone(), opt._m.cpu().clone(), opt._v.cpu().clone(), losses, opt.iterations)
        if graph:
            assert sum(1 for e in model._graphs.values() if e.get("graph") is not None) >= 2  # both sizes were recorded
    (t0, m0, v0, l0, i0), (t1, m1, v1, l1, i1) = runs[False], runs[True]
    assert i0 == i1 == len(batches)
    assert l0 == l1
    assert torch.equal(t0, t1) and torch.equal(m0, m1) and torch.equal(v0, v1)


@pytest.mark.parametrize("case", ["siren_2x30", "nif_tanh_si3_so2", "siren_res_so3_n12_K3_sine"])
def test_hessian_layer_matches_reference_golden(case):
    """HessianLayer via second-order forward tangents (nif_forward_tangent2) vs the reference's
    compute_output_and_grad_and_hessian (golden `jac`, `hess`, made by running the reference source)."""
    import nif_b200
    d, cls, cfg_s, cfg_p, spec, prm, _ = load_golden(case)
    net = getattr(nif_b200, cls)(cfg_s, cfg_p, seed=0, device="cuda:0")
    net.set_weights({k: v.float().numpy() for k, v in prm.items()})
    yi, xi = list(d["jac_y_index"]), list(d["jac_x_index"])
    y, J, H = nif_b200.HessianLayer(net.build(), yi, xi)(d["inputs"].astype(np.float32))
    assert tuple(J.shape) == d["jac"].shape and tuple(H.shape) == d["hess"].shape
    prm32 = {k: v.float() for k, v in prm.items()}
    _, J32, H32 = O.hessian(spec, prm32, torch.as_tensor(d["inputs"]).float(), yi, xi)
    assert _gate(rel_err(y.cpu(), d["y"]), rel_err(d["y32"], d["y"]))
    assert _gate(rel_err(J.cpu(), d["jac"]), rel_err(J32, d["jac"]), floor=2e-5)
    assert _gate(rel_err(H.cpu(), d["hess"]), rel_err(H32, d["hess"]), floor=5e-5)


@pytest.mark.parametrize("cls,cfg_s,cfg_p,B,xi", [
    ("NIFMultiScale",  # C4 shape: derivatives w.r.t. t (through the trunk, second order) and x
     {"use_resblock": False, "connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 64, "nlayers": 4,
      "weight_init_factor": 0.01, "omega_0": 30.0},
     {"use_resblock": False, "input_dim": 1, "latent_dim": 32, "units": 64, "nlayers": 4, "activation": "swish"}, 200, [0, 1]),
    ("NIFMultiScale",  # spatial Hessian of a 3-D field, width 128
     {"use_resblock": False, "connectivity": "full", "input_dim": 3, "output_dim": 2, "units": 128, "nlayers": 1,
      "weight_init_factor": 0.05, "omega_0": 10.0},
     {"use_resblock": False, "input_dim": 1, "latent_dim": 4, "units": 16, "nlayers": 1, "activation": "swish"}, 90, [1, 2, 3]),
    ("NIF",
     {"connectivity": "full", "input_dim": 2, "output_dim": 2, "units": 30, "nlayers": 2, "activation": "swish"},
     {"input_dim": 1, "latent_dim": 2, "units": 30, "nlayers": 2, "activation": "tanh"}, 150, [2, 0]),
])
def test_hessian_layer_random(cls, cfg_s, cfg_p, B, xi):
    import nif_b200
    spec = O.spec_from_cfg(cls, cfg_s, cfg_p)
    prm = O.init_params(spec, 17)
    net = getattr(nif_b200, cls)(cfg_s, cfg_p, seed=0, device="cuda:0")
    net.set_weights({k: v.numpy() for k, v in prm.items()})
    rng = np.random.default_rng(B)
    X = rng.uniform(-1, 1, (B, spec.pi + spec.si)).astype(np.float32)
    yi = list(range(spec.so))
    y, J, H = nif_b200.HessianLayer(net.build(), yi, xi)(X)
    prm64 = {k: v.double() for k, v in prm.items()}
    y64, J64, H64 = O.hessian(spec, prm64, torch.as_tensor(X).double(), yi, xi)
    y32, J32, H32 = O.hessian(spec, prm, torch.as_tensor(X), yi, xi)
    assert tuple(H.shape) == (B, len(yi), len(xi), len(xi))
    assert torch.equal(H, H.transpose(2, 3))
    assert _gate(rel_err(y.cpu(), y64), rel_err(y32, y64))
    assert _gate(rel_err(J.cpu(), J64), rel_err(J32, J64), floor=2e-5)
    assert _gate(rel_err(H.cpu(), H64), rel_err(H32, H64), floor=5e-5)


def test_training_with_jac_reg_matches_oracle_trainer():
    """cfg_parameter_net['jac_reg'] (JacRegLatentLayer, nif/model.py:353-375): the training step adds
    jac_reg * mean((d latent / d input_p)^2) to the loss; steps against the oracle trainer with the same term."""
    import nif_b200
    cfg_s = {"connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
    cfg_p = {"input_dim": 1, "latent_dim": 2, "units": 30, "nlayers": 2, "activation": "swish", "jac_reg": 30.0}
    spec = O.spec_from_cfg("NIF", cfg_s, cfg_p)
    prm = O.init_params(spec, 23)
    net = nif_b200.NIF(cfg_s, cfg_p, seed=0, device="cuda:0")
    net.set_weights({k: v.numpy() for k, v in prm.items()})
    model = net.build()
    model.compile(nif_b200.Adam(1e-3), loss="mse")
    ref = O.MaterialisedTrainer(spec, {k: v.double() for k, v in prm.items()}, lr=1e-3, jac_reg=30.0)
    ref0 = O.MaterialisedTrainer(spec, {k: v.double() for k, v in prm.items()}, lr=1e-3)
    rng = np.random.default_rng(6)
    B = 256
    for step in range(3):
        X = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
        Y = rng.uniform(-1, 1, (B, 1)).astype(np.float32)
        l_gpu = model.train_on_batch(X, Y)
        l_ref = ref.step(torch.as_tensor(X).double(), torch.as_tensor(Y).double())
        if step == 0:
            l_plain = ref0.step(torch.as_tensor(X).double(), torch.as_tensor(Y).double())
            assert l_ref - l_plain > 1e-2 * l_plain  # the regulariser is a visible part of the loss in this set-up
        assert abs(l_gpu - l_ref) <= 1e-4 * max(1.0, abs(l_ref)), (step, l_gpu, l_ref)
    got = net.get_weights()
    diffs = np.concatenate([np.abs(got[k] - v.detach().numpy()).ravel() for k, v in ref.prm.items()])
    assert float(np.quantile(diffs, 0.999)) < 1e-4, float(np.quantile(diffs, 0.999))
    assert float(diffs.max()) < 4 * 2e-3


def test_fit_matches_oracle_trainer_on_tutorial1_data():
    """Model.fit (tutorial/1_simple_1d_wave.ipynb: NIF swish 2x30, batch 512 over the 2000-point travelling wave, so
    a short last batch; LearningRateScheduler; shuffled tf.data-style Dataset) against the oracle trainer fed the same
    batch sequence: per-epoch loss (Keras' sample-weighted running mean) and final parameters."""
    import nif_b200
    cfg_s = {"connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
    cfg_p = {"input_dim": 1, "latent_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
    spec = O.spec_from_cfg("NIF", cfg_s, cfg_p)
    prm = O.init_params(spec, 2)
    raw = O.traveling_wave_raw(4.0)
    data, _, _ = O.standard_normalize(raw)
    data = data.astype(np.float32)
    X, Y = data[:, :2], data[:, 2:3]
    net = nif_b200.NIF(cfg_s, cfg_p, seed=0, device="cuda:0")
    net.set_weights({k: v.numpy() for k, v in prm.items()})
    model = net.build()
    model.compile(nif_b200.Adam(1e-3), loss="mse")
    sched = lambda epoch, lr: 1e-3 if epoch < 2 else 5e-4
    ds = nif_b200.Dataset.from_tensor_slices((X, Y)).shuffle(2000, seed=5).batch(512).prefetch(1)
    hist = model.fit(ds, epochs=3, callbacks=[nif_b200.LearningRateScheduler(sched)], verbose=0)
    # the same batches through the oracle
    ref = O.MaterialisedTrainer(spec, {k: v.double() for k, v in prm.items()}, lr=1e-3)
    ds2 = nif_b200.Dataset.from_tensor_slices((X, Y)).shuffle(2000, seed=5).batch(512)
    ref_hist = []
    for epoch in range(3):
        ref.lr = sched(epoch, ref.lr)
        tot, rows = 0.0, 0
        sizes = []
        for gb, parts in ds2.batches(epoch):
            l = ref.step(parts[0].double(), parts[1].double())
            tot += l * gb
            rows += gb
            sizes.append(gb)
        assert sizes == [512, 512, 512, 464]
        ref_hist.append(tot / rows)
    assert len(hist.history["loss"]) == 3 and hist.history["lr"] == [1e-3, 1e-3, 5e-4]
    for a, b in zip(hist.history["loss"], ref_hist):
        assert abs(a - b) <= 2e-4 * max(1.0, abs(b)), (hist.history["loss"], ref_hist)
    got = net.get_weights()
    diffs = np.concatenate([np.abs(got[k] - v.detach().numpy()).ravel() for k, v in ref.prm.items()])
    # (Adam moves every weight by ~lr per step whatever the gradient scale; a weight whose gradient sits at the rounding
    # level may step the other way, so compare a high quantile tightly and the maximum loosely)
    assert float(np.quantile(diffs, 0.995)) < 3e-4, float(np.quantile(diffs, 0.995))
    assert float(diffs.max()) < 12 * 2e-3
    assert model.optimizer.iterations == 12
    assert ds.arrays[0].is_cuda  # fit() keeps a data set of this size resident in HBM


def test_full_batch_head_gradient_against_oracle():
    """Parity AT the benchmark batch (65 536 rows), where the tensor-core weight-gradient kernel runs its longest
    accumulation chains: head gradients, dz and the loss against the fp64 oracle summed over 4096-row chunks
    (tests/golden/fullbatch/make_fullbatch_ref.py).

    Measured on B200 (round 1, deterministic kernels): loss 1.4e-7, dz 2.5e-6, dw 4.8e-6, db 5.9e-6 with at most 4096 rows
    per TMEM accumulation chain (the default).  With 16 384 rows per chain (one wave of CTAs, what the kernel did first)
    dw / db were 1.95e-5 / 2.38e-5: the tensor core truncates its fp32 accumulator on every instruction."""
    import os
    from tests.helpers import GOLDEN, fullbatch_problem
    ref = np.load(os.path.join(GOLDEN, "fullbatch", "c2_fullbatch_head_grad.npz"))
    spec, prm, z, x, tgt = fullbatch_problem()
    dev = torch.device("cuda:0")
    wn, bn = O.last_layer_names(spec)
    eng = _engine_tc(spec)
    w_h, b_h = prm[wn].to(dev), prm[bn].to(dev)
    packed = eng.pack(w_h, b_h)
    zd, xd = z.to(dev), x.to(dev)
    u, stash = eng.forward(zd, xd, packed, save=True)
    loss = torch.zeros(1, device=dev)
    dw, db = torch.empty_like(w_h), torch.empty_like(b_h)
    dz = eng.mse_backward(zd, xd, packed, u, stash, tgt.to(dev), None, 1.0 / z.shape[0], loss, dw, db)
    errs = {"dw": rel_err(dw.cpu(), ref["dw"]), "db": rel_err(db.cpu(), ref["db"]),
            "dz": rel_err(dz[:256].cpu(), ref["dz_head"]), "loss": abs(float(loss) - float(ref["loss"])) / float(ref["loss"])}
    print("full-batch errors vs fp64 oracle:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["loss"] < 1e-5 and errs["dz"] < 1e-5 and errs["dw"] < 1e-5 and errs["db"] < 1e-5, errs


@pytest.mark.parametrize("B", [113664, 65536])
def test_full_batch_whole_model_step_against_oracle(B):
    """Parity of the WHOLE model AT the benchmark batch: one train_on_batch on bench.py's configuration at bench.py's batch
    (65 536 rows; and at 113 664 rows = 148 SMs x 3 pairs of 128-row tiles, the wave-aligned batch of `--batch 113664`), then every
    variable's gradient (trunk kernels and biases included) as it sits in the flat gradient buffer BEFORE Adam reads it,
    against the fp64 oracle summed over 4096-row chunks (tests/golden/fullbatch/make_fullmodel_ref.py).  Gate 1e-5."""
    import os
    import bench
    import nif_b200
    from tests.helpers import GOLDEN, C2_CFG_S, C2_CFG_P, fullmodel_problem
    assert bench.BATCH in (113664, 65536) and bench.CFG_S == C2_CFG_S and bench.CFG_P == C2_CFG_P  # the benchmark's own configuration
    ref = np.load(os.path.join(GOLDEN, "fullbatch", "c2_fullbatch_model_grad.npz" if B == 65536 else f"c2_fullbatch_model_grad_{B}.npz"))
    spec, prm, inputs, tgt = fullmodel_problem(B)
    dev = torch.device("cuda:0")
    net = nif_b200.NIFMultiScale(C2_CFG_S, C2_CFG_P, "float32", seed=0, device=dev)
    net.set_weights({k: v.numpy() for k, v in prm.items()})
    assert net.engine.kernel_path == "fp16x3"
    model = net.build()
    model.compile(nif_b200.Adam(1e-3), loss="mse", graph=False)
    loss = model.train_on_batch(inputs, tgt)
    errs = {"loss": abs(loss - float(ref["loss"])) / float(ref["loss"])}
    for k in net.variables:
        errs[k] = rel_err(net._gviews[k].cpu(), ref["g:" + k])
    print("whole-model full-batch errors vs fp64 oracle:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert all(v < 1e-5 for v in errs.values()), errs


def test_fit_without_shuffle_and_odd_batch_size():
    """Device-resident data set + shuffle=False hands the step row-slice views at arbitrary byte offsets: the host
    layer must re-align them for the C ABI.  Batch 50 with one output column puts the second target batch at offset 200 B."""
    import nif_b200
    dev = torch.device("cuda:0")
    cfg_s = {"input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
    cfg_p = {"input_dim": 1, "latent_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (230, 2)).astype(np.float32)
    Y = np.sin(3 * X[:, :1] + X[:, 1:]).astype(np.float32)
    sw = rng.uniform(0.5, 1.5, (230,)).astype(np.float32)
    losses = {}
    for graph in (False, True):
        net = nif_b200.NIF(cfg_s, cfg_p, seed=0, device=dev)
        model = net.build()
        model.compile(nif_b200.Adam(1e-3), loss="mse", graph=graph)
        h = model.fit(X, Y, batch_size=50, epochs=3, shuffle=False, sample_weight=sw)
        losses[graph] = h.history["loss"]
        assert np.all(np.isfinite(losses[graph]))
    assert losses[False] == losses[True]
    # user-provided CUDA slices through train_on_batch
    Xd, Yd = torch.as_tensor(X).to(dev), torch.as_tensor(Y).to(dev)
    assert np.isfinite(model.train_on_batch(Xd[3:53], Yd[3:53]))


def test_predict_latent_grid_on_a_tensor_core_model():
    """Model.predict_latent_grid on a 64-unit model whose training engine is the FP16x3 tensor-core one (the K = 0
    engine it derives must not inherit a compute mode that has no grouped kernels)."""
    import nif_b200
    dev = torch.device("cuda:0")
    cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 2, "output_dim": 1, "units": 64, "nlayers": 2,
             "weight_init_factor": 0.01, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 4, "units": 16, "nlayers": 1, "activation": "swish"}
    net = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=0, device=dev)
    assert net.engine.kernel_path == "fp16x3"
    model = net.build()
    rng = np.random.default_rng(1)
    lat = rng.normal(size=(3, 4)).astype(np.float32)
    grid = rng.uniform(-1, 1, (200, 2)).astype(np.float32)
    u = model.predict_latent_grid(lat, grid).cpu()
    spec = O.spec_from_cfg("NIFMultiScale", cfg_s, cfg_p)
    prm = {k: torch.as_tensor(v).double() for k, v in net.get_weights().items()}
    wn, bn = O.last_layer_names(spec)
    wg = O.hyper_linear(torch.as_tensor(lat).double(), prm[wn], prm[bn])
    for g in range(3):
        ref = O.shape_net(spec, torch.as_tensor(grid).double(), wg[g:g + 1].expand(200, -1))
        assert rel_err(u[g], ref) < 2e-5


def test_two_rank_nccl_data_parallel_equals_single_process():
    """N ranks produce the single-process parameters (tools/dp_check.py's body as a driver-run test): spawns 2 ranks over
    NCCL when at least 2 GPUs are visible, skips otherwise."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "dp_check.py")],
                       capture_output=True, text=True, cwd=root, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "DP check: all passed" in r.stdout and "params equal True" in r.stdout
