"""Small invocations of every kernel family, the command compute-sanitizer wraps (tools/sanitize.sh):
FP16x3 tensor-core forward / reverse / weight / thin-term kernels, the bf16 kernels (padded widths 64 and 128), the
tensor-core trunk, the CUDA-core tile kernels, tangents, Adam.  Prints one line per family; shapes are tiny because the
sanitizer runs the kernels 10-100x slower."""
import os
import sys

import torch

sys.path.insert(0, '.')
from nif_b200.ops import FusedShapeNet, FusedTrunk, adam_step  # noqa: E402

dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)


def head(variant, si, so, n, l, K, B, compute):
    eng = FusedShapeNet(variant, si, so, n, l, K, "swish", 30.0, compute=compute)
    w = (torch.rand(K, eng.po_dim, generator=g) - 0.5).mul(0.05).to(dev)
    b = (torch.rand(eng.po_dim, generator=g) - 0.5).mul(0.05).to(dev)
    z = (torch.rand(B, K, generator=g) - 0.5).to(dev)
    x = (torch.rand(B, si, generator=g) * 2 - 1).to(dev)
    t = (torch.rand(B, so, generator=g) * 2 - 1).to(dev)
    packed = eng.pack(w, b)
    u, stash = eng.forward(z, x, packed, save=True)
    loss = torch.zeros(1, device=dev)
    dw, db = torch.empty_like(w), torch.empty_like(b)
    dz = eng.mse_backward(z, x, packed, u, stash, t, None, 1.0 / B, loss, dw, db)
    torch.cuda.synchronize()
    print(f"{compute:7s} {variant:6s} n={n:3d} K={K:2d} B={B}: path {eng.kernel_path}, loss {float(loss):.4f}, |dz| {float(dz.abs().max()):.3e}")
    return eng, z, x, packed


def tc_sobolev_and_wide_trunk():
    """Sobolev step on the tensor cores (tangent mode of the forward kernel, reverse-over-forward modes of the data kernel,
    two directions so that the accumulating mode runs too) and the fused elementwise kernels of a wide bf16 trunk."""
    import ctypes as C
    from nif_b200 import _lib
    eng, z, x, packed = head("siren", 2, 1, 64, 2, 5, 300, "fp16x3")
    xd = torch.zeros(2, 300, 2, device=dev)
    xd[0, :, 0] = 1
    xd[1, :, 1] = 1
    u, ud, stash = eng.forward_tangent(z, x, packed, None, xd, save=True)
    dw, db = torch.empty(5, eng.po_dim, device=dev), torch.empty(eng.po_dim, device=dev)
    dz = eng.sobolev_backward(z, x, xd, packed, stash, u * 1e-3, ud * 1e-3, dw, db, 0.0)
    u2, ud2 = eng.forward_tangent(z, x, packed, None, xd)
    torch.cuda.synchronize()
    print("tangent sobolev tensor cores |dz|", float(dz.abs().max()), "|dw|", float(dw.abs().max()), "same outputs", bool(torch.equal(ud, ud2)))
    L = _lib.lib()
    B, n = 333, 128
    y = torch.randn(B, n, generator=g).to(dev).to(torch.bfloat16)
    bias = torch.randn(n, generator=g).to(dev)
    h = torch.randn(B, n, generator=g).to(dev)
    ho, hb = torch.empty_like(h), torch.empty(B, n, dtype=torch.bfloat16, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.nif_trunk_ew_forward(B, n, 2, y.data_ptr(), bias.data_ptr(), h.data_ptr(), ho.data_ptr(), hb.data_ptr(), st), "ew fwd")
    gq, dbb = torch.empty_like(hb), torch.empty(n, device=dev)
    ws = torch.empty(int(L.nif_trunk_ew_ws_floats(n)), device=dev)
    _lib.check(L.nif_trunk_ew_backward(B, n, 2, y.data_ptr(), bias.data_ptr(), h.data_ptr(), hb.data_ptr(), h.data_ptr(),
                                       gq.data_ptr(), dbb.data_ptr(), ws.data_ptr(), st), "ew bwd")
    torch.cuda.synchronize()
    print("trunk elementwise |h|", float(ho.abs().max()), "|db|", float(dbb.abs().max()))


if os.environ.get("NIF_SANITIZE_ONLY") == "new":  # the kernels added last (the other families: the full run's logs)
    tc_sobolev_and_wide_trunk()
    sys.exit(0)
head("siren", 2, 1, 64, 2, 5, 300, "fp16x3")
head("nif", 2, 2, 48, 2, 3, 200, "fp16x3")
head("siren", 2, 1, 64, 2, 5, 300, "bf16")
head("siren", 3, 3, 128, 2, 9, 300, "bf16")
eng, z, x, packed = head("siren", 2, 1, 24, 2, 3, 200, "fp32")
xd = torch.zeros(1, 200, 2, device=dev)
xd[0, :, 0] = 1
u, ud = eng.forward_tangent(z, x, packed, None, xd)
torch.cuda.synchronize()
print("tangent |udot|", float(ud.abs().max()))
# reverse-over-forward over two directions, one of them moving the latent code (nif_sobolev_backward_dirs)
xd2 = torch.zeros(2, 200, 2, device=dev)
xd2[1, :, 1] = 1
zd2 = torch.zeros(2, 200, 3, device=dev)
zd2[0] = (torch.rand(200, 3, generator=g) - 0.5).to(dev)
u, ud, stash = eng.forward_tangent(z, x, packed, zd2, xd2, save=True)
dw, db = torch.empty(3, eng.po_dim, device=dev), torch.empty(eng.po_dim, device=dev)
dz, dzd = eng.sobolev_backward(z, x, xd2, packed, stash, u * 1e-3, ud * 1e-3, dw, db, 0.0, zdot=zd2)
torch.cuda.synchronize()
print("tangent sobolev |dz|", float(dz.abs().max()), "|dzdot|", float(dzd.abs().max()))
eng0 = FusedShapeNet("siren", 3, 1, 128, 2, 0, None, 30.0, compute="bf16")
wv = (torch.rand(2, eng0.po_dim, generator=g) - 0.5).mul(0.05).to(dev)
ug = eng0.forward(None, (torch.rand(200, 3, generator=g) * 2 - 1).to(dev), eng0.pack(None, wv), groups=2, x_shared=True)
torch.cuda.synchronize()
print("grouped bf16 |u|", float(ug.abs().max()))
tr = FusedTrunk(1, 32, 64, 4, "swish")
theta = (torch.rand(tr.n_theta, generator=g) - 0.5).mul(0.2).to(dev)
p = (torch.rand(300, 1, generator=g) * 2 - 1).to(dev)
zz, st = tr.forward(p, theta, save=True)
gt = torch.empty_like(theta)
tr.backward(p, theta, st, torch.rand(300, 32, generator=g).to(dev), gt, 0.0)
torch.cuda.synchronize()
print("trunk", tr.kernel_path, "|z|", float(zz.abs().max()), "|g|", float(gt.abs().max()))
m, v = torch.zeros_like(theta), torch.zeros_like(theta)
adam_step(theta, gt, m, v, 1e-3, 1)
torch.cuda.synchronize()
print("adam ok")
tc_sobolev_and_wide_trunk()
