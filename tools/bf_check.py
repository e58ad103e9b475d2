"""bf16 tensor-core path against the oracle (GPU): forward, stash-free; prints errors per shape.

    python tools/bf_check.py
"""
import sys

import torch

sys.path.insert(0, '.')
from oracle import nif_oracle as O  # noqa: E402
from nif_b200.ops import FusedShapeNet  # noqa: E402

dev = torch.device('cuda:0')


def rel(a, b):
    return float((a.double().cpu() - b).abs().max() / b.abs().max())


cases = [("siren", 2, 1, 64, 2, 3, 300), ("siren", 3, 3, 128, 2, 5, 300), ("siren", 2, 1, 64, 4, 32, 1000),
         ("siren", 3, 3, 128, 6, 64, 700), ("nif", 2, 2, 48, 3, 7, 90), ("nif", 2, 2, 100, 2, 6, 257),
         ("siren", 3, 1, 128, 3, 0, 500)]
for variant, si, so, n, l, K, B in cases:
    spec = O.Spec(variant=variant, pi=1, si=si, so=so, n=n, l=l, K=max(K, 1), n_st=16, l_st=1, p_act="swish", omega0=30.0,
                  weight_init_factor=0.01, s_act="swish")
    prm = {k: v.double() for k, v in O.init_params(spec, 1).items()}
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, si, generator=g, dtype=torch.float64) * 2 - 1
    z = torch.rand(B, max(K, 1), generator=g, dtype=torch.float64) - 0.5
    wn, bn = O.last_layer_names(spec)
    w_h, b_h = prm[wn], prm[bn]
    if K == 0:  # grouped: 3 groups of explicit weight vectors, shared grid
        G = 3
        zg = torch.rand(G, 1, generator=g, dtype=torch.float64) - 0.5
        wv = zg @ w_h + b_h  # [G, P]
        eng = FusedShapeNet(variant, si, so, n, l, 0, "swish", 30.0, compute="bf16")
        print(eng.kernel_path, end=" ")
        packed = eng.pack(None, wv.float().to(dev))
        u = eng.forward(None, x.float().to(dev), packed, groups=G, x_shared=True).view(G, B, so)
        torch.cuda.synchronize()
        for gi in range(G):
            zz = torch.zeros(B, 0, dtype=torch.float64)
            ref = O.shape_net_factored(spec, x, zz, torch.zeros(0, wv.shape[1], dtype=torch.float64), wv[gi])
            emu = O.shape_net_factored(spec, x, zz, torch.zeros(0, wv.shape[1], dtype=torch.float64), wv[gi], quant="bf16")
            print(f"grouped g={gi}: vs fp64 {rel(u[gi], ref):.2e}  vs bf16-emulation {rel(u[gi], emu):.2e}")
        continue
    eng = FusedShapeNet(variant, si, so, n, l, K, "swish", 30.0, compute="bf16")
    packed = eng.pack(w_h.float().to(dev), b_h.float().to(dev))
    u = eng.forward(z.float().to(dev), x.float().to(dev), packed)
    u2, stash = eng.forward(z.float().to(dev), x.float().to(dev), packed, save=True)
    torch.cuda.synchronize()
    ref = O.shape_net_factored(spec, x, z, w_h, b_h)
    emu = O.shape_net_factored(spec, x, z.float().double(), w_h.float().double(), b_h.float().double(), quant="bf16")
    print(f"{eng.kernel_path} {variant} si={si} so={so} n={n} l={l} K={K} B={B}: vs fp64 {rel(u, ref):.2e}  vs bf16-emulation "
          f"{rel(u, emu):.2e}  save==nosave {bool((u == u2).all())}")
    # reverse pass against fp64 autograd of the factored form
    tgt = torch.rand(B, so, generator=g, dtype=torch.float64) * 2 - 1
    zq, wq, bq = z.clone().requires_grad_(True), w_h.clone().requires_grad_(True), b_h.clone().requires_grad_(True)
    y = O.shape_net_factored(spec, x, zq, wq, bq)
    loss64 = ((y - tgt) ** 2).mean(-1).mean()
    loss64.backward()
    loss = torch.zeros(1, device=dev)
    dw, db = torch.empty_like(w_h, dtype=torch.float32, device=dev), torch.empty_like(b_h, dtype=torch.float32, device=dev)
    dz = eng.mse_backward(z.float().to(dev), x.float().to(dev), packed, u2, stash, tgt.float().to(dev), None, 1.0 / B, loss, dw, db)
    torch.cuda.synchronize()
    print(f"    loss {abs(float(loss) - float(loss64)) / float(loss64):.2e}  dw {rel(dw, wq.grad):.2e}  db {rel(db, bq.grad):.2e}  "
          f"dz {rel(dz, zq.grad):.2e}")
