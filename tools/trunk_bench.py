"""Micro-benchmark of the tensor-core trunk kernels: python tools/trunk_bench.py"""
import sys
import torch
sys.path.insert(0, '.')
from nif_b200.ops import FusedTrunk

dev = torch.device('cuda:0')
tr = FusedTrunk(1, 32, 64, 4, "swish")
g = torch.Generator().manual_seed(0)
theta = (torch.rand(tr.n_theta, generator=g) - 0.5).mul(0.2).to(dev)


def ev(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for B in (256 * 148, 65536, 4 * 256 * 148):
    p = (torch.rand(B, 1, generator=g) * 2 - 1).to(dev)
    dz = torch.rand(B, 32, generator=g).to(dev)
    gt = torch.empty_like(theta)
    z, st = tr.forward(p, theta, save=True)
    print(f"B={B}: fwd+stash {ev(lambda: tr.forward(p, theta, save=True)):.1f} us  fwd only {ev(lambda: tr.forward(p, theta)):.1f} us  "
          f"bwd {ev(lambda: tr.backward(p, theta, st, dz, gt, 0.0)):.1f} us")
