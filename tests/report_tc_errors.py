"""Error report (not a pytest module): fp32 CUDA-core and FP16x3 tensor-core kernels against the fp64 oracle, next to what
fp32 on the CPU gives.    python tests/report_tc_errors.py"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import nif_oracle as O
from nif_b200.ops import FusedShapeNet
from tests.helpers import rel_err
from tests.test_gpu_parity import _random_problem
dev = torch.device('cuda:0')
for (variant, si, so, n, l, K, B, seed) in [("siren",2,1,64,4,32,2048,1), ("siren",2,1,64,4,32,2048,2), ("siren_res",2,2,64,2,6,150,270), ("siren",1,1,64,4,32,128,196)]:
    spec, prm, inputs, target, sw = _random_problem(variant, si, so, n, l, K, B, seed=seed)
    wn, bn = O.last_layer_names(spec)
    loss64, g64, gz64, y64 = O.loss_and_grads(spec, prm, inputs, target, sw)
    prm32 = {k: v.float() for k, v in prm.items()}
    loss32, g32, gz32, y32 = O.loss_and_grads(spec, prm32, inputs.float(), target.float(), sw.float())
    z = O.latent(spec, prm, inputs[:, :1]).float().to(dev); x = inputs[:, 1:].float().contiguous().to(dev)
    w_h, b_h = prm[wn].float().to(dev), prm[bn].float().to(dev)
    out = {"cpu32": (rel_err(y32, y64), rel_err(gz32, gz64), rel_err(g32[wn], g64[wn]))}
    for comp in ("fp32", "fp16x3"):
        eng = FusedShapeNet(spec.variant, spec.si, spec.so, spec.n, spec.l, spec.K, spec.s_act, spec.omega0, compute=comp)
        packed = eng.pack(w_h, b_h)
        u, stash = eng.forward(z, x, packed, save=True)
        loss = torch.zeros(1, device=dev); dw, db = torch.empty_like(w_h), torch.empty_like(b_h)
        dz = eng.mse_backward(z, x, packed, u, stash, target.float().to(dev), sw.float().to(dev), 1.0 / B, loss, dw, db)
        out[comp] = (rel_err(u.cpu(), y64), rel_err(dz.cpu(), gz64), rel_err(dw.cpu(), g64[wn]))
    print(variant, n, l, K, B, {k: tuple(f"{e:.2e}" for e in v) for k, v in out.items()})
