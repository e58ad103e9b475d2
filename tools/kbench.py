"""Kernel micro-benchmark at the C2 shape: CUDA events around individual C-ABI calls."""
import sys, torch
sys.path.insert(0, '.')
import bench
import nif_b200
from nif_b200.ops import FusedShapeNet

dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
net = nif_b200.NIFMultiScale(bench.CFG_S, bench.CFG_P, "float32", seed=0, device=torch.device("cuda:0"))  # the product's own initialiser
g = torch.Generator().manual_seed(0)
z = (torch.rand(B, 32, generator=g) - 0.5).to(dev)
x = (torch.rand(B, 2, generator=g) * 2 - 1).to(dev)
tgt = torch.rand(B, 1, generator=g).to(dev)
w_h, b_h = net.w_h.detach(), net.b_h.detach()
flops = (2 * 32 * 16897 + 2 * (128 + 4 * 4096 + 64)) * B

def ev(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for comp in ("fp32", "fp16x3"):
    eng = FusedShapeNet("siren", 2, 1, 64, 4, 32, omega0=30.0, compute=comp)
    packed = eng.pack(w_h, b_h)
    t_pack = ev(lambda: eng.pack(w_h, b_h, out=packed))
    t_inf = ev(lambda: eng.forward(z, x, packed))
    t_fwd = ev(lambda: eng.forward(z, x, packed, save=True))
    u, stash = eng.forward(z, x, packed, save=True)
    loss = torch.zeros(1, device=dev); dw, db = torch.empty_like(w_h), torch.empty_like(b_h)
    t_bwd = ev(lambda: eng.mse_backward(z, x, packed, u, stash, tgt, None, 1.0 / B, loss, dw, db))
    print(f"{comp:7s} B={B}: pack {t_pack*1e3:7.1f} us | fwd(inference) {t_inf:7.3f} ms = {flops/t_inf/1e9:7.1f} TFLOP/s | "
          f"fwd(+stash) {t_fwd:7.3f} ms = {flops/t_fwd/1e9:7.1f} TFLOP/s | reverse {t_bwd:7.3f} ms = {2*flops/t_bwd/1e9:7.1f} TFLOP/s")
