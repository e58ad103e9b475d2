"""Kernel table of one C4 (Sobolev) step: python tools/c4_profile.py [cols]  (cols: grad columns, default 2 = du/dx)"""
import sys, json
import numpy as np
import torch
sys.path.insert(0, '.')
import nif_b200
from nif_b200.ops import kernel_profile
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 64, "nlayers": 4,
         "weight_init_factor": 0.01, "omega_0": 30.0}
cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 32, "units": 64, "nlayers": 4, "activation": "swish"}
cols = [int(c) for c in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2]
net = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=0, device=dev)
m = nif_b200.JacobianLayer(net.build(), [0], [0, 1]).as_model()
m.compile(nif_b200.Adam(1e-3), loss=nif_b200.SobolevMSE(1e-3, [0], cols))
B = 65536
X = torch.as_tensor(rng.uniform(-1, 1, (B, 2)).astype(np.float32)).to(dev)
Y = torch.as_tensor(rng.uniform(-1, 1, (B, 3)).astype(np.float32)).to(dev)
for _ in range(3): m._train_step(X, Y, None, B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): m._train_step(X, Y, None, B)
e1.record(); torch.cuda.synchronize()
print("ms/step", e0.elapsed_time(e1) / 5, "grad cols", cols)
m.use_graph = False  # the per-kernel events need eager launches
m._train_step(X, Y, None, B)
with kernel_profile() as prof:
    for _ in range(3): m._train_step(X, Y, None, B)
    torch.cuda.synchronize()
for k, c, t in sorted(prof.table, key=lambda r: -r[2]): print(f"{k:34s} {c / 3:5.1f} launches/step {t * 1e3 / 3:9.1f} us/step")
