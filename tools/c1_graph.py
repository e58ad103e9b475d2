"""C1 (tutorial 1: NIF swish 2x30, latent 1, batch 512): eager launches vs CUDA-graph replay of the fused step."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import nif_b200  # noqa: E402

dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
cfg_s = {"input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
cfg_p = {"input_dim": 1, "latent_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
X = torch.as_tensor(rng.uniform(-1, 1, (512, 2)).astype(np.float32)).to(dev)
Y = torch.as_tensor(rng.uniform(-1, 1, (512, 1)).astype(np.float32)).to(dev)
out = {}
for graph in (False, True):
    net = nif_b200.NIF(cfg_s, cfg_p, seed=0, device=dev)
    m = net.build()
    m.compile(nif_b200.Adam(1e-3), loss="mse", graph=graph)
    for _ in range(5):
        m._train_step(X, Y, None, 512)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        m._train_step(X, Y, None, 512)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    out["graph" if graph else "eager"] = {"ms_per_step": ms, "points_per_s": 512 / ms * 1e3}
print(json.dumps({"config": "C1 tutorial-1 NIF swish 2x30, latent 1, batch 512", **out}))
