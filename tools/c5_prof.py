"""One grouped latent x grid sweep call (C5 shapes, smaller grid) for an ncu capture of nif_bf_group_fwd_kernel."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import nif_b200
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 3, "output_dim": 3, "units": 128, "nlayers": 6,
         "weight_init_factor": 0.01, "omega_0": 30.0}
cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 64, "units": 128, "nlayers": 4, "activation": "swish"}
net = nif_b200.NIFMultiScale(cfg_s, cfg_p, "mixed_bfloat16", seed=0, device=dev)
m = net.build()
G, side = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 64
lin = np.linspace(-1, 1, side, dtype=np.float32)
grid = torch.as_tensor(np.stack(np.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(-1, 3)).to(dev)
lat = torch.as_tensor(rng.normal(size=(G, 64)).astype(np.float32)).to(dev)
for _ in range(2):
    out = m.predict_latent_grid(lat, grid)
torch.cuda.synchronize()
print(out.shape)
