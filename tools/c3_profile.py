"""torch.profiler kernel list of one C3 step (eager launches): where the time outside the library kernels goes."""
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
m = net.build(); m.compile(nif_b200.Adam(1e-3), loss="mse", graph=False)
B = 65536
X = torch.as_tensor(rng.uniform(-1, 1, (B, 4)).astype(np.float32)).to(dev)
Y = torch.as_tensor(rng.uniform(-1, 1, (B, 3)).astype(np.float32)).to(dev)
for _ in range(5): m._train_step(X, Y, None, B)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): m._train_step(X, Y, None, B)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
