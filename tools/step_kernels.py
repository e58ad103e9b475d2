import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench, nif_b200
from nif_b200.ops import kernel_profile
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
B = int(sys.argv[1])
net = nif_b200.NIFMultiScale(bench.CFG_S, bench.CFG_P, "float32", seed=0, device=dev)
m = net.build(); m.compile(nif_b200.Adam(1e-3), loss="mse", graph=False)
X = torch.as_tensor(rng.uniform(-1, 1, (B, 3)).astype(np.float32)).to(dev)
Y = torch.as_tensor(rng.uniform(-1, 1, (B, 1)).astype(np.float32)).to(dev)
for _ in range(5): m._train_step(X, Y, None, B)
torch.cuda.synchronize()
with kernel_profile() as prof:
    for _ in range(5): m._train_step(X, Y, None, B)
    torch.cuda.synchronize()
tot = 0
for k, c, t in sorted(prof.table, key=lambda r: -r[2]):
    print(f"{k:34s} {c / 5:4.1f} {t * 1e3 / 5:8.1f} us"); tot += t * 1e3 / 5
print("sum", tot)
