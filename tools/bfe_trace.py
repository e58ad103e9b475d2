"""Timeline of nif_bf_bwd_edge_kernel, CTA (0, 0) (needs the -DNIF_TRACE build):
    make -C nif_b200/csrc trace && NIF_B200_LIB=nif_b200/libnif_b200_trace.so python tools/bfe_trace.py [events]
"""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import nif_b200
from nif_b200 import _lib
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
for _ in range(2):
    m._train_step(X, Y, None, B)
torch.cuda.synchronize()
host = np.zeros((4, 2048), dtype=np.int64)
cnt = np.zeros(4, dtype=np.int32)
_lib.lib().nif_debug_read_trace_bfe(host.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
ev = []
names = {0: "A-gen ", 1: "B-gen0", 2: "mma   "}
ph = {0: "loop top", 1: "packed", 2: "slot free", 3: "stored + arrived", 4: "loop top", 5: "slot free", 6: "stored + arrived", 7: "operands ready: issue"}
for role in range(3):
    for i in range(0, cnt[role], 2):
        ev.append((int(host[role, i + 1]), role, int(host[role, i])))
ev.sort()
t0 = ev[0][0]
last = {}
for t, role, tag in ev[: int(sys.argv[1]) if len(sys.argv) > 1 else 200]:
    d = t - last.get(role, t)
    last[role] = t
    print(f"{t - t0:8d} (+{d:5d}) {names[role]} sub-tile {tag // 8:3d} {ph[tag % 8]}")
