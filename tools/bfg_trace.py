"""Timeline of the first tile pairs of nif_bf_group_fwd_kernel on CTA 0 (needs the -DNIF_TRACE build):
    make -C nif_b200/csrc trace && NIF_B200_LIB=nif_b200/libnif_b200_trace.so python tools/bfg_trace.py [events]
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
m = net.build()
G, side = 16, 64
lin = np.linspace(-1, 1, side, dtype=np.float32)
grid = torch.as_tensor(np.stack(np.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(-1, 3)).to(dev)
lat = torch.as_tensor(rng.normal(size=(G, 64)).astype(np.float32)).to(dev)
for _ in range(2):
    m.predict_latent_grid(lat, grid)
torch.cuda.synchronize()
host = np.zeros((4, 2048), dtype=np.int64)
cnt = np.zeros(4, dtype=np.int32)
_lib.lib().nif_debug_read_trace_bfg(host.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
ev = []
names = {0: "epi.w0 ", 1: "epi.w13", 2: "mma    "}
ph = {0: "wait t_full", 1: "got t_full", 2: "acc drained (ld + fma done)", 3: "sin done", 4: "published", 5: "next accumulator requested"}
for role in range(3):
    for i in range(0, cnt[role], 2):
        ev.append((int(host[role, i + 1]), role, int(host[role, i])))
ev.sort()
t0 = ev[0][0]
last = {}
for t, role, tag in ev[: int(sys.argv[1]) if len(sys.argv) > 1 else 300]:
    step, p = tag // 8, tag % 8
    d = t - last.get(role, t)
    last[role] = t
    what = ph.get(p, str(p)) if role < 2 else "issue MMAs"
    print(f"{t - t0:8d} (+{d:5d}) {names[role]} layer/stage {step // 2:2d} tile {step % 2} {what}")
