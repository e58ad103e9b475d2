"""Timeline of the first tile pair of nif_tc_fwd_kernel on CTA 0 (needs a -DNIF_TRACE build of the library):
    make -C nif_b200/csrc trace && NIF_B200_LIB=nif_b200/libnif_b200_trace.so python tools/tc_trace.py
"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import bench
import nif_b200
from nif_b200 import _lib
from nif_b200.ops import FusedShapeNet

dev = torch.device('cuda:0')
B = 65536
net = nif_b200.NIFMultiScale(bench.CFG_S, bench.CFG_P, "float32", seed=0, device=torch.device("cuda:0"))  # the product's own initialiser
g = torch.Generator().manual_seed(0)
z = (torch.rand(B, 32, generator=g) - 0.5).to(dev)
x = (torch.rand(B, 2, generator=g) * 2 - 1).to(dev)
eng = FusedShapeNet("siren", 2, 1, 64, 4, 32, omega0=30.0, compute="fp16x3")
packed = eng.pack(net.w_h.detach(), net.b_h.detach())
L = _lib.lib()
which = sys.argv[2] if len(sys.argv) > 2 else "fwd"
host = np.zeros((4, 2048), dtype=np.int64)
cnt = np.zeros(4, dtype=np.int32)
if which == "fwd":
    for it in range(2):
        eng.forward(z, x, packed, save=True)
        torch.cuda.synchronize()
        L.nif_debug_read_trace(host.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
else:  # reverse data pass (nif_tc_bwd_data_kernel)
    u, stash = eng.forward(z, x, packed, save=True)
    tgt = torch.zeros_like(u)
    loss = torch.zeros(1, device=dev)
    dw = torch.empty(32, eng.po_dim, device=dev)
    db = torch.empty(eng.po_dim, device=dev)
    for it in range(2):
        eng.mse_backward(z, x, packed, u, stash, tgt, None, 1.0 / B, loss, dw, db)
        torch.cuda.synchronize()
        L.nif_debug_read_trace_bwd(host.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
ev = []
names = {0: "epi0", 1: "epi1", 2: "mma", 3: "prod"}
for role in range(4):
    for i in range(0, cnt[role], 2):
        ev.append((int(host[role, i + 1]), names[role], int(host[role, i])))
ev.sort()
t0 = ev[0][0]
sub = {"epi": {0: "wait_full", 1: "got_full", 2: "arrived_empty", 3: "layer_finished (fwd: before publish_h; bwd: da_m stored)"}, "mma": {0: "got_b_full", 1: "waits done", 2: "accumulator free (bwd)", 3: "committed0", 4: "loop top (bwd)", 5: "mmas issued (bwd)"}}
for t, who, tag in ev[:int(sys.argv[1]) if len(sys.argv) > 1 else 400]:
    if who.startswith("epi"):
        print(f"{t - t0:8d} {who} chunk {tag // 4:3d} {sub['epi'][tag % 4]}")
    elif who == "mma":
        print(f"{t - t0:8d} {who}  chunk {tag // 8:3d} {sub['mma'][tag % 8]}")
    else:
        print(f"{t - t0:8d} {who} chunk {tag:3d} stage free, load issued")
