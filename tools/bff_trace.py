"""Timeline of nif_bf_fwd_kernel, CTA 0 (needs the -DNIF_TRACE build):
    make -C nif_b200/csrc trace && NIF_B200_LIB=nif_b200/libnif_b200_trace.so python tools/bff_trace.py [first] [count]
"""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import nif_b200
from nif_b200 import _lib
from nif_b200.ops import FusedShapeNet
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
eng = FusedShapeNet("siren", 3, 3, 128, 6, 64, None, 30.0, compute="bf16")
B = 65536
w = (torch.rand(64, eng.po_dim, generator=g) - 0.5).mul(0.02).to(dev)
b = (torch.rand(eng.po_dim, generator=g) - 0.5).mul(0.02).to(dev)
z = (torch.rand(B, 64, generator=g) - 0.5).to(dev)
x = (torch.rand(B, 3, generator=g) * 2 - 1).to(dev)
packed = eng.pack(w, b)
for _ in range(2):
    eng.forward(z, x, packed, save=True)
torch.cuda.synchronize()
host = np.zeros((4, 2048), dtype=np.int64)
cnt = np.zeros(4, dtype=np.int32)
_lib.lib().nif_debug_read_trace_bfg(host.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
ev = []
names = {0: "epi ", 2: "mma ", 3: "prod"}
ph = {0: "loop top", 1: "accumulator free", 2: "weights landed: issue", 3: "wait t_full", 4: "got t_full", 5: "drained"}
for role in (0, 2, 3):
    for i in range(0, cnt[role], 2):
        ev.append((int(host[role, i + 1]), role, int(host[role, i])))
ev.sort()
t0 = ev[0][0]
last = {}
first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
for t, role, tag in ev[first:first + count]:
    d = t - last.get(role, t)
    last[role] = t
    what = "stage free" if role == 3 else ph[tag % 8]
    print(f"{t - t0:8d} (+{d:5d}) {names[role]} chunk {tag // 8:4d} {what}")
