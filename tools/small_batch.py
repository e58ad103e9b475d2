"""Step time of the bench.py model (C2) at small batches, graph-replayed: where the fixed per-step costs (weight packing,
trunk kernels, un-packing, Adam, launch latencies) dominate.  python tools/small_batch.py [out.json]
NIF_BATCHES=37888,65536,75776 measures other batches: the tensor-core forward / reverse kernels work on pairs of 128-row tiles,
one CTA per SM, so a batch that is a multiple of 148 x 256 = 37 888 rows fills their last wave."""
import json
import os
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
import nif_b200
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
out = []
batches = [int(b) for b in os.environ.get("NIF_BATCHES", "512,1024,4096,8192,16384,65536").split(",")]
for B in batches:
    net = nif_b200.NIFMultiScale(bench.CFG_S, bench.CFG_P, "float32", seed=0, device=dev)
    m = net.build(); m.compile(nif_b200.Adam(1e-3), loss="mse")
    X = torch.as_tensor(rng.uniform(-1, 1, (B, 3)).astype(np.float32)).to(dev)
    Y = torch.as_tensor(rng.uniform(-1, 1, (B, 1)).astype(np.float32)).to(dev)
    for _ in range(5):
        m._train_step(X, Y, None, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200 if B <= 8192 else 50
    e0.record()
    for _ in range(n):
        m._train_step(X, Y, None, B)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out.append({"batch": B, "ms_per_step": ms, "rows_per_s": B / ms * 1e3})
    print(json.dumps(out[-1]))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
