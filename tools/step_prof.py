"""A few C2 training steps (bench.py's workload, no timing, no CPU arm): the command ncu wraps.

    python tools/step_prof.py [steps] [batch]
"""
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
import nif_b200  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.BATCH
dev = torch.device('cuda:0')
net = nif_b200.NIFMultiScale(bench.CFG_S, bench.CFG_P, "float32", seed=0, device=dev)
model = net.build()
model.compile(nif_b200.Adam(1e-3), loss="mse")
inp, tgt = bench.synth_c2(B * 2, seed=100)
inp, tgt = torch.as_tensor(inp).to(dev), torch.as_tensor(tgt).to(dev)
for i in range(steps):
    j = (i % 2) * B
    loss = model._train_step(inp[j:j + B], tgt[j:j + B], None, B)
torch.cuda.synchronize()
print("loss", float(loss))
