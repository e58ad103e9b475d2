"""fp64 oracle gradient of the hyper-network head at the FULL C2 batch (65 536 rows), summed over 4096-row chunks:
the fixture behind tests/test_gpu_parity.py::test_full_batch_head_gradient_against_oracle (the tensor-core weight-
gradient kernel accumulates 16 384 rows per TMEM accumulator; small-batch parity says nothing about that chain).

    python tests/golden/fullbatch/make_fullbatch_ref.py     (CPU, ~1 min; writes c2_fullbatch_head_grad.npz next to it)

Inputs are regenerated from seeds by the test (tests/helpers.py::fullbatch_problem), only the reference gradients and
the loss are stored (fp32 storage of fp64 sums: 6e-8 relative, far below the 1e-5 gate)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
from oracle import nif_oracle as O  # noqa: E402
from tests.helpers import fullbatch_problem  # noqa: E402

spec, prm, z, x, tgt = fullbatch_problem()
wn, bn = O.last_layer_names(spec)
w = prm[wn].double().requires_grad_(True)
b = prm[bn].double().requires_grad_(True)
B = z.shape[0]
loss_tot = 0.0
gw = torch.zeros_like(w)
gb = torch.zeros_like(b)
dz = []
for s in range(0, B, 4096):
    zz = z[s:s + 4096].double().requires_grad_(True)
    u = O.shape_net(spec, x[s:s + 4096].double(), O.hyper_linear(zz, w, b))
    loss = ((u - tgt[s:s + 4096].double()) ** 2).mean(-1).sum() / B
    g = torch.autograd.grad(loss, [w, b, zz])
    gw += g[0]
    gb += g[1]
    dz.append(g[2])
    loss_tot += float(loss)
    print(s, flush=True)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fullbatch", "c2_fullbatch_head_grad.npz"), dw=gw.numpy().astype(np.float32),
                    db=gb.numpy().astype(np.float32), dz_head=torch.cat(dz)[:256].numpy().astype(np.float32),
                    loss=np.float64(loss_tot))
print("loss", loss_tot)
