"""fp64 oracle gradients of the WHOLE C2 model (ParameterNet trunk + hyper-network head + ShapeNet) at the full benchmark
batch (65 536 rows), summed over 4096-row chunks: the fixture behind
tests/test_gpu_parity.py::test_full_batch_whole_model_step_against_oracle.

    python tests/golden/fullbatch/make_fullmodel_ref.py [B]  (CPU, a few minutes; writes c2_fullbatch_model_grad.npz next to
                                                              it, or c2_fullbatch_model_grad_<B>.npz for another batch:
                                                              113 664 rows = 148 SMs x 3 pairs of 128-row tiles is bench.py's)

Inputs are regenerated from seeds by the test (tests/helpers.py::fullmodel_problem); only the reference gradients and the
loss are stored (fp32 storage of fp64 sums: 6e-8 relative, far below the 1e-5 gate)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
from oracle import nif_oracle as O  # noqa: E402
from tests.helpers import fullmodel_problem  # noqa: E402

B_ARG = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
spec, prm, inputs, tgt = fullmodel_problem(B_ARG)
B = inputs.shape[0]
p64 = {k: v.double().requires_grad_(True) for k, v in prm.items()}
names = list(p64)
tot = {k: torch.zeros_like(v) for k, v in p64.items()}
loss_tot = 0.0
for s in range(0, B, 4096):
    y = O.forward(spec, p64, inputs[s:s + 4096].double())
    loss = ((y - tgt[s:s + 4096].double()) ** 2).mean(-1).sum() / B
    g = torch.autograd.grad(loss, [p64[k] for k in names])
    for k, gk in zip(names, g):
        tot[k] += gk
    loss_tot += float(loss)
    print(s, flush=True)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fullbatch",
                                 "c2_fullbatch_model_grad.npz" if B == 65536 else f"c2_fullbatch_model_grad_{B}.npz"),
                    loss=np.float64(loss_tot), **{"g:" + k: v.numpy().astype(np.float32) for k, v in tot.items()})
print("loss", loss_tot)
