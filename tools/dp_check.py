"""torchrun --nproc-per-node 2 tools/dp_check.py : data-parallel step == single-process step on the same global batch."""
import sys, torch
sys.path.insert(0, '.')
import nif_b200
from nif_b200.distributed import DataParallel
import numpy as np

dp = DataParallel("nccl")
dev = dp.device
torch.cuda.set_device(dev)
cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 2, "output_dim": 1, "units": 64, "nlayers": 4, "weight_init_factor": 0.01, "omega_0": 30.0}
cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 32, "units": 64, "nlayers": 4, "activation": "swish"}
rng = np.random.default_rng(0)
GB = 4096
X = rng.uniform(-1, 1, (GB, 3)).astype(np.float32); Y = rng.uniform(-1, 1, (GB, 1)).astype(np.float32)

def run(parallel, graph=None, steps=3):
    net = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=0, device=dev)
    m = net.build(); m.compile(nif_b200.Adam(1e-3), loss="mse", graph=graph)
    if parallel:
        dp.attach(m)
        xs, ys = X[dp.rank::dp.world], Y[dp.rank::dp.world]
    else:
        xs, ys = X, Y
    losses = []
    for _ in range(steps):
        l = m._train_step(torch.as_tensor(xs).to(dev), torch.as_tensor(ys).to(dev), None, GB).clone()
        if parallel: dp.allreduce_(l)
        losses.append(float(l))
    return net.theta.clone(), losses

th_dp, l_dp = run(True)
th_1, l_1 = run(False)
err = float((th_dp - th_1).abs().max() / th_1.abs().max())
if dp.rank == 0:
    print("DP check: world", dp.world, "losses dp", l_dp, "single", l_1, "max rel param diff after 3 steps", err)
assert err < 1e-5 and all(abs(a - b) < 1e-5 * max(1, abs(b)) for a, b in zip(l_dp, l_1)), (err, l_dp, l_1)
# the graph-replayed data-parallel step (two graphs around the first all-reduce) == the eager one, bit for bit
th_g, l_g = run(True, graph=True, steps=5)
th_e, l_e = run(True, graph=False, steps=5)
if dp.rank == 0:
    print("DP graph vs eager: params equal", bool(torch.equal(th_g, th_e)), "losses equal", l_g == l_e)
assert torch.equal(th_g, th_e) and l_g == l_e
dp.shutdown()
