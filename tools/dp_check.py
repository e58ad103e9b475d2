"""torchrun --nproc-per-node N tools/dp_check.py : the data-parallel step == the single-process step on the same global
batch, for both update paths (NVSwitch multicast fused reduce-scatter + Adam + all-gather kernel | NCCL all-reduce + Adam),
eager and graph-replayed; replicas bit-identical."""
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


def run(parallel, graph=None, steps=3, symmetric=None):
    net = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=0, device=dev)
    m = net.build(); m.compile(nif_b200.Adam(1e-3), loss="mse", graph=graph)
    mode = "single"
    if parallel:
        dp.attach(m, symmetric=symmetric)
        mode = "multimem" if m._symm is not None else "nccl"
        xs, ys = X[dp.rank::dp.world], Y[dp.rank::dp.world]
    else:
        xs, ys = X, Y
    losses = []
    for _ in range(steps):
        l = m._train_step(torch.as_tensor(xs).to(dev), torch.as_tensor(ys).to(dev), None, GB).clone()
        if parallel: dp.allreduce_(l)
        losses.append(float(l))
    torch.cuda.synchronize()
    th = net.theta.clone()
    lo, hi = th.clone(), th.clone()
    if parallel and dp.world > 1:
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
    return th, losses, mode, bool(torch.equal(lo, hi))


th_1, l_1, _, _ = run(False, graph=False)
for symmetric in (True, False):
    th_dp, l_dp, mode, same = run(True, graph=False, symmetric=symmetric)
    err = float((th_dp - th_1).abs().max() / th_1.abs().max())
    if dp.rank == 0:
        print(f"DP check [{mode}]: world {dp.world} losses dp {l_dp} single {l_1} max rel param diff after 3 steps {err:.3e} replicas identical {same}")
    tol = 1e-5 if dp.world <= 2 else 1e-4  # more partial sums, and Adam's m / (sqrt(v) + eps) amplifies them where g ~ 0
    assert err < tol and same and all(abs(a - b) < 1e-5 * max(1, abs(b)) for a, b in zip(l_dp, l_1)), (mode, err, l_dp, l_1)
    # graph-replayed data-parallel steps == eager ones, bit for bit
    th_g, l_g, mode_g, same_g = run(True, graph=True, steps=5, symmetric=symmetric)
    th_e, l_e, _, _ = run(True, graph=False, steps=5, symmetric=symmetric)
    if dp.rank == 0:
        print(f"DP graph vs eager [{mode_g}]: params equal {bool(torch.equal(th_g, th_e))} losses equal {l_g == l_e} replicas identical {same_g}")
    assert torch.equal(th_g, th_e) and l_g == l_e and same_g
if dp.rank == 0:
    print("DP check: all passed; symmetric-memory error:", getattr(dp, "_symm_error", None))
dp.shutdown()
