"""Multi-GPU measurements of the BASELINE configurations bench.py does not cover (one process per GPU, torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/scale_configs.py [out.json]

  C3  turbulence ShapeNet 6x128 SIREN, latent 64, mixed_bfloat16, data parallel WEAK scaling (65 536 rows per GPU, one NCCL
      all-reduce of the flat gradient per step), rows/s of the whole job
  C5  latent sweep 6x128: G latents x 64^3 grid points, the latent axis sharded over the ranks (no collective), evals/s
  C2  STRONG scaling: bench.py's model at a fixed global batch of 65 536 rows split over the ranks
Timing: CUDA events, barrier + synchronize on both sides, max over ranks."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import nif_b200  # noqa: E402
from nif_b200.distributed import DataParallel  # noqa: E402

dp = DataParallel("nccl")
dev = dp.device
torch.cuda.set_device(dev)
rank, world = dp.rank, dp.world


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(); dp.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(); dp.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dp.max_(ms)
    return float(ms) / steps


out = []
rng = np.random.default_rng(100 + rank)
cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 3, "output_dim": 3, "units": 128, "nlayers": 6,
         "weight_init_factor": 0.01, "omega_0": 30.0}
cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 64, "units": 128, "nlayers": 4, "activation": "swish"}
net = nif_b200.NIFMultiScale(cfg_s, cfg_p, "mixed_bfloat16", seed=0, device=dev)
m = net.build(); m.compile(nif_b200.Adam(1e-3), loss="mse")
if world > 1:
    dp.attach(m)
B = 65536
X = torch.as_tensor(rng.uniform(-1, 1, (B, 4)).astype(np.float32)).to(dev)
Y = torch.as_tensor(rng.uniform(-1, 1, (B, 3)).astype(np.float32)).to(dev)
ms = timed(lambda: m._train_step(X, Y, None, B * world), 10)
lo, hi = net.theta.clone(), net.theta.clone()
if world > 1:
    torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
    torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
out.append({"config": "C3 turbulence 6x128 SIREN latent 64 mixed_bfloat16, weak scaling 65536 rows/GPU", "n_gpus": world,
            "kernels": net.engine.kernel_path, "ms_per_step": ms, "rows_per_s": B * world / ms * 1e3,
            "params_equal_across_ranks": bool(torch.equal(lo, hi))})

cfg_s5 = dict(cfg_s, output_dim=1)
net5 = nif_b200.NIFMultiScale(cfg_s5, cfg_p, "mixed_bfloat16", seed=0, device=dev)
m5 = net5.build()
if world > 1:
    dp.attach(m5)
G, side = 64 * world, 64
lin = np.linspace(-1, 1, side, dtype=np.float32)
grid = torch.as_tensor(np.stack(np.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(-1, 3)).to(dev)
lat = torch.as_tensor(np.random.default_rng(7).normal(size=(G, 64)).astype(np.float32)).to(dev)
ms5 = timed(lambda: m5.predict_latent_grid(lat, grid, shard=True), 5)
out.append({"config": f"C5 latent sweep 6x128 mixed_bfloat16: {G} latents x {side}^3 grid, latent axis sharded over the ranks",
            "n_gpus": world, "ms_per_call": ms5, "evals_per_s": G * grid.shape[0] / ms5 * 1e3,
            "full_sweep_4096x256^3_seconds_at_this_rate": 4096 * 256**3 / (G * grid.shape[0] / ms5 * 1e3)})
# ---- C2 STRONG scaling: the bench.py model at a FIXED global batch of 65 536 rows, split over the ranks ----
import bench  # noqa: E402
net2 = nif_b200.NIFMultiScale(bench.CFG_S, bench.CFG_P, "float32", seed=0, device=dev)
m2 = net2.build(); m2.compile(nif_b200.Adam(1e-3), loss="mse")
if world > 1:
    dp.attach(m2)
GB2 = 65536
b2 = GB2 // world
X2 = torch.as_tensor(rng.uniform(-1, 1, (b2, 3)).astype(np.float32)).to(dev)
Y2 = torch.as_tensor(rng.uniform(-1, 1, (b2, 1)).astype(np.float32)).to(dev)
ms2 = timed(lambda: m2._train_step(X2, Y2, None, GB2), 50, 5)
out.append({"config": f"C2 strong scaling: global batch {GB2} rows split over the ranks ({b2} rows per GPU)", "n_gpus": world,
            "scaling": "strong", "ms_per_step": ms2, "rows_per_s": GB2 / ms2 * 1e3,
            "dp_update": "single" if world == 1 else ("multimem" if m2._symm is not None else "nccl")})
if rank == 0:
    for o in out:
        print(json.dumps(o))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
dp.shutdown()
