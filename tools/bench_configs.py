"""Secondary measurements: the other BASELINE.json configurations (C1, C3, C4, C5) on one GPU.

    python tools/bench_configs.py [out.json]        (NIF_CONFIGS=C3,C4 restricts the run to the named configurations)

Prints one JSON object per configuration: points/s of one optimisation step (C1, C3, C4) or of the grouped
inference sweep (C5, scaled to what one call holds), CUDA-event timed after warm-up.  Synthetic inputs,
random-initialised weights of the named architecture.  bench.py stays the headline (C2).
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import nif_b200  # noqa: E402

dev = torch.device('cuda:0')


def ev_time(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def flops_step(pi, si, so, n, l, K, n_st, l_st, tangents=0, tangents_p=0):
    W_s = si * n + l * n * n + n * so
    P = W_s + (l + 1) * n + so
    F_trunk = 2 * (pi * n_st + l_st * n_st * n_st + n_st * K)
    F_fwd = (1 + tangents_p) * 2 * K * P + (1 + tangents) * 2 * W_s + (1 + tangents_p) * F_trunk
    return 3 * F_fwd, P


out = []
rng = np.random.default_rng(0)
_want = [c for c in os.environ.get("NIF_CONFIGS", "").split(",") if c]


def want(c):
    return not _want or c in _want


# ---- C1: tutorial 1, NIF swish 2x30, latent 1, batch 512 (launch-latency bound) ----
if want("C1"):
    cfg_s = {"input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
    cfg_p = {"input_dim": 1, "latent_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
    net = nif_b200.NIF(cfg_s, cfg_p, seed=0, device=dev)
    m = net.build(); m.compile(nif_b200.Adam(1e-3), loss="mse")
    X = torch.as_tensor(rng.uniform(-1, 1, (512, 2)).astype(np.float32)).to(dev)
    Y = torch.as_tensor(rng.uniform(-1, 1, (512, 1)).astype(np.float32)).to(dev)
    ms = ev_time(lambda: m._train_step(X, Y, None, 512), 50)
    F, P = flops_step(1, 1, 1, 30, 2, 1, 30, 2)
    out.append({"config": "C1 tutorial-1 NIF swish 2x30, latent 1, batch 512", "po_dim": P, "ms_per_step": ms,
                "points_per_s": 512 / ms * 1e3, "note": "graph-replayed step (17 kernels per step; 0.178 ms with eager launches)"})

# ---- C3: turbulence, ShapeNet 3->6x128->3 SIREN, ParameterNet 1->4x128->latent 64 ----
if want("C3"):
    cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 3, "output_dim": 3, "units": 128, "nlayers": 6,
             "weight_init_factor": 0.01, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 64, "units": 128, "nlayers": 4, "activation": "swish"}
    net = nif_b200.NIFMultiScale(cfg_s, cfg_p, "mixed_bfloat16", seed=0, device=dev)
    m = net.build(); m.compile(nif_b200.Adam(1e-3), loss="mse")
    B = int(os.environ.get("NIF_C3_BATCH", "65536"))  # (75 776 = 148 SMs x 4 tiles of 128 rows fills the bf16 kernels' last wave)
    X = torch.as_tensor(rng.uniform(-1, 1, (B, 4)).astype(np.float32)).to(dev)
    Y = torch.as_tensor(rng.uniform(-1, 1, (B, 3)).astype(np.float32)).to(dev)
    ms = ev_time(lambda: m._train_step(X, Y, None, B), 5)
    F, P = flops_step(1, 3, 3, 128, 6, 64, 128, 4)
    from nif_b200.ops import kernel_profile  # noqa: E402
    m.use_graph = False  # the per-kernel events need eager launches
    m._train_step(X, Y, None, B)
    with kernel_profile() as prof:
        for _ in range(3):
            m._train_step(X, Y, None, B)
    table = [{"kernel": k, "us_per_step": t * 1e3 / 3, "launches_per_step": c / 3} for k, c, t in sorted(prof.table, key=lambda r: -r[2])]
    out.append({"config": f"C3 turbulence ShapeNet 6x128 SIREN, latent 64, batch {B}, mixed_bfloat16 (kernels: "
                          f"{net.engine.kernel_path})", "po_dim": P, "ms_per_step": ms, "points_per_s": B / ms * 1e3,
                "algorithmic_tflops": F * B / ms / 1e9, "library_kernels": table,
                "library_kernels_us": sum(r["us_per_step"] for r in table)})

# ---- C4: Sobolev training, ShapeNet 1->4x64->1, JacobianLayer(y=[0], x=[0,1]) ----
if want("C4"):
    cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 64, "nlayers": 4,
             "weight_init_factor": 0.01, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 32, "units": 64, "nlayers": 4, "activation": "swish"}
    net = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=0, device=dev)
    m = nif_b200.JacobianLayer(net.build(), [0], [0, 1]).as_model()
    m.compile(nif_b200.Adam(1e-3), loss=nif_b200.SobolevMSE(1e-3, [0], [2]))
    B = int(os.environ.get("NIF_C4_BATCH", "65536"))  # (a multiple of 37 888 rows fills the last wave of the tile-pair kernels)
    X = torch.as_tensor(rng.uniform(-1, 1, (B, 2)).astype(np.float32)).to(dev)
    Y = torch.as_tensor(rng.uniform(-1, 1, (B, 3)).astype(np.float32)).to(dev)
    ms = ev_time(lambda: m._train_step(X, Y, None, B), 5)
    F, P = flops_step(1, 1, 1, 64, 4, 32, 64, 4, tangents=1, tangents_p=0)
    from nif_b200.ops import kernel_profile  # noqa: E402
    m.use_graph = False  # the per-kernel events need eager launches
    ms_eager = ev_time(lambda: m._train_step(X, Y, None, B), 5)
    with kernel_profile() as prof:
        for _ in range(3):
            m._train_step(X, Y, None, B)
    table = [{"kernel": k, "us_per_step": t * 1e3 / 3, "launches_per_step": c / 3} for k, c, t in sorted(prof.table, key=lambda r: -r[2])]
    out.append({"config": f"C4 Sobolev training ShapeNet 4x64, latent 32, batch {B}, loss on u and du/dx (kernels: "
                          f"{net.engine.kernel_path}; tangent forward and both adjoint passes)", "po_dim": P, "ms_per_step": ms,
                "points_per_s": B / ms * 1e3, "ms_per_step_eager_launches": ms_eager, "library_kernels": table,
                "library_kernels_us": sum(r["us_per_step"] for r in table)})

# ---- C5: latent-sweep inference, ShapeNet 3->6x128->1, G latents x N grid points (scaled: 64 x 64^3 per call) ----
if want("C5"):
    cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 3, "output_dim": 1, "units": 128, "nlayers": 6,
             "weight_init_factor": 0.01, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 64, "units": 128, "nlayers": 4, "activation": "swish"}
    net = nif_b200.NIFMultiScale(cfg_s, cfg_p, "mixed_bfloat16", seed=0, device=dev)
    m = net.build()
    G, side = 64, 64
    lin = np.linspace(-1, 1, side, dtype=np.float32)
    grid = torch.as_tensor(np.stack(np.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(-1, 3)).to(dev)
    lat = torch.as_tensor(rng.normal(size=(G, 64)).astype(np.float32)).to(dev)
    ms = ev_time(lambda: m.predict_latent_grid(lat, grid), 3)
    W_s = 3 * 128 + 6 * 128 * 128 + 128
    out.append({"config": f"C5 latent sweep ShapeNet 6x128, {G} latents x {side}^3 grid per call, mixed_bfloat16 (factored form: "
                          f"weights generated once per latent, grouped launches over the shared grid; kernels: {m._eng0.kernel_path})",
                "ms_per_call": ms,
                "evals_per_s": G * grid.shape[0] / ms * 1e3, "algorithmic_tflops": 2 * W_s * G * grid.shape[0] / ms / 1e9})

for o in out:
    print(json.dumps(o))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
