"""Shared helpers for the tests (CPU side)."""
import glob
import json
import os

import numpy as np
import torch

from oracle import nif_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(case, dtype=torch.float64):
    d = np.load(os.path.join(GOLDEN, case + ".npz"))
    cfg_s = json.loads(str(d["cfg_shape_net"]))
    cfg_p = json.loads(str(d["cfg_parameter_net"]))
    cls = str(d["cls"])
    spec = O.spec_from_cfg(cls, cfg_s, cfg_p)
    prm = {k[2:]: torch.as_tensor(d[k]).to(dtype) for k in d.files if k.startswith("w:")}
    grads = {k[2:]: torch.as_tensor(d[k]) for k in d.files if k.startswith("g:")}
    return d, cls, cfg_s, cfg_p, spec, prm, grads


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def fullbatch_problem(B=65536):
    """The C2 head (ShapeNet 2 -> 4x64 -> 1 SIREN, latent 32) at the full benchmark batch with seeded inputs:
    (spec, parameters, z [B,32], x [B,2], target [B,1]), all fp32 on the CPU."""
    spec = O.Spec(variant="siren", pi=1, si=2, so=1, n=64, l=4, K=32, n_st=64, l_st=4, p_act="swish", omega0=30.0,
                  weight_init_factor=0.01)
    prm = O.init_params(spec, 0)
    g = torch.Generator().manual_seed(1234)
    z = torch.rand(B, 32, generator=g) - 0.5
    x = torch.rand(B, 2, generator=g) * 2 - 1
    tgt = torch.rand(B, 1, generator=g) * 2 - 1
    return spec, prm, z, x, tgt


C2_CFG_S = {"use_resblock": False, "connectivity": "full", "input_dim": 2, "output_dim": 1, "units": 64, "nlayers": 4,
            "weight_init_factor": 0.01, "omega_0": 30.0}
C2_CFG_P = {"use_resblock": False, "input_dim": 1, "latent_dim": 32, "units": 64, "nlayers": 4, "activation": "swish"}


def fullmodel_problem(B=65536):
    """The whole C2 model (bench.py's configuration: trunk 1 -> 64 x 4 swish -> 32, head, ShapeNet 2 -> 4 x 64 -> 1) at the
    full benchmark batch with seeded inputs: (spec, parameters, inputs [B,3], target [B,1]), fp32 on the CPU."""
    spec = O.spec_from_cfg("NIFMultiScale", C2_CFG_S, C2_CFG_P)
    prm = O.init_params(spec, 0)
    g = torch.Generator().manual_seed(4321)
    inputs = torch.rand(B, 3, generator=g) * 2 - 1
    tgt = torch.rand(B, 1, generator=g) * 2 - 1
    return spec, prm, inputs, tgt
