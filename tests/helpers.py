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
