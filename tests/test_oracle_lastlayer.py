"""The oracle's restatement of NIFMultiScaleLastLayerParameterized against vectors produced by the reference class itself
(tests/golden/make_golden_lastlayer.py), and the restated optimiser rules against closed-form single steps."""
import glob
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import nif_oracle as O
from tests.helpers import GOLDEN, rel_err

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "lastlayer", "*.npz")))


def load_lastlayer(case):
    d = np.load(os.path.join(GOLDEN, "lastlayer", case + ".npz"))
    cfg_s, cfg_p = json.loads(str(d["cfg_shape_net"])), json.loads(str(d["cfg_parameter_net"]))
    prm = {k[2:]: torch.as_tensor(d[k]) for k in d.files if k.startswith("w:")}
    grads = {k[2:]: torch.as_tensor(d[k]) for k in d.files if k.startswith("g:")}
    return d, cfg_s, cfg_p, prm, grads


def test_fixtures_exist():
    assert CASES == ["ll_plain", "ll_resblock_sine"]


@pytest.mark.parametrize("case", CASES)
def test_last_layer_forward_and_gradients_match_reference_source(case):
    d, cfg_s, cfg_p, prm, grads = load_lastlayer(case)
    leaves = {k: v.clone().requires_grad_(True) for k, v in prm.items()}
    u, phi, pout = O.last_layer_forward(cfg_s, cfg_p, leaves, torch.as_tensor(d["inputs"]))
    assert rel_err(u.detach(), d["y"]) < 1e-12 and rel_err(phi.detach(), d["phi"]) < 1e-12
    assert rel_err(pout.detach(), d["pnet_output"]) < 1e-12
    loss = (((u - torch.as_tensor(d["target"])) ** 2).mean(-1) * torch.as_tensor(d["sample_weight"])).mean()
    assert abs(float(loss) - float(d["loss"])) < 1e-13
    loss.backward()
    for k, g in grads.items():
        assert rel_err(leaves[k].grad, g) < 1e-10, k


def test_adabelief_first_steps_closed_form():
    """Step 1 of the rectified rule is below the SMA threshold (sma_1 = 1): p -= lr * m_corr = lr * g.  A later step of the
    unrectified rule: p -= lr * m^ / (sqrt(v^) + eps) with v tracking (g - m)^2."""
    p, g = torch.tensor([1.0, -2.0], dtype=torch.float64), torch.tensor([0.5, 0.25], dtype=torch.float64)
    m, v = torch.zeros(2, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
    O.adabelief_step(p, g, m, v, 1, lr=0.1)
    assert torch.allclose(p, torch.tensor([1.0 - 0.05, -2.0 - 0.025], dtype=torch.float64), atol=1e-15)
    p2, m2, v2 = torch.tensor([1.0], dtype=torch.float64), torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64)
    O.adabelief_step(p2, torch.tensor([0.5], dtype=torch.float64), m2, v2, 1, lr=0.1, rectify=False, epsilon=1e-14)
    mt = 0.05
    vt = 0.001 * (0.5 - mt) ** 2 + 1e-14
    want = 1.0 - 0.1 * (mt / 0.1) / (math.sqrt(vt / 0.001) + 1e-14)
    assert abs(float(p2) - want) < 1e-12


def test_adabelief_warmup_rectification_switch_and_amsgrad():
    g = torch.Generator().manual_seed(0)
    p = torch.randn(50, generator=g, dtype=torch.float64)
    m, v, vh = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    norms = []
    for t in range(1, 12):
        before = p.clone()
        O.adabelief_step(p, torch.randn(50, generator=g, dtype=torch.float64), m, v, t, lr=1e-2, amsgrad=True, vhat=vh,
                         total_steps=10, warmup_proportion=0.5, min_lr=1e-4, weight_decay=1e-3)
        norms.append(float((p - before).abs().max()))
        assert torch.all(vh >= v - 1e-30)
    assert all(np.isfinite(norms)) and norms[0] < norms[3]  # warm-up: the first step uses lr / 5


def test_lion_and_centralisation():
    p, g, m = torch.tensor([1.0, -1.0, 0.0]), torch.tensor([0.3, -0.2, 0.0]), torch.tensor([-1.0, 0.1, 0.0])
    O.lion_step(p, g, m, lr=0.1, beta_1=0.9, beta_2=0.99, wd=0.5)
    # sign(0.9 m + 0.1 g) = sign([-0.87, 0.07, 0]) ; p -= 0.1 (sign + 0.5 p)
    assert torch.allclose(p, torch.tensor([1.0 - 0.1 * (-1 + 0.5), -1.0 - 0.1 * (1 - 0.5), 0.0]))
    assert torch.allclose(m, torch.tensor([-0.99 + 0.003, 0.099 - 0.002, 0.0]))
    G = torch.arange(12.0).reshape(3, 4)
    C = O.centralize_gradient(G)
    assert torch.allclose(C.mean(0), torch.zeros(4)) and torch.equal(O.centralize_gradient(G[0]), G[0])
