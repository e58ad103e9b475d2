"""Pin the oracle: it must reproduce what the reference's own source computes
(golden vectors made by tests/golden/make_golden.py) and the structural known
answers stored in the reference notebooks."""
import numpy as np
import pytest
import torch

from oracle import nif_oracle as O
from tests.helpers import golden_cases, load_golden, rel_err


@pytest.mark.parametrize("case", golden_cases())
def test_forward_matches_reference_source(case):
    d, cls, cfg_s, cfg_p, spec, prm, _ = load_golden(case)
    assert spec.po_dim == int(d["po_dim"])
    assert sorted(prm) == sorted(O.trunk_param_names(spec))
    x = torch.as_tensor(d["inputs"])
    z = O.latent(spec, prm, x[:, : spec.pi])
    assert rel_err(z, d["latent"]) < 1e-13
    wn, bn = O.last_layer_names(spec)
    p = O.hyper_linear(z, prm[wn], prm[bn])
    assert rel_err(p, d["pnet_output"]) < 1e-13
    y = O.forward(spec, prm, x)
    assert rel_err(y, d["y"]) < 1e-12


@pytest.mark.parametrize("case", golden_cases())
def test_fp32_forward_matches_reference_source(case):
    d, cls, cfg_s, cfg_p, spec, prm, _ = load_golden(case, torch.float32)
    y = O.forward(spec, prm, torch.as_tensor(d["inputs"]).float())
    # same ops, same order, same library: fp32 results agree to rounding
    assert rel_err(y, d["y32"]) < 2e-6
    assert rel_err(d["y32"], d["y"]) < 5e-4


@pytest.mark.parametrize("case", golden_cases())
def test_gradients_match_reference_source(case):
    d, cls, cfg_s, cfg_p, spec, prm, grads = load_golden(case)
    loss, g, gz, y = O.loss_and_grads(
        spec, prm, torch.as_tensor(d["inputs"]), torch.as_tensor(d["target"]), torch.as_tensor(d["sample_weight"])
    )
    assert abs(float(loss) - float(d["loss"])) < 1e-13
    assert rel_err(gz, d["g_latent"]) < 1e-11
    for k in grads:
        assert rel_err(g[k], grads[k]) < 1e-11, k


@pytest.mark.parametrize("case", [c for c in golden_cases() if "jac" in np.load(f"tests/golden/{c}.npz").files]
                         if True else [])
def test_jacobian_hessian_match_reference_source(case):
    d, cls, cfg_s, cfg_p, spec, prm, _ = load_golden(case)
    yi, xi = list(d["jac_y_index"]), list(d["jac_x_index"])
    y, J = O.jacobian(spec, prm, torch.as_tensor(d["inputs"]), yi, xi)
    assert J.shape == (24, len(yi), len(xi))
    assert rel_err(J, d["jac"]) < 1e-11
    _, J2, H = O.hessian(spec, prm, torch.as_tensor(d["inputs"]), yi, xi)
    assert H.shape == (24, len(yi), len(xi), len(xi))
    assert rel_err(H, d["hess"]) < 1e-10


@pytest.mark.parametrize("case", golden_cases())
def test_jac_reg_loss_matches_reference_source(case):
    """JacRegLatentLayer's add_loss term: the golden `latent_jac` is what the reference's
    compute_output_and_augment_grad returns for the model augmented with the latent code (gradient.py:183-204 as wired
    by model.py:353-375); the layer then adds l1 * reduce_mean(square(.)) (gradient.py:101-104).  The oracle, and the
    product's forward-mode formulation of the same term (trunk only, runs on CPU), are pinned against it."""
    import nif_b200
    d, cls, cfg_s, cfg_p, spec, prm, _ = load_golden(case)
    l1 = 0.37
    want = l1 * float(np.mean(np.square(d["latent_jac"])))
    got = float(O.jac_reg_loss(spec, prm, torch.as_tensor(d["inputs"]), l1))
    assert abs(got - want) <= 1e-12 * max(1.0, abs(want))
    net = getattr(nif_b200, cls)(cfg_s, dict(cfg_p, jac_reg=l1), seed=0, device="cpu")
    net.set_weights({k: v.float().numpy() for k, v in prm.items()})
    prod = float(net._jac_reg_loss(torch.as_tensor(d["inputs"][:, : spec.pi]).float()))
    assert abs(prod - want) <= 2e-5 * abs(want) + 1e-12


def test_notebook_known_answers():
    # tutorial/1_simple_1d_wave.ipynb cells 27/31/33: po_dim 1951; p->lr 1951 params; lr->w 3902 params
    cfg_s = {"connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
    cfg_p = {"input_dim": 1, "latent_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
    spec = O.spec_from_cfg("NIF", cfg_s, cfg_p)
    assert spec.po_dim == 1951
    prm = O.init_params(spec, 0)
    wn, bn = O.last_layer_names(spec)
    n_last = prm[wn].numel() + prm[bn].numel()
    n_all = sum(v.numel() for v in prm.values())
    assert n_last == 3902
    assert n_all - n_last == 1951
    assert n_all == 5853
    # SURVEY section 8 config sizes
    assert O.po_dim(2, 1, 64, 4, False) == 16897
    assert O.po_dim(3, 3, 128, 6, False) == 99971
    assert O.po_dim(1, 1, 64, 4, False) == 16833
    assert O.po_dim(3, 1, 128, 6, False) == 99713


def test_layout_is_a_partition():
    for (si, so, n, l, res) in [(1, 1, 30, 2, False), (2, 3, 12, 2, True), (3, 2, 7, 0, False)]:
        L = O.layout(si, so, n, l, res)
        cover = np.zeros(L.P, int)
        for (o, a, c) in L.w:
            cover[o : o + a * c] += 1
        for (o, c) in L.b:
            cover[o : o + c] += 1
        assert (cover == 1).all()


def test_gradient_finite_difference():
    d, cls, cfg_s, cfg_p, spec, prm, _ = load_golden("siren_si2_n16_K4")
    x, t = torch.as_tensor(d["inputs"]), torch.as_tensor(d["target"])
    loss, g, gz, _ = O.loss_and_grads(spec, prm, x, t)
    wn, bn = O.last_layer_names(spec)
    rng = np.random.default_rng(0)
    for name in (wn, bn):
        for _ in range(5):
            idx = tuple(int(rng.integers(0, s)) for s in prm[name].shape)
            eps = 1e-6
            pp = {k: v.clone() for k, v in prm.items()}
            pp[name][idx] += eps
            lp = O.mse(O.forward(spec, pp, x), t)
            pp[name][idx] -= 2 * eps
            lm = O.mse(O.forward(spec, pp, x), t)
            fd = float(lp - lm) / (2 * eps)
            assert abs(fd - float(g[name][idx])) < 1e-6 * max(1.0, abs(fd))


def test_init_bounds_follow_reference_rules():
    # siren.py:36-62: w ~ U(+-sqrt(6/K) f); b bounds per column block
    cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 2, "output_dim": 1, "units": 16,
             "nlayers": 3, "weight_init_factor": 0.05, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 4, "units": 12, "nlayers": 2, "activation": "swish"}
    spec = O.spec_from_cfg("NIFMultiScale", cfg_s, cfg_p)
    wb, bb = O.hyper_init_bounds(spec)
    assert abs(wb - np.sqrt(6 / 4) * 0.05) < 1e-15
    L = O.layout(2, 1, 16, 3, False)
    assert np.allclose(bb[: 32], 0.5)
    assert np.allclose(bb[32 : 32 + 3 * 256], np.sqrt(6 / 16) / 30)
    assert np.allclose(bb[L.w[-1][0] : L.w[-1][0] + 16], np.sqrt(6 / 32))
    assert np.allclose(bb[L.b[0][0] :], 1 / 16)
    prm = O.init_params(spec, 3)
    assert float(prm["HyperLinearForSIREN_w"].abs().max()) <= wb
    assert bool((prm["HyperLinearForSIREN_b"].abs().double() <= torch.as_tensor(bb) * (1 + 1e-6)).all())
    assert float(prm["mlp_first_pnet/kernel"].abs().max()) <= 0.2 + 1e-7  # truncated at 2 sigma


def test_traveling_wave_regenerates():
    raw = O.traveling_wave_raw(4.0)
    assert raw.shape == (2000, 3) and raw.dtype == np.float32
    nd, mean, std = O.standard_normalize(raw)
    assert np.allclose(nd.mean(0), 0, atol=1e-5) and np.allclose(nd.std(0), 1, atol=1e-5)
    hf, m2, s2 = O.minmax_normalize(O.traveling_wave_raw(400.0), 1, 1, 1)
    assert abs(hf[:, 0].min() + 1) < 1e-6 and abs(hf[:, 1].max() - 1) < 1e-6
    assert abs(np.abs(hf[:, 2]).max() - 1) < 1e-6


def test_adam_tf_semantics():
    p = torch.tensor([1.0, -2.0], dtype=torch.float64)
    g = torch.tensor([0.5, 0.25], dtype=torch.float64)
    m, v = torch.zeros(2, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
    O.adam_tf(p, g, m, v, 1, 1e-3)
    # first step: m = 0.1 g, v = 0.001 g^2, alpha = lr*sqrt(0.001)/0.1
    exp = torch.tensor([1.0, -2.0], dtype=torch.float64) - 1e-3 * np.sqrt(0.001) / 0.1 * (0.1 * g) / ((0.001 * g * g).sqrt() + 1e-7)
    assert torch.allclose(p, exp, atol=1e-15)
