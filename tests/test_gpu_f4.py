"""SURVEY 8 row f4 and the host-semantics items of the round-1 review, on the GPU through the public API:
NIFMultiScaleLastLayerParameterized, AdaBeliefOptimizer / Lion / gradient centralisation / the L-BFGS fine-tuner, kernel and
activity regularisers, per-epoch metrics."""
import math

import numpy as np
import pytest
import torch

from oracle import nif_oracle as O
from tests.helpers import rel_err
from tests.test_oracle_lastlayer import CASES as LL_CASES, load_lastlayer

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


# ---- NIFMultiScaleLastLayerParameterized ------------------------------------------------------------------------------
@pytest.mark.parametrize("case", LL_CASES)
def test_last_layer_class_matches_reference_golden(case):
    import nif_b200
    d, cfg_s, cfg_p, prm, grads = load_lastlayer(case)
    net = nif_b200.NIFMultiScaleLastLayerParameterized(cfg_s, cfg_p, "float32", seed=0, device=DEV)
    assert sorted(net.variables) == sorted(prm)  # the reference's variable names
    net.set_weights({k: v.numpy() for k, v in prm.items()})
    pi, si = cfg_p["input_dim"], cfg_s["input_dim"]
    X = d["inputs"].astype(np.float32)
    assert rel_err(net.model().predict(X), d["y"]) < 2e-5
    assert rel_err(net.model_x_to_phi().predict(X[:, pi:pi + si]).reshape(d["phi"].shape), d["phi"]) < 2e-5
    assert rel_err(net.model_p_to_lr().predict(X[:, :pi]), d["pnet_output"]) < 2e-5
    u_w = net.model_x_to_u_given_w().predict([X[:, pi:pi + si], d["pnet_output"].astype(np.float32)])
    assert rel_err(u_w, d["y"]) < 2e-5
    with pytest.raises(ValueError):
        net.model_lr_to_w()
    # gradients of Keras 'mse' with sample weights, every variable
    m = net.build()
    m.compile(nif_b200.Adam(1e-3), loss="mse")
    dev = torch.device(DEV)
    loss = m._loss_and_grad(torch.as_tensor(X).to(dev), torch.as_tensor(d["target"]).float().to(dev), "mse",
                            torch.as_tensor(d["sample_weight"]).float().to(dev))
    assert abs(float(loss) - float(d["loss"])) < 1e-5 * float(d["loss"])
    for k, g in grads.items():
        assert rel_err(net._gviews[k].cpu(), g) < 5e-5, k


def test_last_layer_class_trains():
    import nif_b200
    cfg_s = {"use_resblock": False, "connectivity": "last_layer", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2,
             "weight_init_factor": 0.01, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 10, "units": 30, "nlayers": 2, "activation": "swish"}
    net = nif_b200.NIFMultiScaleLastLayerParameterized(cfg_s, cfg_p, seed=0, device=DEV)
    m = net.build()
    m.compile(nif_b200.optimizers.AdaBeliefOptimizer(1e-3), loss="mse")
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (1024, 2)).astype(np.float32)
    Y = (np.sin(4 * X[:, 1:]) * np.exp(-X[:, :1] ** 2)).astype(np.float32)
    h = m.fit(X, Y, batch_size=256, epochs=30)
    assert h.history["loss"][-1] < 0.5 * h.history["loss"][0]


# ---- optimiser kernels against the restated rules ---------------------------------------------------------------------
@pytest.mark.parametrize("rectify,amsgrad,wd,total", [(True, False, 0.0, 0), (False, False, 0.0, 0), (True, True, 1e-2, 0),
                                                      (True, False, 0.0, 40)])
def test_adabelief_kernel_matches_reference_rule(rectify, amsgrad, wd, total):
    from nif_b200.optimizers import AdaBeliefOptimizer
    g = torch.Generator().manual_seed(3)
    n = 1003
    p64 = torch.randn(n, generator=g, dtype=torch.float64)
    m64, v64, vh64 = torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    p = p64.float().to(DEV)
    opt = AdaBeliefOptimizer(1e-2, weight_decay=wd, rectify=rectify, amsgrad=amsgrad, total_steps=total, warmup_proportion=0.25,
                             min_lr=1e-4, epsilon=1e-10)
    for t in range(1, 31):
        gr = torch.randn(n, generator=g, dtype=torch.float64) * (1.0 + 0.1 * t)
        O.adabelief_step(p64, gr, m64, v64, t, lr=1e-2, epsilon=1e-10, weight_decay=wd, rectify=rectify, amsgrad=amsgrad,
                         vhat=vh64, total_steps=total, warmup_proportion=0.25, min_lr=1e-4)
        opt.apply(p, gr.float().to(DEV))
        assert rel_err(p.cpu(), p64) < 2e-6, t
    assert rel_err(opt._m.cpu(), m64) < 2e-6 and rel_err(opt._v.cpu(), v64) < 2e-5


def test_lion_and_centralise_kernels():
    from nif_b200 import _lib
    from nif_b200.optimizers import Lion
    import ctypes as C
    g = torch.Generator().manual_seed(4)
    p64, m64 = torch.randn(777, generator=g, dtype=torch.float64), torch.zeros(777, dtype=torch.float64)
    p = p64.float().to(DEV)
    opt = Lion(1e-3, wd=0.1)
    for t in range(5):
        gr = torch.randn(777, generator=g, dtype=torch.float64)
        O.lion_step(p64, gr, m64, lr=1e-3, wd=0.1)
        opt.apply(p, gr.float().to(DEV))
    assert rel_err(p.cpu(), p64) < 1e-6 and rel_err(opt._m.cpu(), m64) < 1e-6
    G = torch.randn(33, 1000, generator=g)
    Gd = G.to(DEV).contiguous()
    assert _lib.lib().nif_centralize_gradient(33, 1000, C.c_void_p(Gd.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    assert rel_err(Gd.cpu(), O.centralize_gradient(G.double())) < 1e-6


def _tiny_nif(seed=0, **extra_p):
    import nif_b200
    cfg_s = {"connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
    cfg_p = {"input_dim": 1, "latent_dim": 2, "units": 30, "nlayers": 2, "activation": "swish", **extra_p}
    return nif_b200.NIF(cfg_s, cfg_p, seed=seed, device=DEV), cfg_s, cfg_p


def _oracle_grads(spec, prm, X, Y, extra=None):
    leaves = {k: v.clone().requires_grad_(True) for k, v in prm.items()}
    y = O.forward(spec, leaves, X)
    loss = ((y - Y) ** 2).mean(-1).mean()
    if extra is not None:
        loss = loss + extra(leaves)
    loss.backward()
    return float(loss), {k: v.grad for k, v in leaves.items()}


def test_centralised_adabelief_fit_follows_the_restated_rules():
    """tutorial 1's optimiser line (AdaBelief + centralized_gradients_for_optimizer) with centralisation switched on."""
    import nif_b200
    from nif_b200.optimizers import AdaBeliefOptimizer, centralized_gradients_for_optimizer
    net, cfg_s, cfg_p = _tiny_nif()
    spec = O.spec_from_cfg("NIF", cfg_s, cfg_p)
    prm = {k: torch.as_tensor(v).double() for k, v in net.get_weights().items()}
    opt = AdaBeliefOptimizer(1e-3)
    opt.get_gradients = centralized_gradients_for_optimizer(opt, apply_in_fit=True)
    m = net.build()
    m.compile(opt, loss="mse")
    rng = np.random.default_rng(1)
    X = torch.as_tensor(rng.uniform(-1, 1, (300, 2))).double()
    Y = torch.sin(3 * X[:, :1] + X[:, 1:])
    ms = {k: torch.zeros_like(v) for k, v in prm.items()}
    vs = {k: torch.zeros_like(v) for k, v in prm.items()}
    for t in range(1, 5):  # steps below the SMA threshold: p -= lr * m_corr, insensitive to the epsilon floor
        _, gr = _oracle_grads(spec, prm, X, Y)
        for k in prm:
            O.adabelief_step(prm[k], O.centralize_gradient(gr[k]), ms[k], vs[k], t, lr=1e-3)
        m.train_on_batch(X.float(), Y.float())
        for k in prm:
            assert rel_err(net.variables[k].detach().cpu(), prm[k]) < 2e-5, (t, k)
    # without opting in, the assignment is inert (as under TF 2): plain gradients
    net2, _, _ = _tiny_nif()
    opt2 = AdaBeliefOptimizer(1e-3)
    opt2.get_gradients = centralized_gradients_for_optimizer(opt2)
    m2 = net2.build()
    m2.compile(opt2, loss="mse")
    m2.train_on_batch(X.float(), Y.float())
    prm2 = {k: torch.as_tensor(v).double() for k, v in _tiny_nif()[0].get_weights().items()}
    _, gr = _oracle_grads(spec, prm2, X, Y)
    for k in prm2:
        O.adabelief_step(prm2[k], gr[k], torch.zeros_like(prm2[k]), torch.zeros_like(prm2[k]), 1, lr=1e-3)
        assert rel_err(net2.variables[k].detach().cpu(), prm2[k]) < 2e-5, k


def test_lbfgs_fine_tuner_reduces_the_loss():
    from nif_b200.optimizers import TFPLBFGS
    net, cfg_s, cfg_p = _tiny_nif()
    m = net.build()
    rng = np.random.default_rng(2)
    X = rng.uniform(-1, 1, (512, 2)).astype(np.float32)
    Y = np.sin(3 * X[:, :1] + X[:, 1:]).astype(np.float32)

    def loss_fun(y_true, y_pred):
        return ((y_true - y_pred) ** 2).mean()

    ft = TFPLBFGS(m, loss_fun, X, Y, display_epoch=0)
    ft.minimize(rounds=2, max_iter=25)
    h = ft.history["loss"]
    assert len(h) > 10 and h[-1] < 0.2 * h[0] and np.isfinite(h).all()
    assert abs(float(m.evaluate(X, Y)) - min(h)) < 1e-3 * max(min(h), 1e-6) + 1e-6  # the model holds the best iterate


# ---- regularisers -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key,val", [("l2_reg", 1e-3), ("l1_reg", 1e-4)])
def test_kernel_regularisers_enter_the_loss_and_the_update(key, val):
    import nif_b200
    net, cfg_s, cfg_p = _tiny_nif(**{key: val})
    spec = O.spec_from_cfg("NIF", cfg_s, cfg_p)
    prm = {k: torch.as_tensor(v).double() for k, v in net.get_weights().items()}
    rng = np.random.default_rng(5)
    X = torch.as_tensor(rng.uniform(-1, 1, (400, 2))).double()
    Y = torch.sin(3 * X[:, :1] + X[:, 1:])

    def reg(leaves):  # every ParameterNet kernel and bias, last layer included (nif/model.py:107-117, 220-230)
        return sum(val * ((v * v).sum() if key == "l2_reg" else v.abs().sum()) for v in leaves.values())

    m = net.build()
    m.compile(nif_b200.Adam(1e-3), loss="mse", graph=False)
    ms = {k: torch.zeros_like(v) for k, v in prm.items()}
    vs = {k: torch.zeros_like(v) for k, v in prm.items()}
    for t in range(1, 4):
        loss64, gr = _oracle_grads(spec, prm, X, Y, reg)
        for k in prm:
            O.adam_tf(prm[k], gr[k], ms[k], vs[k], t, 1e-3)
        got = m.train_on_batch(X.float(), Y.float())
        assert abs(got - loss64) < 2e-5 * abs(loss64), (t, got, loss64)
    for k in prm:
        assert rel_err(net.variables[k].detach().cpu(), prm[k]) < 1e-4, k


@pytest.mark.parametrize("key,val,fused", [("act_l2_reg", 1e-4, True), ("act_l1_reg", 1e-5, True), ("act_l2_reg", 1e-4, False)])
def test_activity_regulariser_of_the_last_parameter_net_layer(key, val, fused):
    """activity_regularizer on pnet_output = z W_h + b_h (nif/model.py:118-125, 229): Keras adds reg(output) / batch."""
    import nif_b200
    net, cfg_s, cfg_p = _tiny_nif(**{key: val})
    spec = O.spec_from_cfg("NIF", cfg_s, cfg_p)
    prm = {k: torch.as_tensor(v).double() for k, v in net.get_weights().items()}
    rng = np.random.default_rng(6)
    X = torch.as_tensor(rng.uniform(-1, 1, (333, 2))).double()
    Y = torch.sin(3 * X[:, :1] + X[:, 1:])
    wn, bn = O.last_layer_names(spec)

    def reg(leaves):
        pout = O.hyper_linear(O.latent(spec, leaves, X[:, :1]), leaves[wn], leaves[bn])
        return val * ((pout * pout).sum() if key == "act_l2_reg" else pout.abs().sum()) / X.shape[0]

    loss64, gr = _oracle_grads(spec, prm, X, Y, reg)
    m = net.build()
    m.compile(nif_b200.Adam(0.0), loss="mse" if fused else (lambda yt, yp: ((yt - yp) ** 2).mean(-1).mean()), graph=False)
    got = m.train_on_batch(X.float(), Y.float())  # lr = 0: the gradient buffer is all that changes
    assert abs(got - loss64) < 2e-5 * abs(loss64)
    for k in prm:
        assert rel_err(net._gviews[k].cpu(), gr[k]) < 5e-5, k


def test_fit_reports_callable_metrics():
    import nif_b200
    net, _, _ = _tiny_nif()
    m = net.build()

    def mae(y_true, y_pred):
        return (y_true - y_pred).abs().mean()

    m.compile(nif_b200.Adam(1e-3), loss="mse", metrics=[mae])
    rng = np.random.default_rng(7)
    X = rng.uniform(-1, 1, (256, 2)).astype(np.float32)
    Y = np.sin(3 * X[:, :1] + X[:, 1:]).astype(np.float32)
    h = m.fit(X, Y, batch_size=64, epochs=3)
    assert len(h.history["mae"]) == 3 and h.history["mae"][-1] <= h.history["mae"][0]
