"""Host-side mirror of the reference API (no GPU needed): configuration, variable names and shapes,
initialisers, trunk arithmetic, checkpoints, dataset semantics, error behaviour."""
import json

import numpy as np
import pytest
import torch

import nif_b200
from oracle import nif_oracle as O
from tests.helpers import golden_cases, load_golden, rel_err

CFG_S = {"use_resblock": False, "connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2,
         "weight_init_factor": 0.01, "omega_0": 30.0}
CFG_P = {"use_resblock": False, "input_dim": 1, "latent_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}


def test_known_answers_from_notebooks():
    net = nif_b200.NIFMultiScale(CFG_S, CFG_P, seed=0, device="cpu")
    assert net.po_dim == 1951
    assert net.model_p_to_lr().count_params() == 1951  # tutorial/1 cell 31
    assert net.model_lr_to_w().count_params() == 3902  # tutorial/1 cell 33
    assert net.build().count_params() == 5853
    nif = nif_b200.NIF({"connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2,
                        "activation": "swish"},
                       {"input_dim": 1, "latent_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}, seed=0,
                       device="cpu")
    assert nif.po_dim == 1951 and nif.count_params() == 5853


@pytest.mark.parametrize("case", golden_cases())
def test_variable_names_shapes_and_trunk_match_reference(case):
    d, cls, cfg_s, cfg_p, spec, prm, _ = load_golden(case, torch.float32)
    net = getattr(nif_b200, cls)(cfg_s, cfg_p, seed=1, device="cpu")
    assert list(net.variables) == O.trunk_param_names(spec)
    for k, v in net.variables.items():
        assert tuple(v.shape) == tuple(prm[k].shape), k
    net.set_weights({k: v.numpy() for k, v in prm.items()})
    x = torch.as_tensor(d["inputs"]).float()
    with torch.no_grad():
        z = net._latent(x[:, : net.pi_dim])
    assert rel_err(z, d["latent"]) < 1e-5
    # model_lr_to_w is `z @ w + b`: compare with the reference's pnet_output
    with torch.no_grad():
        p = torch.addmm(net.b_h, z, net.w_h)
    assert rel_err(p, d["pnet_output"]) < 1e-5


def test_initialisers_follow_reference_distributions():
    cfg_s = dict(CFG_S, units=16, input_dim=2, nlayers=3, weight_init_factor=0.05)
    cfg_p = dict(CFG_P, latent_dim=4, units=12)
    net = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=3, device="cpu")
    spec = O.spec_from_cfg("NIFMultiScale", cfg_s, cfg_p)
    wb, bb = O.hyper_init_bounds(spec)
    w, b = net.w_h.detach(), net.b_h.detach()
    assert float(w.abs().max()) <= wb and float(w.abs().max()) > 0.9 * wb
    assert bool((b.abs().double() <= torch.as_tensor(bb) * (1 + 1e-6)).all())
    assert abs(float(w.mean())) < 0.05 * wb
    k = net.variables["mlp_first_pnet/kernel"].detach()
    assert float(k.abs().max()) <= 0.2 + 1e-6  # TruncatedNormal(0.1): cut at 2 sigma
    # seeds are reproducible and distinct
    n2 = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=3, device="cpu")
    n3 = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=4, device="cpu")
    assert torch.equal(n2.theta, net.theta) and not torch.equal(n3.theta, net.theta)


def test_sine_trunk_resblock_copies_w_into_w2():
    cfg_p = {"use_resblock": True, "input_dim": 2, "latent_dim": 3, "units": 8, "nlayers": 2, "activation": "sine",
             "omega_0": 5.0}
    net = nif_b200.NIFMultiScale(dict(CFG_S, use_resblock=True), cfg_p, seed=0, device="cpu")
    v = net.variables
    assert torch.equal(v["siren_hidden_resblock_pnet_0_w"], v["siren_hidden_resblock_pnet_0_w2"])  # siren.py:370-379
    assert torch.equal(v["siren_hidden_resblock_pnet_1_b"], v["siren_hidden_resblock_pnet_1_b2"])


def test_cfg_validation_errors_match_reference():
    with pytest.raises(TypeError):
        nif_b200.NIFMultiScale(CFG_S, "nope", device="cpu")
    with pytest.raises(TypeError):
        nif_b200.NIFMultiScale([1], CFG_P, device="cpu")
    bad = dict(CFG_S)
    del bad["use_resblock"]
    with pytest.raises(AssertionError):
        nif_b200.NIFMultiScale(bad, CFG_P, device="cpu")
    with pytest.raises(AssertionError):
        nif_b200.NIFMultiScale(dict(CFG_S, use_resblock=1), CFG_P, device="cpu")
    with pytest.raises(ValueError):
        nif_b200.NIFMultiScale(dict(CFG_S, connectivity="banana"), CFG_P, device="cpu")
    with pytest.raises(ValueError):  # 'last_layer' belongs to NIFMultiScaleLastLayerParameterized
        nif_b200.NIFMultiScale(dict(CFG_S, connectivity="last_layer"), CFG_P, device="cpu")
    with pytest.raises(AssertionError):  # nif/model.py:1020-1022
        nif_b200.NIFMultiScaleLastLayerParameterized(CFG_S, CFG_P, device="cpu")
    ll = nif_b200.NIFMultiScaleLastLayerParameterized(dict(CFG_S, connectivity="last_layer"), CFG_P, device="cpu")
    assert ll.po_dim == CFG_P["latent_dim"]  # nif/model.py:583-585
    with pytest.raises(ValueError):
        nif_b200.NIFMultiScale(CFG_S, CFG_P, "float64", device="cpu")


def test_save_config_roundtrip(tmp_path):
    net = nif_b200.NIFMultiScale(CFG_S, CFG_P, "float32", seed=0, device="cpu")
    f = tmp_path / "config.json"
    net.save_config(str(f))
    c = json.load(open(f))
    assert c == {"cfg_shape_net": CFG_S, "cfg_parameter_net": CFG_P, "mixed_policy": "float32"}
    again = nif_b200.NIFMultiScale(c["cfg_shape_net"], c["cfg_parameter_net"], c["mixed_policy"], device="cpu")
    assert again.po_dim == net.po_dim


def test_weights_checkpoint_roundtrip(tmp_path):
    a = nif_b200.NIFMultiScale(CFG_S, CFG_P, seed=0, device="cpu")
    b = nif_b200.NIFMultiScale(CFG_S, CFG_P, seed=1, device="cpu")
    a.build().save_weights(str(tmp_path / "ckpt-0" / "ckpt"))
    b.build().load_weights(str(tmp_path / "ckpt-0" / "ckpt"))
    assert torch.equal(a.theta, b.theta)
    with pytest.raises(KeyError):
        b.set_weights({"nope": np.zeros(1)})
    with pytest.raises(ValueError):
        b.set_weights({"bottleneck_pnet/bias": np.zeros(7)})


def test_derived_models_share_parameters():
    net = nif_b200.NIFMultiScale(CFG_S, CFG_P, seed=0, device="cpu")
    m1, m2 = net.build(), net.model_lr_to_w()
    assert m1.net is m2.net
    assert [n for n, _ in m1.inputs] == ["input_tot"] and m1.inputs[0][1] == 2
    assert [n for n, _ in net.model_x_to_u_given_w().inputs] == ["input_x_to_u_given_w", "input_w_and_b_from_pnet"]
    assert net.model_x_to_u_given_w().inputs[1][1] == 1951


def test_no_cpu_fallback():
    net = nif_b200.NIFMultiScale(CFG_S, CFG_P, seed=0, device="cpu")
    with pytest.raises(nif_b200._lib.NifError):
        net.build().predict(np.zeros((4, 2), np.float32))
    m = net.build()
    m.compile(nif_b200.Adam(1e-3), loss="mse")
    with pytest.raises(nif_b200._lib.NifError):
        m.fit(np.zeros((4, 2), np.float32), np.zeros((4, 1), np.float32), batch_size=2, epochs=1)


def test_dataset_semantics():
    x = np.arange(10, dtype=np.float32).reshape(10, 1)
    y = -x
    ds = nif_b200.Dataset.from_tensor_slices((x, y)).shuffle(10, seed=7).batch(4).prefetch(1)
    e0 = [(gb, r[0].numpy().ravel().copy()) for gb, r in ds.batches(0)]
    e1 = [(gb, r[0].numpy().ravel().copy()) for gb, r in ds.batches(1)]
    assert [g for g, _ in e0] == [4, 4, 2]  # short last batch is kept
    assert sorted(np.concatenate([r for _, r in e0]).tolist()) == list(range(10))
    assert not np.array_equal(np.concatenate([r for _, r in e0]), np.concatenate([r for _, r in e1]))  # reshuffled
    again = [(gb, r[0].numpy().ravel().copy()) for gb, r in ds.batches(0)]
    assert all(np.array_equal(a[1], b[1]) for a, b in zip(e0, again))  # deterministic per (seed, epoch)
    for gb, r in ds.batches(0):
        assert np.array_equal(r[1].numpy(), -r[0].numpy())
    # data-parallel sharding: the two ranks partition every global batch
    d0 = nif_b200.Dataset.from_tensor_slices((x, y)).shuffle(10, seed=7).batch(4).shard(2, 0)
    d1 = nif_b200.Dataset.from_tensor_slices((x, y)).shuffle(10, seed=7).batch(4).shard(2, 1)
    for (gb, full), (_, a), (_, b) in zip(e0, d0.batches(0), d1.batches(0)):
        assert sorted(np.concatenate([a[0].numpy().ravel(), b[0].numpy().ravel()]).tolist()) == sorted(full.tolist())


def test_dataset_cache_on_device_yields_the_same_batches():
    """Dataset.cache_on(device) (what fit() does for data sets that fit in HBM): same batches in the same order; exercised
    here with the CPU as the 'device' -- the code path is device-agnostic."""
    import torch
    x = np.arange(40, dtype=np.float32).reshape(20, 2)
    y = x.sum(1, keepdims=True)
    a = nif_b200.Dataset.from_tensor_slices((x, y)).shuffle(20, seed=3).batch(6)
    b = nif_b200.Dataset.from_tensor_slices((x, y)).shuffle(20, seed=3).batch(6).cache_on(torch.device("cpu"))
    assert b.nbytes() == x.nbytes + y.nbytes
    for epoch in range(2):
        for (ga, ra), (gb, rb) in zip(a.batches(epoch), b.batches(epoch)):
            assert ga == gb and all(torch.equal(p, q) for p, q in zip(ra, rb))


def test_lr_scheduler_and_adam_object():
    opt = nif_b200.Adam(1e-3)
    assert (opt.beta_1, opt.beta_2, opt.epsilon) == (0.9, 0.999, 1e-7)  # tf.keras defaults

    class M:
        optimizer = opt
    cb = nif_b200.LearningRateScheduler(lambda epoch, lr: lr if epoch < 2 else 1e-4)
    cb.set_model(M())
    cb.on_epoch_begin(0)
    assert opt.learning_rate == 1e-3
    cb.on_epoch_begin(2)
    assert opt.learning_rate == 1e-4


def test_pointwise_normalisers_match_oracle():
    from nif_b200.data import PointWiseData
    from nif_b200.demo import TravelingWave, TravelingWaveHighFreq
    raw = O.traveling_wave_raw(4.0)
    a, m, s = PointWiseData.standard_normalize(raw.copy())
    b, m2, s2 = O.standard_normalize(raw.copy())
    assert np.array_equal(a, b)
    hf = O.traveling_wave_raw(400.0)
    a, m, s = PointWiseData.minmax_normalize(hf.copy(), 1, 1, 1)
    b, m2, s2 = O.minmax_normalize(hf.copy(), 1, 1, 1)
    assert np.allclose(a, b, atol=1e-7) and np.allclose(m, m2) and np.allclose(s, s2)
    assert TravelingWave().data.shape == (2000, 3) and TravelingWaveHighFreq().u.shape == (2000, 1)
    # area-weighted variant: last column divided by its mean, not centred
    rng = np.random.default_rng(0)
    rawa = np.hstack([rng.normal(size=(50, 3)), rng.uniform(1, 2, (50, 1))])
    d, mean, std, w = PointWiseData.minmax_normalize(rawa.copy(), 1, 1, 1, area_weighted=True)
    assert d.shape == (50, 3) and abs(w.mean() - 1) < 1e-12 and mean[-1] == 0


@pytest.mark.parametrize("cls,cfg_s,cfg_p", [
    ("NIF", {"connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"},
     {"input_dim": 2, "latent_dim": 3, "units": 30, "nlayers": 2, "activation": "swish", "jac_reg": 1e-2}),
    ("NIFMultiScale", CFG_S, dict(CFG_P, jac_reg=0.5)),
])
def test_jac_reg_latent_loss_matches_oracle(cls, cfg_s, cfg_p):
    """The add_loss term of JacRegLatentLayer (gradient.py:52-113, model.py:353-375): value and trunk gradient of the
    product's forward-mode formulation against the oracle's batch-Jacobian restatement (trunk only: runs on CPU)."""
    import torch
    from oracle import nif_oracle as O
    spec = O.spec_from_cfg(cls, cfg_s, cfg_p)
    prm = {k: v.double() for k, v in O.init_params(spec, 5).items()}
    net = getattr(nif_b200, cls)(cfg_s, cfg_p, seed=0, device="cpu")
    assert net._trunk is None and not net._fused_trunk_supported()  # the regulariser needs the differentiable trunk
    net.set_weights({k: v.float().numpy() for k, v in prm.items()})
    rng = np.random.default_rng(0)
    X = torch.as_tensor(rng.uniform(-1, 1, (64, spec.pi + spec.si)))
    ref_prm = {k: v.clone().requires_grad_(True) for k, v in prm.items()}
    ref = O.jac_reg_loss(spec, ref_prm, X, cfg_p["jac_reg"])
    ref.backward()
    got = net._jac_reg_loss(X[:, : spec.pi].float())
    net.grad.zero_()
    got.backward()
    assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref))
    last = set(net._last_names)
    for k, v in ref_prm.items():
        g = net._gviews[k]
        if k in last or v.grad is None:
            # the last linear layer is not part of the latent code; the bottleneck bias does not enter its Jacobian
            assert v.grad is None and float(g.abs().max()) == 0.0, k
        else:
            assert float((g.double() - v.grad).abs().max()) <= 2e-5 * float(v.grad.abs().max()) + 1e-12, k


def test_pointwise_data_matches_the_reference_file():
    """PointWiseData against vectors produced by the unmodified nif/data/point_wise_data.py
    (tests/golden/make_golden_pointwise.py)."""
    import os
    from nif_b200.data import PointWiseData as P
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pointwise", "pointwise.npz"))
    raw, rawa = g["raw"], g["rawa"]
    d, m, s = P.standard_normalize(raw.copy())
    assert np.allclose(d, g["std_data"], atol=1e-12) and np.allclose(m, g["std_mean"]) and np.allclose(s, g["std_std"])
    d, m, s, w = P.standard_normalize(rawa.copy(), area_weighted=True)
    assert np.allclose(d, g["stda_data"], atol=1e-12) and np.allclose(m, g["stda_mean"]) and np.allclose(s, g["stda_std"])
    assert np.allclose(w, g["stda_w"])
    d, m, s = P.minmax_normalize(raw.copy(), 2, 3, 2)
    assert np.allclose(d, g["mm_data"], atol=1e-12) and np.allclose(m, g["mm_mean"]) and np.allclose(s, g["mm_std"])
    d, m, s, w = P.minmax_normalize(rawa.copy(), 2, 3, 2, area_weighted=True)
    assert np.allclose(d, g["mma_data"], atol=1e-12) and np.allclose(m, g["mma_mean"]) and np.allclose(s, g["mma_std"])
    assert np.allclose(w, g["mma_w"])
    obj = P(raw[:, :2], raw[:, 2:5], raw[:, 5:7], rawa[:, -1:])
    obj.data = obj.data_raw
    assert np.array_equal(obj.data_raw, g["obj_raw"]) and np.array_equal(obj.parameter, g["obj_parameter"])
    assert np.array_equal(obj.x, g["obj_x"]) and np.array_equal(obj.u, g["obj_u"])


def test_jacobian_model_loss_planning():
    """compile() of a JacobianLayer model maps the loss onto tangent directions: one per differentiated input column, in the
    order the grad columns name them; ParameterNet inputs move the latent code; 'mse' / a callable see every Jacobian entry;
    more than four differentiated inputs, or columns outside the output, are rejected (host logic, no GPU)."""
    cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 2, "output_dim": 2, "units": 30, "nlayers": 2,
             "weight_init_factor": 0.01, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 2, "units": 30, "nlayers": 2, "activation": "swish"}
    net = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=0, device="cpu")
    # output columns: [u0, u1 | du0/dt, du0/dx, du0/dy, du1/dt, du1/dx, du1/dy]
    model = nif_b200.JacobianLayer(net.build(), y_index=[0, 1], x_index=[0, 1, 2]).as_model()
    model.compile(nif_b200.Adam(1e-3), loss=nif_b200.SobolevMSE(0.1, value_cols=[0, 1], grad_cols=[3, 7]))
    plan = model._sobolev_plan
    assert plan["dirs"] == [1, 2] and plan["pairs"] == [(0, 0), (1, 1)] and not plan["latent_moves"] and not plan["general"]
    model.compile(nif_b200.Adam(1e-3), loss=nif_b200.SobolevMSE(0.1, value_cols=[0], grad_cols=[5, 2, 4]))
    plan = model._sobolev_plan
    # columns 5, 2, 4 = du1/dt, du0/dt, du0/dy: directions t (a ParameterNet input) and y
    assert plan["dirs"] == [0, 2] and plan["pairs"] == [(0, 1), (0, 0), (1, 0)] and plan["latent_moves"]
    for loss in ("mse", lambda yt, yp: ((yt - yp) ** 2).mean()):
        model.compile(nif_b200.Adam(1e-3), loss=loss)
        plan = model._sobolev_plan
        assert plan["general"] and plan["dirs"] == [0, 1, 2] and plan["latent_moves"]
        assert plan["pairs"] == [(0, 0), (1, 0), (2, 0), (0, 1), (1, 1), (2, 1)]
    with pytest.raises(nif_b200._lib.NifError):
        model.compile(nif_b200.Adam(1e-3), loss=nif_b200.SobolevMSE(0.1, value_cols=[0], grad_cols=[8]))
    with pytest.raises(nif_b200._lib.NifError):
        model.compile(nif_b200.Adam(1e-3), loss=nif_b200.SobolevMSE(0.1, value_cols=[2], grad_cols=[3]))
    cfg_s5 = dict(cfg_s, input_dim=4)
    net5 = nif_b200.NIFMultiScale(cfg_s5, cfg_p, seed=0, device="cpu")
    wide = nif_b200.JacobianLayer(net5.build(), y_index=[0], x_index=[0, 1, 2, 3, 4]).as_model()
    with pytest.raises(nif_b200._lib.NifError):
        wide.compile(nif_b200.Adam(1e-3), loss="mse")  # five differentiated inputs
