"""Generate golden vectors by executing the UNMODIFIED reference sources
(/root/reference/nif/model.py, nif/layers/*.py) on top of the torch-backed TF
shim in tests/golden/tf_shim.  Run in the build container only (the GPU box has
no /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/<case>.npz with: the two cfg dicts (JSON), every weight the
reference layers created (named by the reference layer names), a seeded input
batch, the reference output `y`, the pnet_output `(B, po_dim)` tensor, the
latent, the gradients of Keras-'mse' w.r.t. every weight and the latent, and
the JacobianLayer / HessianLayer outputs for the cases that have one, and the latent-code Jacobian that
JacRegLatentLayer regularises.  fp64 is stored; the
fp32 run of the same weights is stored as `y32`.
"""
import importlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/nif"


def import_reference():
    sys.path.insert(0, os.path.join(HERE, "tf_shim"))
    pkg = types.ModuleType("nif")  # skip nif/__init__.py (imports tfp, optimizers)
    pkg.__path__ = [REF]
    sys.modules["nif"] = pkg
    model = importlib.import_module("nif.model")
    grad = importlib.import_module("nif.layers.gradient")
    return model, grad


CASES = {
    # tutorial/1_simple_1d_wave.ipynb:253-265
    "nif_swish_2x30": ("NIF",
        {"connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"},
        {"input_dim": 1, "latent_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}),
    # tutorial/2_multi_scale_NIF.ipynb:567-582
    "siren_2x30": ("NIFMultiScale",
        {"use_resblock": False, "connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30,
         "nlayers": 2, "weight_init_factor": 0.01, "omega_0": 30.0},
        {"use_resblock": False, "input_dim": 1, "latent_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}),
    "siren_si2_n16_K4": ("NIFMultiScale",
        {"use_resblock": False, "connectivity": "full", "input_dim": 2, "output_dim": 1, "units": 16,
         "nlayers": 3, "weight_init_factor": 0.05, "omega_0": 30.0},
        {"use_resblock": True, "input_dim": 1, "latent_dim": 4, "units": 12, "nlayers": 2, "activation": "swish"}),
    "siren_res_so3_n12_K3_sine": ("NIFMultiScale",
        {"use_resblock": True, "connectivity": "full", "input_dim": 2, "output_dim": 3, "units": 12,
         "nlayers": 2, "weight_init_factor": 0.1, "omega_0": 10.0},
        {"use_resblock": True, "input_dim": 2, "latent_dim": 3, "units": 8, "nlayers": 2, "activation": "sine",
         "omega_0": 5.0}),
    "siren_sine_trunk_plain": ("NIFMultiScale",
        {"use_resblock": False, "connectivity": "full", "input_dim": 3, "output_dim": 2, "units": 9,
         "nlayers": 1, "weight_init_factor": 0.1, "omega_0": 30.0},
        {"use_resblock": False, "input_dim": 1, "latent_dim": 2, "units": 7, "nlayers": 1, "activation": "sine",
         "omega_0": 3.0}),
    "nif_tanh_si3_so2": ("NIF",
        {"connectivity": "full", "input_dim": 3, "output_dim": 2, "units": 10, "nlayers": 3, "activation": "tanh"},
        {"input_dim": 2, "latent_dim": 5, "units": 6, "nlayers": 1, "activation": "tanh"}),
}
JAC = {"siren_2x30": ([0], [0, 1]), "nif_tanh_si3_so2": ([0, 1], [0, 2, 4]), "siren_res_so3_n12_K3_sine": ([2], [3])}


def named_weights(net):
    """name -> tensor, using the reference layer names."""
    out = {}
    for layer in net.pnet_list:
        cls = type(layer).__name__
        nm = layer.name
        if cls == "Dense":
            out[nm + "/kernel"], out[nm + "/bias"] = layer.kernel, layer.bias
        elif cls == "MLP_SimpleShortCut":
            out[nm + "/kernel"], out[nm + "/bias"] = layer.L1.kernel, layer.L1.bias
        elif cls == "MLP_ResNet":
            out[nm + "_dense_1/kernel"], out[nm + "_dense_1/bias"] = layer.L1.kernel, layer.L1.bias
            out[nm + "_dense_2/kernel"], out[nm + "_dense_2/bias"] = layer.L2.kernel, layer.L2.bias
        elif cls == "SIREN_ResNet":
            out[nm + "_w"], out[nm + "_b"], out[nm + "_w2"], out[nm + "_b2"] = layer.w, layer.b, layer.w2, layer.b2
        elif cls in ("SIREN", "HyperLinearForSIREN"):
            out[nm + "_w"], out[nm + "_b"] = layer.w, layer.b
        else:
            raise RuntimeError(cls)
    return out


def main():
    model, grad = import_reference()
    for case, (cls, cfg_s, cfg_p) in CASES.items():
        torch.manual_seed(abs(hash(case)) % 1000 if False else sum(map(ord, case)))
        net = getattr(model, cls)(cfg_s, cfg_p, "float64")
        B = 24
        rng = np.random.default_rng(sum(map(ord, case)))
        pi, si, so = cfg_p["input_dim"], cfg_s["input_dim"], cfg_s["output_dim"]
        inputs = torch.as_tensor(rng.uniform(-1, 1, (B, pi + si)))
        target = torch.as_tensor(rng.uniform(-1, 1, (B, so)))
        sw = torch.as_tensor(rng.uniform(0.5, 1.5, (B,)))
        y = net.call(inputs)  # builds lazy Dense kernels
        W = named_weights(net)
        # de-correlate w2/b2 from w/b (the reference initialises them equal) so a
        # swapped pair cannot pass unnoticed
        with torch.no_grad():
            for k, v in W.items():
                if k.endswith("_w2") or k.endswith("_b2"):
                    v.mul_(0.7).add_(0.01)
        for v in W.values():
            v.grad = None
        # forward again keeping the intermediates the oracle is checked against
        p_in = inputs[:, :pi]
        pout, lat = net._call_parameter_net(p_in, net.pnet_list)
        lat.retain_grad()
        pout.retain_grad()
        if cls == "NIF":
            y = net._call_shape_net(inputs[:, pi:pi + si], pout, si, so, net.n_sx, net.l_sx,
                                    cfg_s["activation"], "float64")
        else:
            y = net._call_shape_net_mres(inputs[:, pi:pi + si], pout, cfg_s["use_resblock"],
                                         torch.tensor(cfg_s["omega_0"], dtype=torch.float64),
                                         si, so, net.n_sx, net.l_sx, "float64")
        assert torch.equal(y, net.call(inputs))
        # Keras 'mse' with sample_weight (third-party semantics, restated)
        loss = (((y - target) ** 2).mean(-1) * sw).mean()
        loss.backward()
        out = {
            "cls": cls, "cfg_shape_net": json.dumps(cfg_s), "cfg_parameter_net": json.dumps(cfg_p),
            "inputs": inputs.numpy(), "target": target.numpy(), "sample_weight": sw.numpy(),
            "y": y.detach().numpy(), "pnet_output": pout.detach().numpy(), "latent": lat.detach().numpy(),
            "loss": loss.detach().numpy(), "g_latent": lat.grad.numpy(), "po_dim": net.po_dim,
        }
        for k, v in W.items():
            out["w:" + k] = v.detach().numpy()
            out["g:" + k] = v.grad.numpy()
        # same weights, fp32 policy arithmetic
        with torch.no_grad():
            net32 = getattr(model, cls)(cfg_s, cfg_p, "float32")
            net32.call(inputs.float())
            for (k, v), (k2, v2) in zip(named_weights(net32).items(), W.items()):
                assert k == k2
                v.copy_(v2.float())
            out["y32"] = net32.call(inputs.float()).numpy()
        if case in JAC:
            yi, xi = JAC[case]

            class M:  # JacobianLayer only needs model(x)
                def __call__(self, x):
                    return net.call(x)

            yj, J = grad.compute_output_and_grad(M(), inputs.clone().requires_grad_(True), xi, yi)
            out["jac_y_index"], out["jac_x_index"] = np.array(yi), np.array(xi)
            out["jac"] = J.detach().numpy()
            _, J2, H = grad.compute_output_and_grad_and_hessian(M(), inputs.clone().requires_grad_(True), xi, yi)
            out["hess"] = H.detach().numpy()
        # what JacRegLatentLayer differentiates (nif/model.py:353-375 wires y_index = every latent unit, x_index = every
        # ParameterNet input on the model augmented with the latent code): d latent / d input_p, [B, K, pi]
        class M2:
            def __call__(self, x):
                return net.call(x), net._call_parameter_net(x[:, :pi], net.pnet_list)[1]

        _, dl = grad.compute_output_and_augment_grad(M2(), inputs.clone().requires_grad_(True), list(range(pi)),
                                                     list(range(cfg_p["latent_dim"])))
        out["latent_jac"] = dl.detach().numpy()
        path = os.path.join(HERE, case + ".npz")
        np.savez_compressed(path, **out)
        print(case, "po_dim", net.po_dim, "loss", float(loss), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
