"""Golden vectors for NIFMultiScaleLastLayerParameterized (nif/model.py:989-1269), produced by executing the UNMODIFIED
reference class on the torch-backed TF shim (same mechanism as make_golden.py; build container only):

    python tests/golden/make_golden_lastlayer.py

Writes tests/golden/lastlayer/<case>.npz: cfg dicts, every variable the reference created (reference names), a seeded
input batch, y, phi(x), pnet_output, and the gradients of Keras-'mse' (with sample weights) w.r.t. every variable."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, named_weights  # noqa: E402

CASES = {
    # tutorial/3_multi_scale_linear_NIF.ipynb shape (2 x 30, latent 10), scaled down
    "ll_plain": ({"use_resblock": False, "connectivity": "last_layer", "input_dim": 2, "output_dim": 2, "units": 16,
                  "nlayers": 2, "weight_init_factor": 0.01, "omega_0": 30.0},
                 {"use_resblock": False, "input_dim": 1, "latent_dim": 5, "units": 8, "nlayers": 1, "activation": "swish"}),
    "ll_resblock_sine": ({"use_resblock": True, "connectivity": "last_layer", "input_dim": 1, "output_dim": 1, "units": 12,
                          "nlayers": 1, "weight_init_factor": 0.1, "omega_0": 10.0},
                         {"use_resblock": True, "input_dim": 2, "latent_dim": 3, "units": 6, "nlayers": 1,
                          "activation": "sine", "omega_0": 4.0}),
}


def snet_weights(net):
    out = {}
    for layer in net.snet_list:
        nm = layer.name
        if type(layer).__name__ == "SIREN_ResNet":
            out[nm + "_w"], out[nm + "_b"], out[nm + "_w2"], out[nm + "_b2"] = layer.w, layer.b, layer.w2, layer.b2
        else:
            out[nm + "_w"], out[nm + "_b"] = layer.w, layer.b
    out["last_layer_bias_snet"] = net.last_bias_layer.last_layer_bias
    return out


def main():
    model, _ = import_reference()
    os.makedirs(os.path.join(HERE, "lastlayer"), exist_ok=True)
    for case, (cfg_s, cfg_p) in CASES.items():
        torch.manual_seed(sum(map(ord, case)))
        net = model.NIFMultiScaleLastLayerParameterized(cfg_s, cfg_p, "float64")
        B = 20
        rng = np.random.default_rng(sum(map(ord, case)))
        pi, si, so = cfg_p["input_dim"], cfg_s["input_dim"], cfg_s["output_dim"]
        inputs = torch.as_tensor(rng.uniform(-1, 1, (B, pi + si)))
        target = torch.as_tensor(rng.uniform(-1, 1, (B, so)))
        sw = torch.as_tensor(rng.uniform(0.5, 1.5, (B,)))
        net.call(inputs)  # builds lazy Dense kernels
        W = {**named_weights(net), **snet_weights(net)}
        with torch.no_grad():  # de-correlate the res-block copies so a swapped pair cannot pass unnoticed
            for k, v in W.items():
                if k.endswith("_w2") or k.endswith("_b2"):
                    v.mul_(0.7).add_(0.01)
        for v in W.values():
            v.grad = None
        y = net.call(inputs)
        phi = net._call_shape_net_get_phi_x(inputs[:, pi:pi + si], net.snet_list, so, net.pi_hidden)
        pout = net._call_parameter_net(inputs[:, :pi], net.pnet_list)[0]
        loss = (((y - target) ** 2).mean(-1) * sw).mean()
        loss.backward()
        out = {"cfg_shape_net": json.dumps(cfg_s), "cfg_parameter_net": json.dumps(cfg_p), "inputs": inputs.numpy(),
               "target": target.numpy(), "sample_weight": sw.numpy(), "y": y.detach().numpy(), "phi": phi.detach().numpy(),
               "pnet_output": pout.detach().numpy(), "loss": loss.detach().numpy()}
        for k, v in W.items():
            out["w:" + k] = v.detach().numpy()
            out["g:" + k] = v.grad.numpy()
        path = os.path.join(HERE, "lastlayer", case + ".npz")
        np.savez_compressed(path, **out)
        print(case, "loss", float(loss), os.path.getsize(path), "bytes", sorted(W))


if __name__ == "__main__":
    main()
