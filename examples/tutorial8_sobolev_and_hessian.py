"""tutorial/8_NIF_with_Sobolov_training.ipynb on nif_b200, plus HessianLayer (needs a B200).

Sobolev training wraps the model in JacobianLayer and puts du/dx into the loss; here that trains through the
reverse-over-forward kernels.  HessianLayer (README.md:119-144 of the reference, PDE residuals) runs second-order forward mode."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nif_b200 as nif  # noqa: E402

cfg_shape_net = {"use_resblock": False, "connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 64, "nlayers": 4,
                 "weight_init_factor": 0.01, "omega_0": 30.0}
cfg_parameter_net = {"use_resblock": False, "input_dim": 1, "latent_dim": 32, "units": 64, "nlayers": 4, "activation": "swish"}

# travelling wave with omega = 400 and its derivatives (tutorial 8 cells 4, 6), inputs scaled to [-1, 1]
rng = np.random.default_rng(0)
n = 200_000
t = rng.uniform(0, 100, (n, 1))
x = rng.uniform(0, 1, (n, 1))
s = x - 0.2 - 0.006 * t
env, osc = np.exp(-1000 * s**2), np.sin(400 * s)
u = env * osc
du_ds = env * (400 * np.cos(400 * s) - 2000 * s * osc)
tn, xn = (t - 50) / 50, (x - 0.5) / 0.5                        # d/dtn = 50 d/dt, d/dxn = 0.5 d/dx
inp = np.hstack([tn, xn]).astype(np.float32)
tgt = np.hstack([u, du_ds * (-0.006) * 50, du_ds * 0.5]).astype(np.float32)   # [u, du/dtn, du/dxn]

model_ori = nif.NIFMultiScale(cfg_shape_net, cfg_parameter_net, "float32")
model = nif.JacobianLayer(model_ori.build(), y_index=[0], x_index=[0, 1]).as_model()   # output [u, du/dt, du/dx]
model.compile(nif.Adam(1e-3), loss=nif.SobolevMSE(coef_grad=1e-3, value_cols=[0], grad_cols=[2]))
ds = nif.Dataset.from_tensor_slices((inp, tgt)).shuffle(n).batch(65536)
hist = model.fit(ds, epochs=int(os.environ.get("NEPOCH", "20")), verbose=1)

y, J, H = nif.HessianLayer(model_ori.build(), y_index=[0], x_index=[0, 1])(inp[:1000])
print("u", tuple(y.shape), "Jacobian", tuple(J.shape), "Hessian", tuple(H.shape))   # (1000,1) (1000,1,2) (1000,1,2,2)
print("u_xx sample", H[:3, 0, 1, 1].cpu().numpy())
