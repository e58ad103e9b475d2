"""tutorial/1_simple_1d_wave.ipynb of the reference, on nif_b200 (needs a B200; `python examples/tutorial1_simple_1d_wave.py`).

The only edits against the notebook: `import nif_b200 as nif`, the optimiser / dataset / callback classes come from the
same module instead of tf.keras / tf.data, and checkpoints are `.npz` files."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nif_b200 as nif  # noqa: E402
from nif_b200.demo import TravelingWave  # noqa: E402

cfg_shape_net = {"connectivity": "full", "input_dim": 1, "output_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
cfg_parameter_net = {"input_dim": 1, "latent_dim": 1, "units": 30, "nlayers": 2, "activation": "swish"}
nepoch = int(os.environ.get("NEPOCH", "200"))
batch_size = 512

tw = TravelingWave()
train_data = tw.data.astype(np.float32)                       # (2000, 3): t, x, u (standard-normalised)
train_inp, train_tgt = train_data[:, :2], train_data[:, 2:3]
train_dataset = nif.Dataset.from_tensor_slices((train_inp, train_tgt)).shuffle(2000).batch(batch_size).prefetch(1)

model_ori = nif.NIF(cfg_shape_net, cfg_parameter_net, "float32")
model_opt = model_ori.build()                                 # was: model_ori.build() under MirroredStrategy().scope()
model_opt.compile(nif.Adam(1e-3), loss="mse")


def scheduler(epoch, lr):
    return 1e-3 if epoch < 1000 else 5e-4


history = model_opt.fit(train_dataset, epochs=nepoch, verbose=0, callbacks=[nif.LearningRateScheduler(scheduler)])
print("final loss", history.history["loss"][-1])
model_opt.save_weights("./saved_weights/ckpt-{}/ckpt".format(nepoch))

u_pred = model_opt.predict(train_inp)                        # (2000, 1)
print("relative L2 error", float(np.linalg.norm(u_pred - train_tgt) / np.linalg.norm(train_tgt)))

# the extraction models share the trained variables (tutorial cell 22 onwards)
latent = model_ori.model_p_to_lr().predict(train_inp[:10, :1])                       # t -> latent code
w = model_ori.model_lr_to_w().predict(latent)                                         # latent -> ShapeNet weights
u_again = model_ori.model_x_to_u_given_w().predict([train_inp[:10, 1:2], w])          # (x, weights) -> u
print("extraction path consistent:", bool(np.allclose(u_again, u_pred[:10], atol=1e-5)))
