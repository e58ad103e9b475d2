"""Keras-shaped training / inference runtime around the fused kernels.

The reference hands its functional graph to Keras (`Model.compile / fit / predict / save_weights`,
tutorial/2_multi_scale_NIF.ipynb:643-656).  Keras itself is third-party there (tensorflow==2.11.1);
the semantics kept here are the ones SURVEY A.6 lists: 'mse' = mean over outputs then (weighted)
mean over the global batch, Adam with epsilon outside the bias correction, LearningRateScheduler
called at the start of every epoch, short last batch kept, predict() batched.
"""
from __future__ import annotations

import math
import os
import time
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from ._lib import NifError
from . import ops


# --------------------------------------------------------------------------------------------------
# optimiser / callbacks / dataset
# --------------------------------------------------------------------------------------------------
class Adam:
    """tf.keras.optimizers.Adam(learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7)."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate = float(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.iterations = 0
        self._m = None
        self._v = None

    @property
    def lr(self):
        return self.learning_rate

    @lr.setter
    def lr(self, v):
        self.learning_rate = float(v)

    def _ensure(self, theta):
        if self._m is None or self._m.shape != theta.shape or self._m.device != theta.device:
            self._m = torch.zeros_like(theta)
            self._v = torch.zeros_like(theta)

    def apply(self, theta: torch.Tensor, grad: torch.Tensor, l1=0.0, l2=0.0, g_scale=1.0):
        self._ensure(theta)
        self.iterations += 1
        ops.adam_step(theta, grad, self._m, self._v, self.learning_rate, self.iterations, self.beta_1, self.beta_2,
                      self.epsilon, l1, l2, g_scale)

    def apply_multimem(self, theta: torch.Tensor, symm: dict, rank: int, world: int, l1=0.0, l2=0.0, g_scale=1.0):
        """The data-parallel update fused with its collective over NVSwitch multicast memory (nif_adam_step_multimem);
        the caller brackets it with cross-rank barriers."""
        import ctypes as C
        from . import _lib
        self._ensure(theta)
        self.iterations += 1
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().nif_adam_step_multimem(theta.numel(), rank, world, C.c_void_p(symm["p_mc"]), C.c_void_p(symm["g_mc"]),
                                                     C.c_void_p(theta.data_ptr()), C.c_void_p(self._m.data_ptr()),
                                                     C.c_void_p(self._v.data_ptr()), self.learning_rate, None, self.beta_1,
                                                     self.beta_2, self.epsilon, self.iterations, float(l1), float(l2),
                                                     float(g_scale), st), "nif_adam_step_multimem")

    # graph-replayed steps: the update is recorded once with its step size in device memory (record_apply, inside the
    # capture); before every replay the host forms this step's alpha exactly like nif_adam_step does and stores it.
    def record_apply(self, theta: torch.Tensor, grad: torch.Tensor, l1=0.0, l2=0.0, g_scale=1.0):
        ops.adam_step_dev(theta, grad, self._m, self._v, self._alpha_dev, self.beta_1, self.beta_2, self.epsilon, l1, l2,
                          g_scale)

    def prepare_replay(self, theta: torch.Tensor):
        self._ensure(theta)
        if getattr(self, "_alpha_dev", None) is None or self._alpha_dev.device != theta.device:
            self._alpha_dev = torch.zeros(4, dtype=torch.float32, device=theta.device)  # 16 bytes: aligned like the rest

    def advance_replay(self):
        self.iterations += 1
        t = float(self.iterations)
        alpha = self.learning_rate * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)
        self._alpha_dev.fill_(alpha)  # by-value launch argument of a fill kernel: stream-ordered, no host buffer to race on


class _nvtx:
    """NVTX range around a phase of the step when NIF_B200_NVTX=1 (Nsight Systems / `ncu --nvtx` filters); a no-op
    otherwise.  The kernels themselves are named; the ranges say which phase of which step launched them."""
    ON = os.environ.get("NIF_B200_NVTX", "0") == "1"

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if self.ON:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if self.ON:
            torch.cuda.nvtx.range_pop()
        return False


class SobolevMSE:
    """The `Sobolov_MSE` loss of tutorial 8 (tutorial/8_NIF_with_Sobolov_training.ipynb:815-820) for a model built
    with JacobianLayer.as_model(): per row  sum_{c in value_cols} (t_c - p_c)^2 + coef_grad * sum_{c in grad_cols}
    (t_c - p_c)^2, averaged over the batch.  Columns index the concatenated output [y | dy/dx flattened]; the
    tutorial uses value_cols=[0], grad_cols=[2] (u and du/dx; du/dt is only monitored).  Grad columns may differentiate
    with respect to any inputs, ShapeNet coordinates and ParameterNet inputs alike (up to four distinct ones).
    Like the tutorial's `Sobolov_MSE`, each group is a MEAN over its columns (tf.reduce_mean(..., axis=-1))."""

    def __init__(self, coef_grad=1e-3, value_cols=(0,), grad_cols=(2,)):
        self.coef_grad = float(coef_grad)
        self.value_cols = [int(c) for c in value_cols]
        self.grad_cols = [int(c) for c in grad_cols]

    def __call__(self, y_true, y_pred):
        sd = ((y_true[:, self.value_cols] - y_pred[:, self.value_cols]) ** 2).mean(-1)
        sg = ((y_true[:, self.grad_cols] - y_pred[:, self.grad_cols]) ** 2).mean(-1)
        return (sd + self.coef_grad * sg).mean()


class Callback:
    model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None): ...
    def on_train_end(self, logs=None): ...
    def on_epoch_begin(self, epoch, logs=None): ...
    def on_epoch_end(self, epoch, logs=None): ...
    def on_train_batch_end(self, batch, logs=None): ...


class LearningRateScheduler(Callback):
    """tf.keras.callbacks.LearningRateScheduler(schedule): lr = schedule(epoch, lr) at epoch begin
    (tutorial/2_multi_scale_NIF.ipynb:631-641)."""

    def __init__(self, schedule: Callable[[int, float], float], verbose=0):
        self.schedule = schedule
        self.verbose = verbose

    def on_epoch_begin(self, epoch, logs=None):
        opt = self.model.optimizer
        opt.learning_rate = float(self.schedule(epoch, opt.learning_rate))


class History(Callback):
    def on_train_begin(self, logs=None):
        self.history: Dict[str, List[float]] = {}
        self.epoch: List[int] = []

    def on_epoch_end(self, epoch, logs=None):
        self.epoch.append(epoch)
        for k, v in (logs or {}).items():
            self.history.setdefault(k, []).append(v)


class Dataset:
    """The slice of tf.data the tutorials use: from_tensor_slices(...).shuffle(N).batch(B).prefetch(...)
    (tutorial/2_multi_scale_NIF.ipynb:222-224).  Arrays are staged once in pinned host memory;
    every epoch draws a fresh permutation when shuffle() was requested; the short last batch is kept."""

    def __init__(self, arrays: Sequence[np.ndarray]):
        self.arrays = [torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32) for a in arrays]
        n = {a.shape[0] for a in self.arrays}
        if len(n) != 1:
            raise ValueError("all arrays must have the same first dimension")
        self.n = n.pop()
        self._shuffle = False
        self._batch = None
        self._seed = 0
        self._rank, self._world = 0, 1

    @staticmethod
    def from_tensor_slices(arrays) -> "Dataset":
        if isinstance(arrays, (tuple, list)):
            return Dataset(list(arrays))
        return Dataset([arrays])

    def shuffle(self, buffer_size=None, seed=None) -> "Dataset":
        self._shuffle = True
        if seed is not None:
            self._seed = int(seed)
        return self

    def batch(self, batch_size) -> "Dataset":
        self._batch = int(batch_size)
        return self

    def prefetch(self, *_a) -> "Dataset":
        return self

    def shard(self, world: int, rank: int) -> "Dataset":
        """Data parallel: every rank sees rows rank::world of every global batch."""
        self._world, self._rank = int(world), int(rank)
        return self

    def pin(self):
        if torch.cuda.is_available():
            self.arrays = [a if (a.is_cuda or a.is_pinned()) else a.pin_memory() for a in self.arrays]
        return self

    def nbytes(self) -> int:
        return sum(a.numel() * a.element_size() for a in self.arrays)

    def cache_on(self, device) -> "Dataset":
        """Keep the arrays in device memory: the per-epoch permutation and the batch gathers then run on the device and
        no batch crosses PCIe (a tutorial-sized data set is a few MB; 180 GB of HBM hold every data set of the
        reference).  Same batches, in the same order, as the host path."""
        self.arrays = [a.to(device) for a in self.arrays]
        return self

    def batches(self, epoch: int):
        bs = self._batch or self.n
        if self._shuffle:
            g = torch.Generator().manual_seed(self._seed * 1000003 + epoch)
            perm = torch.randperm(self.n, generator=g)
            if self.arrays[0].device != perm.device:
                perm = perm.to(self.arrays[0].device)  # device-resident arrays: one transfer per epoch, gathers on the device
        else:
            perm = None
        for s in range(0, self.n, bs):
            e = min(self.n, s + bs)
            idx = slice(s, e) if perm is None else perm[s:e]
            rows = [a[idx] for a in self.arrays]
            gb = e - s
            if self._world > 1:
                rows = [r[self._rank::self._world] for r in rows]
            yield gb, rows


# --------------------------------------------------------------------------------------------------
# the model object returned by NIF.build() / .model() / .model_*()
# --------------------------------------------------------------------------------------------------
class Model:
    """What `nif.NIF(...).build()` returns in the reference is a tf.keras.Model; this is the same
    surface: __call__/predict, compile, fit, train_on_batch, save_weights/load_weights, summary,
    trainable_variables.  `kind` selects the (sub-)graph (nif/model.py:379-464, 956-986)."""

    def __init__(self, net, kind: str):
        self.net = net
        self.kind = kind
        self.optimizer: Optional[Adam] = None
        self.loss = None
        self.metrics_fns: list = []
        self.stop_training = False
        self.history = None
        self._loss_buf = None
        self._packed = None
        self.dist = None  # set by nif_b200.distributed.DataParallel
        self._symm = None  # symmetric-memory handles of the flat buffers (DataParallel.attach), or None
        self.use_graph: Optional[bool] = None  # None: NIF_B200_GRAPH env (default on); see _train_step_graph
        self._graphs: Dict[tuple, dict] = {}

    # ---- plumbing -----------------------------------------------------------------------------------
    def _dev(self, a, pinned_ok=True) -> torch.Tensor:
        if isinstance(a, torch.Tensor):
            t = a
        else:
            t = torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32)
        if t.dtype != torch.float32:
            t = t.float()
        if self.net.device.type != "cuda":
            raise NifError("nif_b200 needs a CUDA device: there is no CPU fallback for the hot path")
        return t.to(self.net.device, non_blocking=True)

    @property
    def trainable_variables(self):
        return self.net.trainable_variables

    @property
    def inputs(self):
        n = self.net
        return {"full": [("input_tot", n.pi_dim + n.si_dim)], "jacobian": [("input_tot", n.pi_dim + n.si_dim)],
                "p_to_w": [("input_p_to_w", n.pi_dim)],
                "p_to_lr": [("input_p_to_lr", n.pi_dim)], "lr_to_w": [("input_lr_to_w", n.pi_hidden)],
                "x_to_phi": [("input_x_to_phi", n.si_dim)],
                "x_to_u_given_w": [("input_x_to_u_given_w", n.si_dim), ("input_w_and_b_from_pnet", n.po_dim)]}[self.kind]

    def count_params(self) -> int:
        n = self.net
        last = n.pi_hidden * n.po_dim + n.po_dim
        if getattr(n, "_last_layer_only", False):
            return n._count_params_kind(self.kind)
        return {"full": n.count_params(), "jacobian": n.count_params(), "p_to_w": n.count_params(),
                "p_to_lr": n.count_params() - last,
                "lr_to_w": last, "x_to_u_given_w": 0}[self.kind]

    def summary(self, print_fn=print):
        n = self.net
        print_fn(f'Model: "{type(n).__name__}.{self.kind}"')
        print_fn("_" * 65)
        names = list(n.variables)
        if self.kind == "p_to_lr":
            names = names[:-2]
        elif self.kind == "lr_to_w":
            names = names[-2:]
        elif self.kind == "x_to_u_given_w":
            names = []
        for k in names:
            v = n.variables[k]
            print_fn(f"{k:45s} {str(tuple(v.shape)):>12s} {v.numel():>7d}")
        print_fn("=" * 65)
        print_fn(f"Total params: {self.count_params():,}")
        print_fn(f"Trainable params: {self.count_params():,}")

    def _packed_weights(self) -> torch.Tensor:
        n = self.net
        self._packed = n.engine.pack(n.w_h.detach(), n.b_h.detach(), out=self._packed)
        return self._packed

    # ---- forward ------------------------------------------------------------------------------------
    def _latent_nograd(self, p_in: torch.Tensor) -> torch.Tensor:
        n = self.net
        if n._trunk is not None:
            return n._trunk.forward(p_in.contiguous(), n.theta_trunk)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(n.compute_Dtype == "bfloat16")):
            z = n._latent(p_in)
        return z.float()

    @torch.no_grad()
    def _forward_batch(self, x):
        n = self.net
        if getattr(n, "_last_layer_only", False):
            if self.kind == "x_to_u_given_w":
                xs, w = x
                return n._forward_kind(self.kind, self._dev(xs), self._dev(w))
            return n._forward_kind(self.kind, self._dev(x))
        if self.kind == "full":
            inp = self._dev(x)
            z = self._latent_nograd(inp[:, : n.pi_dim])
            xs = inp[:, n.pi_dim: n.pi_dim + n.si_dim].contiguous()
            return n.engine.forward(z.contiguous(), xs, self._packed_weights())
        if self.kind == "jacobian":
            u, J = self._jacobian_forward(self._dev(x))
            return torch.cat([u, J.reshape(u.shape[0], -1)], -1)
        if self.kind == "p_to_lr":
            return self._latent_nograd(self._dev(x))
        if self.kind == "p_to_w":
            z = self._latent_nograd(self._dev(x))
            return torch.addmm(n.b_h.detach(), z, n.w_h.detach())
        if self.kind == "lr_to_w":
            return torch.addmm(n.b_h.detach(), self._dev(x), n.w_h.detach())
        if self.kind == "x_to_u_given_w":
            xs, w = x
            return n.engine.given_w(self._dev(xs), self._dev(w))
        raise ValueError(self.kind)

    def __call__(self, x, training=False):
        return self._forward_batch(x)

    @torch.no_grad()
    def _jacobian_forward(self, inp: torch.Tensor, y_index=None, x_index=None):
        """(y, J[b,a,c] = dy[b, y_index[a]] / d input[b, x_index[c]]) with forward-mode tangents: one direction per
        requested input column; directions on ParameterNet inputs carry the trunk tangent of the latent code
        (JacobianLayer, nif/layers/gradient.py:36-49, 207-231)."""
        n = self.net
        y_index = self.jac_y if y_index is None else y_index
        x_index = self.jac_x if x_index is None else x_index
        B = inp.shape[0]
        p_in = inp[:, : n.pi_dim].contiguous()
        xs = inp[:, n.pi_dim: n.pi_dim + n.si_dim].contiguous()
        nd = len(x_index)
        zdot = torch.zeros(nd, B, n.pi_hidden, device=inp.device)
        xdot = torch.zeros(nd, B, n.si_dim, device=inp.device)
        z = None
        for d, c in enumerate(x_index):
            if c < n.pi_dim:
                e = torch.zeros_like(p_in)
                e[:, c] = 1.0
                z, zd = torch.func.jvp(n._latent, (p_in,), (e,))
                zdot[d] = zd
            elif c < n.pi_dim + n.si_dim:
                xdot[d, :, c - n.pi_dim] = 1.0
            else:
                raise IndexError(f"x_index {c} outside the {n.pi_dim + n.si_dim} model inputs")
        if z is None:
            z = self._latent_nograd(p_in)
        if all(c >= n.pi_dim for c in x_index):
            zdot = None  # the latent code does not move
        u, udot = n.engine.forward_tangent(z.contiguous(), xs, self._packed_weights(), zdot, xdot)  # udot [nd, B, so]
        J = udot.permute(1, 2, 0)[:, y_index, :]  # [B, |y|, |x|]
        return u, J.contiguous()

    @torch.no_grad()
    def _hessian_forward(self, inp: torch.Tensor, y_index, x_index):
        """(y, J[b,a,c], H[b,a,c,e] = d2 y[b, y_index[a]] / d input[b, x_index[c]] d input[b, x_index[e]]) with
        second-order forward mode: one launch per unordered pair of requested input columns carries h, the two first
        tangents and the mixed second tangent through the fused kernel (HessianLayer, nif/layers/gradient.py:130-180,
        234-261).  Directions on ParameterNet inputs take the first and second directional derivatives of the latent
        code from the trunk (torch forward-mode AD)."""
        n = self.net
        B = inp.shape[0]
        dev = inp.device
        p_in = inp[:, : n.pi_dim].contiguous()
        xs = inp[:, n.pi_dim: n.pi_dim + n.si_dim].contiguous()
        nx, ny = len(x_index), len(y_index)
        for c in x_index:
            if not 0 <= c < n.pi_dim + n.si_dim:
                raise IndexError(f"x_index {c} outside the {n.pi_dim + n.si_dim} model inputs")
        jvp = torch.func.jvp

        def unit(c):
            e = torch.zeros_like(p_in)
            e[:, c] = 1.0
            return e

        z = self._latent_nograd(p_in).contiguous()
        zd = {c: jvp(n._latent, (p_in,), (unit(c),))[1] for c in set(x_index) if c < n.pi_dim}
        packed = self._packed_weights()
        u = None
        J = torch.zeros(B, ny, nx, device=dev)
        Hs = torch.zeros(B, ny, nx, nx, device=dev)
        for ia in range(nx):
            for ib in range(ia, nx):
                ca, cb = x_index[ia], x_index[ib]
                zdot = torch.zeros(2, B, n.pi_hidden, device=dev)
                xdot = torch.zeros(2, B, n.si_dim, device=dev)
                for k, c in enumerate((ca, cb)):
                    if c < n.pi_dim:
                        zdot[k] = zd[c]
                    else:
                        xdot[k, :, c - n.pi_dim] = 1.0
                zddot = None
                if ca < n.pi_dim and cb < n.pi_dim:  # both directions act on the trunk: its second directional derivative
                    ea, eb = unit(ca), unit(cb)
                    zddot = jvp(lambda q: jvp(n._latent, (q,), (ea,))[1], (p_in,), (eb,))[1].contiguous()
                u, ud, udd = n.engine.forward_tangent2(z, xs, packed, zdot, xdot, zddot)
                J[:, :, ia] = ud[0][:, y_index]
                J[:, :, ib] = ud[1][:, y_index]
                Hs[:, :, ia, ib] = udd[:, y_index]
                Hs[:, :, ib, ia] = udd[:, y_index]
        if u is None:
            u = n.engine.forward(z, xs, packed)
        return u, J, Hs

    def predict(self, x, batch_size=None, verbose=0, **_kw) -> np.ndarray:
        """Keras predict: batched forward, numpy out.  (Keras' default batch of 32 only affects speed;
        rows are independent, so a larger internal batch returns the same values.)"""
        bs = int(batch_size) if batch_size else 65536
        first = x[0] if isinstance(x, (tuple, list)) else x
        N = first.shape[0]
        outs = []
        for s in range(0, N, bs):
            xb = [a[s:s + bs] for a in x] if isinstance(x, (tuple, list)) else x[s:s + bs]
            outs.append(self._forward_batch(xb).cpu())
        if not outs:
            return np.zeros((0, self.net.so_dim), np.float32)
        return torch.cat(outs, 0).numpy()

    def predict_latent_grid(self, latents, coords, group_chunk: Optional[int] = None, shard: bool = False) -> torch.Tensor:
        """Factored form of model_x_to_u_given_w for sweeps:
        latents [G,K] x coords [N,si] -> u [G,N,so]; per-latent weights are generated once
        (G x P, a plain GEMM) and the G ShapeNets run as grouped launches over the shared grid.

        group_chunk: latents per launch (bounds the packed weight images and the output held at once; default: all).
        shard=True (data parallel, model attached to a DataParallel): this rank evaluates latents[rank::world] only --
        the sweep partitions over the latent axis with no collective -- and returns its [ceil(G/world), N, so] block."""
        n = self.net
        with torch.no_grad():
            zg = self._dev(latents)
            if shard and self.dist is not None and self.dist.world > 1:
                zg = zg[self.dist.rank::self.dist.world].contiguous()
            eng0 = getattr(self, "_eng0", None)
            if eng0 is None:
                eng0 = self._eng0 = n.engine.with_latent(0)
            xs = self._dev(coords)
            G = zg.shape[0]
            step = int(group_chunk) if group_chunk else max(G, 1)
            outs = []
            for s0 in range(0, G, step):
                w = torch.addmm(n.b_h.detach(), zg[s0:s0 + step], n.w_h.detach())  # [g,P]
                packed = eng0.pack(None, w)
                u = eng0.forward(None, xs, packed, groups=w.shape[0], x_shared=True)
                outs.append(u.view(w.shape[0], xs.shape[0], n.so_dim))
            if not outs:
                return torch.empty(0, xs.shape[0], n.so_dim, device=xs.device)
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    # ---- training -----------------------------------------------------------------------------------
    def compile(self, optimizer=None, loss="mse", metrics=None, graph: Optional[bool] = None, **_kw):
        self.use_graph = graph
        self._graphs = {}
        if self.kind == "jacobian":
            # SobolevMSE names its columns; 'mse' or any callable(y_true, y_pred) sees the whole concatenated output
            # [y | dy/dx], so every requested Jacobian entry becomes a direction / seed of the reverse-over-forward pass
            self._sobolev_plan = self._plan_sobolev(loss)
        elif self.kind != "full":
            raise NifError("only the full model is trainable")
        if optimizer is None:
            optimizer = Adam()
        if not callable(getattr(optimizer, "apply", None)) or not hasattr(optimizer, "learning_rate"):
            raise NifError("optimizer must be nif_b200.Adam or one of nif_b200.optimizers (AdaBeliefOptimizer, Lion)")
        if not (isinstance(loss, str) and loss in ("mse", "mean_squared_error") or callable(loss)):
            raise NifError("loss must be 'mse' or a callable(y_true, y_pred) -> scalar tensor")
        self.optimizer, self.loss = optimizer, loss
        # callable metrics(y_true, y_pred) -> scalar are evaluated on every batch's predictions and averaged per epoch
        self.metrics_fns = list(metrics or [])
        for mfn in self.metrics_fns:
            if not callable(mfn):
                raise NifError("metrics must be callables(y_true, y_pred) -> scalar tensor")

    def train_on_batch(self, x, y, sample_weight=None, global_batch: Optional[int] = None) -> float:
        """One optimisation step (what Keras' train_step + apply_gradients do).  Returns the loss of this
        process's rows already divided by the global batch (sum over ranks = global loss)."""
        if (self.kind == "full" and self.optimizer is not None and self._fusable() and self._graph_enabled()
                and all(isinstance(t, torch.Tensor) and t.dtype == torch.float32 and t.device.type == "cpu" and t.is_contiguous()
                        for t in (x, y) + (() if sample_weight is None else (sample_weight,)))):
            # host batches of a graph-replayed step go straight into the graph's input buffers (one host-to-device copy
            # each, no device temporary and device-to-device copy in between)
            loss = self._train_step(x, y, sample_weight, global_batch)
        else:
            loss = self._train_step(self._dev(x), self._dev(y), None if sample_weight is None else self._dev(sample_weight),
                                    global_batch)
        return float(loss)

    def _train_step(self, inp: torch.Tensor, tgt: torch.Tensor, sw: Optional[torch.Tensor],
                    global_batch: Optional[int]) -> torch.Tensor:
        if self.kind == "jacobian":
            return self._train_step_sobolev(inp, tgt, global_batch)
        n = self.net
        B = inp.shape[0]
        gb = int(global_batch) if global_batch else B
        if self._loss_buf is None:
            self._loss_buf = torch.zeros(1, dtype=torch.float32, device=n.device)
        if inp.device.type != "cuda" and not (self._fusable() and B > 0 and self._graph_enabled()):
            inp, tgt, sw = self._dev(inp), self._dev(tgt), (None if sw is None else self._dev(sw))
        if self._fusable():
            if B > 0 and self._graph_enabled():
                return self._train_step_graph(inp, tgt, sw, gb)
            self._loss_buf.zero_()
            self._fused_step(inp, tgt, sw, gb, self.optimizer.apply)
            return self._loss_buf
        if (B > 0 and not callable(self.loss) and self._graph_enabled() and not isinstance(n.p_jac_reg, (float, int))
                and os.environ.get("NIF_B200_GRAPH_GENERAL", "1") != "0"):
            return self._general_step_graph(inp, tgt, sw, gb)
        loss = self._loss_and_grad(inp, tgt, self.loss, sw, gb, with_regularisers=True)
        if self._symm_ready():
            self._symm_update()
            return loss
        if self.dist is not None:
            self.dist.allreduce_(n.grad)
        l1, l2 = n._kernel_regulariser()
        self._apply_update(self.optimizer.apply, l1, l2)
        return loss

    def _symm_ready(self) -> bool:
        return (self.dist is not None and getattr(self, "_symm", None) is not None and isinstance(self.optimizer, Adam)
                and not getattr(self.optimizer, "_centralize_in_fit", False))

    def _symm_update(self):
        """barrier (every replica's gradient is complete) -> one kernel: reduce-scatter + Adam + all-gather over NVSwitch
        multicast memory -> barrier (every parameter slice has landed, every gradient has been read)."""
        n, s = self.net, self._symm
        l1, l2 = n._kernel_regulariser()
        s["hg"].barrier(channel=0)
        self.optimizer.apply_multimem(n.theta, s, self.dist.rank, self.dist.world, l1, l2)
        s["hg"].barrier(channel=1)

    def _fusable(self) -> bool:
        """The fully fused step: 'mse' loss, fused trunk kernels, a hyper-network head (not the last-layer-parameterised
        class, whose ShapeNet is a shared-weight MLP)."""
        n = self.net
        return not callable(self.loss) and n._trunk is not None and not getattr(n, "_last_layer_only", False)

    def _loss_and_grad(self, inp: torch.Tensor, tgt: torch.Tensor, loss, sw: Optional[torch.Tensor] = None,
                       gb: Optional[int] = None, with_regularisers: bool = False) -> torch.Tensor:
        """Loss of this process's rows (already divided by the global batch) and its gradient with respect to every
        variable, left in the flat gradient buffer; no update.  `loss` is 'mse' or a callable(y_true, y_pred) (Keras
        order).  The general path of _train_step and the closure of the L-BFGS fine-tuner."""
        n = self.net
        B = inp.shape[0]
        gb = int(gb) if gb else B
        if self._loss_buf is None:
            self._loss_buf = torch.zeros(1, dtype=torch.float32, device=n.device)
        self._loss_buf.zero_()
        n.grad.zero_()
        p_in = inp[:, : n.pi_dim]
        xs = inp[:, n.pi_dim: n.pi_dim + n.si_dim].contiguous()
        extra = []  # differentiable torch terms added to the loss
        if with_regularisers and isinstance(n.p_jac_reg, (float, int)):
            extra.append(n._jac_reg_loss(p_in) * (B / gb))  # JacRegLatentLayer's add_loss term, this process's share
        if getattr(n, "_last_layer_only", False):
            u = n._forward_train(inp)  # shared-weight ShapeNet: torch autograd over library GEMMs
            lv = (self._mse_torch(tgt, u, sw) if not callable(loss) else loss(tgt, u)) * (B / gb)
            for e in extra:
                lv = lv + e
            lv.backward()
            out = lv.detach().reshape(1)
        else:
            eng = n.engine
            # mixed_bfloat16: the ParameterNet's Dense / SIREN layers compute in bfloat16 too (nif/model.py:101-105)
            n._first_order_only = not extra  # no jac_reg term: the trunk is differentiated once
            try:
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(n.compute_Dtype == "bfloat16")):
                    z = n._latent(p_in)
            finally:
                n._first_order_only = False
            z = z.float()
            act = self._activity_terms(z.detach(), B, gb) if with_regularisers else None
            if not callable(loss):
                packed = self._packed_weights()
                zc = z.detach().contiguous()
                u, stash = eng.forward(zc, xs, packed, save=True)
                dz = eng.mse_backward(zc, xs, packed, u, stash, tgt, sw, 1.0 / gb, self._loss_buf,
                                      n._gviews[n._last_names[0]], n._gviews[n._last_names[1]], 0.0)
                if act is not None:
                    dz = dz + act[1]
                    n._gviews[n._last_names[0]].add_(act[2])
                    n._gviews[n._last_names[1]].add_(act[3])
                    self._loss_buf += act[0]
                torch.autograd.backward([z] + extra, [dz] + [torch.ones_like(e) for e in extra])
                for e in extra:
                    self._loss_buf += e.detach()
                out = self._loss_buf
            else:
                u = ops.fused_shapenet(z, xs, n.w_h, n.b_h, eng)
                lv = loss(tgt, u) * (B / gb)
                for e in extra:
                    lv = lv + e
                if act is not None:
                    torch.autograd.backward([lv, z], [torch.ones_like(lv), act[1]])
                else:
                    lv.backward()
                out = lv.detach().reshape(1)
                if act is not None:
                    n._gviews[n._last_names[0]].add_(act[2])
                    n._gviews[n._last_names[1]].add_(act[3])
                    out = out + act[0]
        if with_regularisers:
            reg = self._reg_loss()
            if reg is not None:
                out = out + reg.detach() * (B / gb)
        return out

    @staticmethod
    def _mse_torch(y_true, y_pred, sw):
        per_row = ((y_true - y_pred) ** 2).mean(-1)
        if sw is not None:
            per_row = per_row * sw.reshape(-1)
        return per_row.mean()

    def _fused_part1(self, inp, tgt, sw, gb):
        """Trunk forward, hyper-network head + ShapeNet, loss, reverse pass of the head: every gradient of the last
        linear layer is written (beta = 0) straight into the flat gradient buffer.  Returns what the trunk's reverse
        pass needs.  Only stream-ordered work (no host read, no allocation outside torch's caching allocator), so the
        same body is what a CUDA graph records."""
        n = self.net
        eng = n.engine
        B = inp.shape[0]
        xs = inp[:, n.pi_dim: n.pi_dim + n.si_dim].contiguous()
        p_in = inp[:, : n.pi_dim].contiguous()
        # the re-laid weight image depends on the parameters only: it is built on a second stream while the trunk runs
        # (a fork / join that a CUDA graph records as two parallel branches)
        cur = torch.cuda.current_stream()
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=n.device)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            packed = self._packed_weights()
        z, tstash = n._trunk.forward(p_in, n.theta_trunk, save=True)
        cur.wait_stream(self._side)
        u, stash = eng.forward(z, xs, packed, save=True)
        act = self._activity_terms(z, B, gb)
        # The trunk's reverse pass needs dz only, which the first kernel of the head's reverse pass produces: it starts on the
        # second stream as soon as dz is final and runs next to the weight-gradient kernels (they leave SMs idle: 135 CTAs
        # on 148 SMs at C2).  Not when dz is modified afterwards (activity regularisers) or when the two kernel sequences
        # are recorded as separate graphs around an NCCL all-reduce.
        overlap = (act is None and (self.dist is None or self._symm_ready())
                   and os.environ.get("NIF_B200_OVERLAP_TRUNK", "1") != "0")
        ev = None
        if overlap:
            ev = getattr(self, "_dz_event", None)
            if ev is None:
                with torch.cuda.stream(self._side):  # (created by its first record, outside any capture of the main stream)
                    ev = self._dz_event = torch.cuda.Event()
                    ev.record()
        dz = eng.mse_backward(z, xs, packed, u, stash, tgt, sw, 1.0 / gb, self._loss_buf,
                              n._gviews[n._last_names[0]], n._gviews[n._last_names[1]], 0.0, dz_event=ev)
        if overlap:
            self._side.wait_event(ev)
            with torch.cuda.stream(self._side):
                n._trunk.backward(p_in, n.theta_trunk, tstash, dz, n.grad_trunk, 0.0)
            reg = self._reg_loss()
            if reg is not None:
                self._loss_buf += reg * (B / gb)
            return p_in, tstash, dz, True
        if act is not None:
            dz = dz + act[1]
            n._gviews[n._last_names[0]].add_(act[2])
            n._gviews[n._last_names[1]].add_(act[3])
            self._loss_buf += act[0]
        reg = self._reg_loss()
        if reg is not None:
            self._loss_buf += reg * (B / gb)
        return p_in, tstash, dz, False

    def _fused_part2(self, ctx):
        n = self.net
        p_in, tstash, dz, launched = ctx
        if launched:  # the trunk's reverse pass is already running on the second stream: join
            torch.cuda.current_stream().wait_stream(self._side)
            return
        n._trunk.backward(p_in, n.theta_trunk, tstash, dz, n.grad_trunk, 0.0)

    def _fused_step(self, inp, tgt, sw, gb, apply_update, part1=None, part2=None):
        """One fully fused optimisation step; `part1` / `part2` replace the two kernel sequences by graph replays.
        Data parallel: the last linear layer's gradient (almost all of the buffer) is summed across ranks while the
        trunk's reverse pass runs; the small trunk gradient follows."""
        n = self.net
        if self._symm_ready() and apply_update is not None:
            with _nvtx("nif.step.forward+head_reverse"):
                ctx = part1() if part1 is not None else self._fused_part1(inp, tgt, sw, gb)
            with _nvtx("nif.step.trunk_reverse"):
                if part2 is not None:
                    part2()
                else:
                    self._fused_part2(ctx)
            with _nvtx("nif.step.multimem_update"):
                self._symm_update()
            return
        with _nvtx("nif.step.forward+head_reverse"):
            ctx = part1() if part1 is not None else self._fused_part1(inp, tgt, sw, gb)
        h_head = self.dist.allreduce_start(n.grad[n._n_trunk:]) if self.dist is not None else None
        with _nvtx("nif.step.trunk_reverse"):
            if part2 is not None:
                part2()
            else:
                self._fused_part2(ctx)
        if self.dist is not None:
            with _nvtx("nif.step.allreduce"):
                h_trunk = self.dist.allreduce_start(n.grad_trunk)
                self.dist.allreduce_finish(h_head)
                self.dist.allreduce_finish(h_trunk)
        l1, l2 = n._kernel_regulariser()
        with _nvtx("nif.step.update"):
            self._apply_update(apply_update, l1, l2)

    # ---- CUDA-graph replay of the fused step ------------------------------------------------------------------
    # One optimisation step is 17 kernel launches from Python; at tutorial-1 sizes (512 rows) the launches, not the
    # kernels, set the step time.  The step body is stream-ordered and allocation-free apart from torch's caching
    # allocator, so it is recorded once per (rows, global batch, sample-weighted?) and replayed; per step the host
    # only copies the batch into the graph's input buffers and stores Adam's bias-corrected step size.
    # Data parallel: the two kernel sequences either side of the first all-reduce are recorded as two graphs sharing
    # one memory pool; the NCCL calls and the Adam launch between / after them stay eager.
    GRAPH_CACHE = 4
    DEVICE_CACHE_BYTES = 8 << 30  # fit() keeps data sets up to this size resident in HBM

    def _graph_enabled(self) -> bool:
        g = self.use_graph
        if g is None:
            g = os.environ.get("NIF_B200_GRAPH", "1") != "0"
        if not hasattr(self.optimizer, "record_apply") or getattr(self.optimizer, "_centralize_in_fit", False):
            return False  # only Adam records its update into a graph; centralisation adds per-variable launches
        return bool(g)

    # ---- regularisers of the ParameterNet (nif/model.py:107-125) ---------------------------------------------------
    def _reg_loss(self) -> Optional[torch.Tensor]:
        """Keras adds every layer's kernel / bias regulariser to the reported loss: l1 * sum|w| or l2 * sum w^2 over every
        ParameterNet kernel and bias, last layer included (nif/model.py:107-117, 220-230; siren.py:515-518).  All of them
        live in the flat parameter buffer (its padding entries are zero).  The gradient is folded into the optimiser."""
        l1, l2 = self.net._kernel_regulariser()
        th = self.net._reg_theta()
        if l2:
            return l2 * (th * th).sum()
        if l1:
            return l1 * th.abs().sum()
        return None

    def _activity_terms(self, z: torch.Tensor, B: int, gb: int):
        """activity_regularizer of the last ParameterNet layer (nif/model.py:118-125, 229; siren.py:467-469): Keras adds
        reg(pnet_output) / batch_size with pnet_output = z W_h + b_h the (B, po_dim) tensor.  Returns
        (loss term of this process's rows, dz [B,K], dW_h [K,P], db_h [P]) or None.
        L2: closed form in the Gram matrices  G = zt^T zt  and  C = W W^T  (zt = [z,1], W = [W_h; b_h]) --
            sum_b |zt_b W|^2 = tr(W^T G W),  dW = 2 G W,  dz_b = 2 (C zt_b)[:K] -- nothing of size (B, po_dim) is formed.
        L1: |.| does not factor: row blocks of pnet_output are formed one at a time by library GEMMs."""
        n = self.net
        a1, a2 = n.p_act_l1_reg, n.p_act_l2_reg
        use_l2 = isinstance(a2, (float, int))
        use_l1 = isinstance(a1, (float, int)) and not use_l2  # l2 wins (nif/model.py:118-125)
        if not (use_l1 or use_l2):
            return None
        W = torch.cat([n.w_h.detach(), n.b_h.detach()[None, :]], 0)  # [K+1, P]
        zt = torch.cat([z, torch.ones_like(z[:, :1])], 1)
        K = z.shape[1]
        # Keras divides by the batch it sees: the per-replica batch B (each replica adds its own term; the replica losses
        # are then summed, and the gradient all-reduce sums the replica gradients, so the scale here is 1 / B per replica
        # times this replica's share B / gb of the global-batch mean -> 1 / gb)
        scale = 1.0 / gb
        if use_l2:
            lam = float(a2)
            G = zt.T @ zt
            GW = G @ W
            term = lam * scale * (W * GW).sum()
            dWt = (2.0 * lam * scale) * GW
            Cm = W @ W.T
            dz = (2.0 * lam * scale) * (zt @ Cm)[:, :K]
        else:
            lam = float(a1)
            term = z.new_zeros(())
            dWt = torch.zeros_like(W)
            dz = torch.empty_like(z)
            step = max(1, (1 << 26) // max(W.shape[1], 1))  # 256 MB of fp32 per block
            for s0 in range(0, B, step):
                blk = zt[s0:s0 + step] @ W
                term = term + blk.abs().sum()
                sg = torch.sign(blk)
                dWt += zt[s0:s0 + step].T @ sg
                dz[s0:s0 + step] = (sg @ W.T)[:, :K]
            term = lam * scale * term
            dWt *= lam * scale
            dz *= lam * scale
        return term, dz.contiguous(), dWt[:K], dWt[K]

    def _apply_update(self, apply_update, l1, l2):
        n = self.net
        if getattr(self.optimizer, "_centralize_in_fit", False):
            from .optimizers import centralize_
            centralize_(n)
        apply_update(n.theta, n.grad, l1, l2)

    def _graph_signature(self):
        n, opt = self.net, self.optimizer
        return (n.theta.data_ptr(), n.grad.data_ptr(), opt._m.data_ptr(), opt._v.data_ptr(), opt._alpha_dev.data_ptr(),
                self._loss_buf.data_ptr(), id(opt), opt.beta_1, opt.beta_2, opt.epsilon, n._kernel_regulariser(),
                self.dist is not None, self._symm_ready())

    def _train_step_graph(self, inp, tgt, sw, gb) -> torch.Tensor:
        n, opt = self.net, self.optimizer
        opt.prepare_replay(n.theta)
        key = (inp.shape[0], gb, sw is not None)
        ent = self._graphs.get(key)
        if ent is not None and ent.get("sig") not in (None, self._graph_signature()):
            ent = None  # parameters / optimiser state were re-created: the recorded pointers are stale
        if (ent is None or ent["sig"] is None) and inp.device.type != "cuda":  # eager visit / recording: device tensors
            inp, tgt, sw = self._dev(inp), self._dev(tgt), (None if sw is None else self._dev(sw))
        if ent is None:
            # first visit of this shape: run it eagerly (this also sizes the persistent workspaces the graph will use)
            while len(self._graphs) >= self.GRAPH_CACHE:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = {"sig": None}
            self._loss_buf.zero_()
            self._fused_step(inp, tgt, sw, gb, opt.apply)
            return self._loss_buf
        if ent["sig"] is None:
            # second visit: record.  (Capture launches nothing; the replay below performs this step.)
            ent["inp"], ent["tgt"] = torch.empty_like(inp), torch.empty_like(tgt)
            ent["sw"] = torch.empty_like(sw) if sw is not None else None
            g = torch.cuda.CUDAGraph()
            if self.dist is None:
                with torch.cuda.graph(g):
                    self._loss_buf.zero_()
                    self._fused_step(ent["inp"], ent["tgt"], ent["sw"], gb, opt.record_apply)
                ent["graph"] = g
            elif self._symm_ready():
                with torch.cuda.graph(g):  # both kernel sequences in one graph; the fused update + barriers stay eager
                    self._loss_buf.zero_()
                    ctx = self._fused_part1(ent["inp"], ent["tgt"], ent["sw"], gb)
                    self._fused_part2(ctx)
                ent["graph"], ent["ctx"] = g, ctx
            else:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._loss_buf.zero_()
                    ctx = self._fused_part1(ent["inp"], ent["tgt"], ent["sw"], gb)
                with torch.cuda.graph(g2, pool=g.pool()):
                    self._fused_part2(ctx)
                ent["graph"], ent["graph2"], ent["ctx"] = g, g2, ctx  # ctx: buffers the second graph reads stay alive
            # buffers whose addresses are baked into the graph stay alive as long as it does
            ent["keep"] = (n.engine._ws, n._trunk._ws, n._trunk._packed, self._packed, self._loss_buf, n.theta, n.grad,
                           opt._m, opt._v, opt._alpha_dev)
            ent["sig"] = self._graph_signature()
        ent["inp"].copy_(inp, non_blocking=True)
        ent["tgt"].copy_(tgt, non_blocking=True)
        if sw is not None:
            ent["sw"].copy_(sw, non_blocking=True)
        if self.dist is None:
            opt.advance_replay()
            ent["graph"].replay()
        elif "graph2" not in ent:
            ent["graph"].replay()
            self._symm_update()
        else:
            self._fused_step(None, None, None, gb, opt.apply, part1=ent["graph"].replay, part2=ent["graph2"].replay)
        return self._loss_buf

    def _general_step_graph(self, inp, tgt, sw, gb) -> torch.Tensor:
        """The general step (trunk as torch ops differentiated by autograd around the fused head kernels: trunks wider than
        64 units, sine / res-block trunks, the last-layer-parameterised class) is a few hundred small launches from Python;
        recorded once per batch shape -- loss, every gradient and, in a single process, the Adam update -- and replayed."""
        n, opt = self.net, self.optimizer
        opt.prepare_replay(n.theta)
        key = ("general", inp.shape[0], gb, sw is not None)
        ent = self._graphs.get(key)
        if ent is not None and ent.get("sig") not in (None, self._graph_signature()):
            ent = None
        single = self.dist is None
        l1, l2 = n._kernel_regulariser()

        def eager_update():
            if self._symm_ready():
                self._symm_update()
            else:
                if self.dist is not None:
                    self.dist.allreduce_(n.grad)
                self._apply_update(opt.apply, l1, l2)

        if ent is None:  # first visit: eager (sizes workspaces, warms up the library handles)
            while len(self._graphs) >= self.GRAPH_CACHE:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = {"sig": None}
            loss = self._loss_and_grad(inp, tgt, self.loss, sw, gb, with_regularisers=True)
            eager_update()
            return loss
        if ent["sig"] is None:
            ent["inp"], ent["tgt"] = torch.empty_like(inp), torch.empty_like(tgt)
            ent["sw"] = torch.empty_like(sw) if sw is not None else None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                ent["out"] = self._loss_and_grad(ent["inp"], ent["tgt"], self.loss, ent["sw"], gb, with_regularisers=True)
                if single:
                    self._apply_update(opt.record_apply, l1, l2)
            ent["graph"] = g
            ent["keep"] = (n.engine._ws if not getattr(n, "_last_layer_only", False) else None, self._packed, self._loss_buf,
                           n.theta, n.grad, opt._m, opt._v, opt._alpha_dev)
            ent["sig"] = self._graph_signature()
        ent["inp"].copy_(inp, non_blocking=True)
        ent["tgt"].copy_(tgt, non_blocking=True)
        if sw is not None:
            ent["sw"].copy_(sw, non_blocking=True)
        if single:
            opt.advance_replay()
            ent["graph"].replay()
        else:
            ent["graph"].replay()
            eager_update()
        return ent["out"]

    # ---- Sobolev training (JacobianLayer inside the loss) ---------------------------------------------------
    def _plan_sobolev(self, loss: "SobolevMSE"):
        """Map the loss columns onto tangent directions (one per differentiated input column) and ShapeNet outputs."""
        n = self.net
        ny, nx = len(self.jac_y), len(self.jac_x)
        general = not isinstance(loss, SobolevMSE)
        if general:  # every Jacobian entry of the model output may enter the loss
            grad_cols = list(range(n.so_dim, n.so_dim + ny * nx))
        else:
            grad_cols = loss.grad_cols
            for c in loss.value_cols:
                if not 0 <= c < n.so_dim:
                    raise NifError(f"value column {c} is not one of the {n.so_dim} model outputs")
        dirs, pairs = [], []  # input column of each direction; (direction, output) of each grad column
        for c in grad_cols:
            k = c - n.so_dim
            if not 0 <= k < ny * nx:
                raise NifError(f"grad column {c} is outside the Jacobian block of the model output")
            a, cc = divmod(k, nx)
            col = self.jac_x[cc]
            if not 0 <= col < n.pi_dim + n.si_dim:
                raise NifError(f"x_index {col} outside the {n.pi_dim + n.si_dim} model inputs")
            if col not in dirs:
                dirs.append(col)
            pairs.append((dirs.index(col), self.jac_y[a]))
        if not dirs:
            raise NifError("SobolevMSE needs at least one grad column")
        from . import _lib
        if len(dirs) > _lib.NIF_MAX_DIR:
            raise NifError(f"a loss on a JacobianLayer model differentiates at most {_lib.NIF_MAX_DIR} inputs")
        return {"dirs": dirs, "pairs": pairs, "grad_cols": grad_cols, "general": general,
                "latent_moves": any(c < n.pi_dim for c in dirs)}

    def _train_step_sobolev(self, inp: torch.Tensor, tgt: torch.Tensor, global_batch: Optional[int]) -> torch.Tensor:
        """One optimisation step of a JacobianLayer model: loss and gradients (_sobolev_loss_and_grad), then the update.
        With the latent code at rest (ShapeNet-input directions only) the step is recorded in a CUDA graph per batch
        shape and replayed, like the general step."""
        n = self.net
        B = inp.shape[0]
        if (B > 0 and inp.is_cuda and not self._sobolev_plan["latent_moves"] and self._graph_enabled()
                and isinstance(self.loss, (str, SobolevMSE))  # (a user callable may do anything, e.g. synchronise)
                and os.environ.get("NIF_B200_GRAPH_GENERAL", "1") != "0"):
            return self._sobolev_step_graph(inp, tgt, int(global_batch) if global_batch else B)
        lv = self._sobolev_loss_and_grad(inp, tgt, global_batch)
        if self.dist is not None:
            self.dist.allreduce_(n.grad)
        l1, l2 = n._kernel_regulariser()
        self._apply_update(self.optimizer.apply, l1, l2)
        return lv

    def _sobolev_step_graph(self, inp, tgt, gb) -> torch.Tensor:
        n, opt = self.net, self.optimizer
        if self._loss_buf is None:
            self._loss_buf = torch.zeros(1, dtype=torch.float32, device=n.device)
        opt.prepare_replay(n.theta)
        key = ("sobolev", inp.shape[0], gb)
        ent = self._graphs.get(key)
        if ent is not None and ent.get("sig") not in (None, self._graph_signature()):
            ent = None
        single = self.dist is None
        l1, l2 = n._kernel_regulariser()

        def eager_update():
            if self.dist is not None:
                self.dist.allreduce_(n.grad)
            self._apply_update(opt.apply, l1, l2)

        if ent is None:  # first visit: eager (sizes workspaces, warms up the library handles)
            while len(self._graphs) >= self.GRAPH_CACHE:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = {"sig": None}
            lv = self._sobolev_loss_and_grad(inp, tgt, gb)
            eager_update()
            return lv
        if ent["sig"] is None:
            ent["inp"], ent["tgt"] = torch.empty_like(inp), torch.empty_like(tgt)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                ent["out"] = self._sobolev_loss_and_grad(ent["inp"], ent["tgt"], gb)
                if single:
                    self._apply_update(opt.record_apply, l1, l2)
            ent["graph"] = g
            ent["keep"] = (n.engine._ws, getattr(n._trunk, "_ws", None), getattr(n._trunk, "_packed", None), self._packed,
                           self._loss_buf, n.theta, n.grad, opt._m, opt._v, opt._alpha_dev)
            ent["sig"] = self._graph_signature()
        ent["inp"].copy_(inp, non_blocking=True)
        ent["tgt"].copy_(tgt, non_blocking=True)
        if single:
            opt.advance_replay()
            ent["graph"].replay()
        else:
            ent["graph"].replay()
            eager_update()
        return ent["out"]

    def _sobolev_loss_and_grad(self, inp: torch.Tensor, tgt: torch.Tensor, global_batch: Optional[int]) -> torch.Tensor:
        """JacobianLayer inside the loss (tutorial 8): forward-mode tangents with a stash, then the reverse-over-forward
        pass of the library.  Grad columns w.r.t. ShapeNet inputs are directions on x; grad columns w.r.t. ParameterNet
        inputs (du/dt) are directions on the latent code, whose tangent -- and the adjoint that comes back for it -- go
        through the trunk by reverse-over-forward autograd.  Leaves every gradient in the flat buffer; returns the loss."""
        n, loss, plan = self.net, self.loss, self._sobolev_plan
        if isinstance(n.p_jac_reg, (float, int)):
            raise NifError("jac_reg is not combined with Sobolev training in this build")
        eng = n.engine
        B = inp.shape[0]
        gb = int(global_batch) if global_batch else B
        xs = inp[:, n.pi_dim: n.pi_dim + n.si_dim].contiguous()
        p_in = inp[:, : n.pi_dim].contiguous()
        dirs, pairs = plan["dirs"], plan["pairs"]
        D = len(dirs)
        fused_trunk = n._trunk is not None and not plan["latent_moves"]
        xdot = torch.zeros(D, B, n.si_dim, device=inp.device)
        zdot, zds = None, []
        if fused_trunk:
            z, tstash = n._trunk.forward(p_in, n.theta_trunk, save=True)
        else:
            n.grad.zero_()
            z = None
            if plan["latent_moves"]:
                zdot = torch.zeros(D, B, n.pi_hidden, device=inp.device)
        for d, col in enumerate(dirs):
            if col >= n.pi_dim:
                xdot[d, :, col - n.pi_dim] = 1.0
            else:
                e = torch.zeros_like(p_in)
                e[:, col] = 1.0
                zz, zd = torch.func.jvp(n._latent, (p_in,), (e,))
                z = zz if z is None else z
                zds.append((d, zd))
                zdot[d] = zd.detach()
        if z is None:
            z = n._latent(p_in)
        zc = z.detach().contiguous()
        packed = self._packed_weights()
        u, udot, stash = eng.forward_tangent(zc, xs, packed, zdot, xdot, save=True)
        # seeds of the batch-mean loss (O(B) elementwise; everything heavier is in the library)
        dud = torch.zeros_like(udot)
        if plan["general"]:
            # 'mse' or a callable(y_true, y_pred) on the concatenated output [u | du/dx...] (a PDE residual, say): its
            # derivative with respect to that output is the seed, taken by autograd over the O(B) loss expression
            with torch.enable_grad():
                yp = torch.cat([u] + [udot[d][:, yc: yc + 1] for d, yc in pairs], -1).requires_grad_(True)
                lv = self._mse_torch(tgt, yp, None) if isinstance(loss, str) else loss(tgt, yp)
                if lv.dim() > 0:
                    lv = lv.mean()
                lv = lv * (B / gb)
                (dy,) = torch.autograd.grad(lv, yp)
            lv = lv.detach()
            du = dy[:, : n.so_dim].contiguous()
            for k, (d, yc) in enumerate(pairs):
                dud[d][:, yc] += dy[:, n.so_dim + k]
        else:
            du = torch.zeros_like(u)
            vc, gc = loss.value_cols, loss.grad_cols
            if plan.get("vc_idx") is None or plan["vc_idx"].device != u.device:  # (a device index: no host copy per step)
                plan["vc_idx"] = torch.as_tensor(list(vc), dtype=torch.long, device=u.device)
            vci = plan["vc_idx"]
            ev = u.index_select(1, vci) - tgt.index_select(1, vci)
            du.index_copy_(1, vci, (2.0 / (gb * len(vc))) * ev)
            sq_g = 0.0
            for (d, yc), c in zip(pairs, gc):
                eg = udot[d][:, yc] - tgt[:, c]
                dud[d][:, yc] += (2.0 * loss.coef_grad / (gb * len(gc))) * eg
                sq_g = sq_g + (eg * eg).sum()
            lv = ((ev * ev).sum() / len(vc) + loss.coef_grad * sq_g / len(gc)) / gb
        out = eng.sobolev_backward(zc, xs, xdot, packed, stash, du, dud, n._gviews[n._last_names[0]],
                                   n._gviews[n._last_names[1]], 0.0, zdot=zdot, zdot_dirs=[d for d, _ in zds])
        if fused_trunk:
            n._trunk.backward(p_in, n.theta_trunk, tstash, out, n.grad_trunk, 0.0)
        elif zdot is None:
            z.backward(out)
        else:
            dz, dzdot = out
            torch.autograd.backward([z] + [zd for _, zd in zds], [dz] + [dzdot[d] for d, _ in zds])
        return lv.reshape(1)

    def fit(self, x=None, y=None, batch_size=None, epochs=1, verbose=0, callbacks=None, shuffle=True,
            sample_weight=None, initial_epoch=0, **_kw):
        if self.optimizer is None:
            raise NifError("call compile() before fit()")
        if isinstance(x, Dataset):
            ds = x  # like Keras, batch_size / shuffle arguments are ignored for a dataset
        else:
            arrays = [np.asarray(x), np.asarray(y)] + ([np.asarray(sample_weight)] if sample_weight is not None else [])
            ds = Dataset(arrays).batch(batch_size or 32)
            if shuffle:
                ds.shuffle(arrays[0].shape[0])
        if self.dist is not None:
            ds.shard(self.dist.world, self.dist.rank)
        if self.net.device.type == "cuda" and ds.nbytes() <= self.DEVICE_CACHE_BYTES:
            ds.cache_on(self.net.device)
        ds.pin()
        hist = History()
        cbs: List[Callback] = [hist] + list(callbacks or [])
        for cb in cbs:
            cb.set_model(self)
            cb.on_train_begin({})
        self.stop_training = False
        for epoch in range(initial_epoch, epochs):
            for cb in cbs:
                cb.on_epoch_begin(epoch, {})
            t0 = time.time()
            tot = torch.zeros(1, dtype=torch.float64, device=self.net.device)
            rows = 0
            mtot, mrows = {}, 0
            for bi, (gb, parts) in enumerate(ds.batches(epoch)):
                inp = parts[0].to(self.net.device, non_blocking=True)
                tgt = parts[1].to(self.net.device, non_blocking=True)
                sw = parts[2].to(self.net.device, non_blocking=True).reshape(-1) if len(parts) > 2 else None
                if sw is not None and self.kind == "jacobian":
                    raise NifError("sample_weight is not supported by the Sobolev training step")
                if self.metrics_fns:  # Keras evaluates metrics on the training forward pass (pre-update weights)
                    pred = self._forward_batch(inp)
                    for mfn in self.metrics_fns:
                        k = getattr(mfn, "__name__", type(mfn).__name__)
                        mtot[k] = mtot.get(k, 0.0) + torch.as_tensor(mfn(tgt, pred)).double().mean() * inp.shape[0]
                    mrows += inp.shape[0]
                loss = self._train_step(inp, tgt, sw, gb)
                tot += loss.double() * gb  # Keras reports the sample-weighted running mean of batch losses
                rows += gb
            if self.dist is not None:
                self.dist.allreduce_(tot)
            logs = {"loss": float(tot) / max(rows, 1), "lr": self.optimizer.learning_rate}
            for k, v in mtot.items():
                logs[k] = float(v) / max(mrows, 1)
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs} - {time.time() - t0:.2f}s - loss: {logs['loss']:.4e}")
            for cb in cbs:
                cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        for cb in cbs:
            cb.on_train_end({})
        self.history = hist
        return hist

    def evaluate(self, x, y, batch_size=None, sample_weight=None, **_kw) -> float:
        pred = self.predict(x, batch_size)
        per_row = ((pred - np.asarray(y, np.float32)) ** 2).mean(-1)
        if sample_weight is not None:
            per_row = per_row * np.asarray(sample_weight, np.float32).reshape(-1)
        return float(per_row.mean())

    # ---- checkpoints ------------------------------------------------------------------------------------
    @staticmethod
    def _ckpt_file(path: str) -> str:
        return path if path.endswith(".npz") else path + ".npz"

    def save_weights(self, path: str, save_format: Optional[str] = None):
        """Keras `save_weights(prefix)` (tutorial/2_multi_scale_NIF.ipynb:629).  Default: one .npz keyed by the reference
        variable names.  save_format="tf" writes a TensorFlow checkpoint (`<prefix>.index` + `.data-00000-of-00001`,
        nif_b200.data.tf_checkpoint) whose object graph carries the same names."""
        if save_format == "tf":
            from .data.tf_checkpoint import write_checkpoint
            write_checkpoint(path, self.net.get_weights())
            return
        f = self._ckpt_file(path)
        d = os.path.dirname(f)
        if d:
            os.makedirs(d, exist_ok=True)
        np.savez(f, **{k.replace("/", "|"): v for k, v in self.net.get_weights().items()})

    def load_weights(self, path: str):
        """An .npz written by save_weights, or a TensorFlow checkpoint prefix as written by the reference's
        `model.save_weights("…/ckpt")` (tutorial/1_simple_1d_wave.ipynb:501, 1280): variables are matched by the names
        the reference layers give them (the checkpoint's object graph records them as full_name)."""
        if not path.endswith(".npz") and os.path.exists(path + ".index"):
            from .data.tf_checkpoint import load_variables
            found = load_variables(path)
            named = {k: v for k, v in found.items() if k in self.net.variables}
            missing = [k for k in self.net.variables if k not in named]
            if missing:
                raise NifError(f"checkpoint {path!r} lacks variables {missing[:4]}{'...' if len(missing) > 4 else ''} "
                               f"(it holds {sorted(found)[:6]}...)")
            self.net.set_weights(named)
            return self
        with np.load(self._ckpt_file(path)) as d:
            self.net.set_weights({k.replace("|", "/"): d[k] for k in d.files})
        return self
