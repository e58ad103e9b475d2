"""Mirror of `nif.optimizers` (nif/optimizers/__init__.py:1-20): AdaBeliefOptimizer, Lion, gradient centralisation and the
L-BFGS fine-tuner, over the flat fp32 parameter / gradient buffers of a nif_b200 model.

The update rules run as one CUDA kernel each in libnif_b200.so (nif_optim.cu); step-dependent scalars are formed here in
Python floats, as Keras forms them once per step.  L4Adam is not built.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, Optional

import numpy as np
import torch

from . import _lib
from ._lib import NifError, check

__all__ = ["AdaBeliefOptimizer", "Lion", "centralized_gradients_for_optimizer", "TFPLBFGS", "function_factory", "L4Adam"]


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _FlatOptimizer:
    """What Model.compile() needs from an optimiser: `.learning_rate`, `.iterations`, `.apply(theta, grad, l1, l2, g_scale)`."""

    learning_rate: float
    iterations: int = 0
    get_gradients = None  # tutorials assign centralized_gradients_for_optimizer(...) here (see that function)

    @property
    def lr(self):
        return self.learning_rate

    @lr.setter
    def lr(self, v):
        self.learning_rate = float(v)

    def _slots(self, theta, names):
        for nm in names:
            t = getattr(self, nm, None)
            if t is None or t.shape != theta.shape or t.device != theta.device:
                setattr(self, nm, torch.zeros_like(theta))


class AdaBeliefOptimizer(_FlatOptimizer):
    """AdaBelief with optional rectification, AMSGrad, decoupled weight decay and linear warm-up / decay of the learning
    rate (nif/optimizers/external_optimizers.py:321-628; update rule :458-528; defaults :398-413)."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-14, weight_decay=0.0, rectify=True,
                 amsgrad=False, sma_threshold=5.0, total_steps=0, warmup_proportion=0.1, min_lr=0.0,
                 name="AdaBeliefOptimizer", print_change_log=True, **kwargs):
        self.learning_rate = float(kwargs.get("lr", learning_rate))
        self.beta_1, self.beta_2 = float(beta_1), float(beta_2)
        self.epsilon = float(epsilon) if epsilon else 1e-7  # `epsilon or tf.keras.backend.epsilon()` (:428)
        self.weight_decay = float(weight_decay)
        self.rectify, self.amsgrad = bool(rectify), bool(amsgrad)
        self.sma_threshold = float(sma_threshold)
        self.total_steps, self.warmup_proportion, self.min_lr = int(total_steps), float(warmup_proportion), float(min_lr)
        self.name = name
        self.iterations = 0
        self._m = self._v = self._vhat = None

    def _lr_t(self, step: int) -> float:
        """:468-479: linear warm-up to lr over total_steps * warmup_proportion steps, then linear decay to min_lr."""
        lr_t = self.learning_rate
        if self.total_steps > 0:
            warmup_steps = self.total_steps * self.warmup_proportion
            decay_steps = max(self.total_steps - warmup_steps, 1)
            decay_rate = (self.min_lr - lr_t) / decay_steps
            if step <= warmup_steps:
                lr_t = lr_t * (step / warmup_steps)
            else:
                lr_t = lr_t + decay_rate * min(step - warmup_steps, decay_steps)
        return lr_t

    def apply(self, theta: torch.Tensor, grad: torch.Tensor, l1=0.0, l2=0.0, g_scale=1.0):
        self._slots(theta, ["_m", "_v"] + (["_vhat"] if self.amsgrad else []))
        self.iterations += 1
        t = self.iterations
        check(_lib.lib().nif_adabelief_step(theta.numel(), _ptr(theta), _ptr(grad), _ptr(self._m), _ptr(self._v),
                                            _ptr(self._vhat) if self.amsgrad else None, float(self._lr_t(t)), self.beta_1,
                                            self.beta_2, self.epsilon, t, 1 if self.rectify else 0, self.sma_threshold,
                                            self.weight_decay, float(l1), float(l2), float(g_scale), _stream()),
              "nif_adabelief_step")


class Lion(_FlatOptimizer):
    """nif/optimizers/external_optimizers.py:631-735: p -= lr (sign(b1 m + (1-b1) g) + wd p); m = b2 m + (1-b2) g."""

    def __init__(self, learning_rate=1e-4, beta_1=0.9, beta_2=0.99, wd=0, name="lion", **kwargs):
        self.learning_rate = float(kwargs.get("lr", learning_rate))
        self.beta_1, self.beta_2, self.wd = float(beta_1), float(beta_2), float(wd)
        self.name = name
        self.iterations = 0
        self._m = None

    def apply(self, theta: torch.Tensor, grad: torch.Tensor, l1=0.0, l2=0.0, g_scale=1.0):
        self._slots(theta, ["_m"])
        self.iterations += 1
        check(_lib.lib().nif_lion_step(theta.numel(), _ptr(theta), _ptr(grad), _ptr(self._m), self.learning_rate, self.beta_1,
                                       self.beta_2, self.wd, float(l1), float(l2), float(g_scale), _stream()), "nif_lion_step")


class L4Adam(_FlatOptimizer):
    def __init__(self, *a, **k):
        raise NotImplementedError("L4Adam (nif/optimizers/external_optimizers.py:18-318) is not built in nif_b200")


# --------------------------------------------------------------------------------------------------------------------
def centralize_(net) -> None:
    """Gradient centralisation of every rank >= 2 variable of `net`, in place in its flat gradient buffer
    (nif/optimizers/gtcf.py:27-32: grad -= mean over every axis but the last)."""
    for name, g in net._gviews.items():
        if g.dim() > 1:
            cols = g.shape[-1]
            check(_lib.lib().nif_centralize_gradient(g.numel() // cols, cols, _ptr(g), _stream()), "nif_centralize_gradient")


def centralized_gradients_for_optimizer(optimizer, apply_in_fit: bool = False) -> Callable:
    """nif/optimizers/gtcf.py:53-67.  Returns `get_centralized_gradients(loss, params)` bound to `optimizer`, the function the
    tutorials assign to `optimizer.get_gradients` (tutorial/1_simple_1d_wave.ipynb:767-768).

    Under TF 2 that assignment does not change training: Keras' train_step differentiates with a GradientTape and never
    calls `Optimizer.get_gradients`, so the reference's `fit()` runs with plain gradients.  The same holds here by
    default.  `apply_in_fit=True` opts in to what the tutorial intends: the model then centralises its gradient buffer
    (one kernel per matrix) before every optimiser update."""

    def get_centralized_gradients(loss, params):
        grads = list(torch.autograd.grad(loss, list(params)))
        out = []
        for g in grads:
            if g.dim() > 1:
                g = g - g.mean(dim=tuple(range(g.dim() - 1)), keepdim=True)
            out.append(g)
        return out

    optimizer._centralize_in_fit = bool(apply_in_fit)
    return get_centralized_gradients


# --------------------------------------------------------------------------------------------------------------------
def function_factory(model, loss, train_x, train_y, display_epoch):
    """nif/optimizers/lbfgs.py:7-95: f(params_1d) -> (loss, gradient) of `loss(model(train_x), train_y)` on the full data
    set, with `.history`, `.iter`, `.assign_new_model_parameters`.  The model's variables already live in one flat buffer,
    so the stitch / partition bookkeeping of the reference is the identity."""
    net = model.net
    x = model._dev(train_x)
    y = model._dev(train_y)

    def assign_new_model_parameters(params_1d):
        with torch.no_grad():
            net.theta.copy_(params_1d)

    def f(params_1d):
        assign_new_model_parameters(params_1d)
        lv = model._loss_and_grad(x, y, loss)
        f.iter += 1
        if display_epoch and f.iter % display_epoch == 0:
            print("Epoch:", f.iter, "loss:", float(lv))
        f.history.append(float(lv))
        return lv, net.grad.clone()

    f.iter = 0
    f.history = []
    f.assign_new_model_parameters = assign_new_model_parameters
    return f


class TFPLBFGS(object):
    """nif/optimizers/lbfgs.py:98-126.  The reference hands `function_factory`'s closure to
    tensorflow_probability's `lbfgs_minimize` (third-party, not under /root/reference) with num_correction_pairs=20,
    tolerance = x_tolerance = f_relative_tolerance = 1e-15, max_line_search_iterations=100.  Here: the two-loop L-BFGS
    recursion with a strong-Wolfe bracketing / zoom line search (Nocedal & Wright, Alg. 7.4, 3.5, 3.6; c1 = 1e-4, c2 = 0.9)
    under the same limits -- the same family of iterates, not TFP's Hager-Zhang step lengths."""

    def __init__(self, model, loss_fun, inps, outs, display_epoch=1):
        self.func = function_factory(model, loss_fun, inps, outs, display_epoch)
        self.model = model

    def minimize(self, rounds=50, max_iter=50):
        for _ in range(rounds):
            x = lbfgs_minimize(self.func, self.model.net.theta.detach().clone(), num_correction_pairs=20, tolerance=1e-15,
                               x_tolerance=1e-15, f_relative_tolerance=1e-15, max_iterations=max_iter,
                               max_line_search_iterations=100)
            self.func.assign_new_model_parameters(x)

    @property
    def history(self):
        history = list(self.func.history)
        return {"iteration": np.arange(1, len(history) + 1), "loss": history}


def lbfgs_minimize(value_and_gradients_function, initial_position, num_correction_pairs=20, tolerance=1e-8,
                   x_tolerance=0.0, f_relative_tolerance=0.0, max_iterations=50, max_line_search_iterations=50):
    """L-BFGS on a flat device vector; returns the final position."""
    fg = value_and_gradients_function
    x = initial_position.clone()
    f, g = fg(x)
    f = float(f)
    S, Y, RHO = [], [], []
    for it in range(max_iterations):
        if float(g.abs().max()) <= tolerance:
            break
        # two-loop recursion
        q = g.clone()
        alphas = []
        for s, y, rho in zip(reversed(S), reversed(Y), reversed(RHO)):
            a = rho * float(torch.dot(s, q))
            alphas.append(a)
            q.add_(y, alpha=-a)
        if S:
            q.mul_(float(torch.dot(S[-1], Y[-1])) / float(torch.dot(Y[-1], Y[-1])))
        for (s, y, rho), a in zip(zip(S, Y, RHO), reversed(alphas)):
            b = rho * float(torch.dot(y, q))
            q.add_(s, alpha=a - b)
        d = -q
        gtd = float(torch.dot(g, d))
        if not (gtd < 0):  # not a descent direction (stale curvature): restart from steepest descent
            S, Y, RHO = [], [], []
            d = -g
            gtd = float(torch.dot(g, d))
        t0 = 1.0 if S else min(1.0, 1.0 / max(float(g.abs().sum()), 1e-30))
        t, f_new, g_new, ok = _strong_wolfe(fg, x, f, g, d, gtd, t0, max_line_search_iterations)
        if not ok:
            break
        s = d * t
        y = g_new - g
        x_new = x + s
        ys = float(torch.dot(y, s))
        if ys > 1e-30:
            S.append(s); Y.append(y); RHO.append(1.0 / ys)
            if len(S) > num_correction_pairs:
                S.pop(0); Y.pop(0); RHO.pop(0)
        dx = float(s.abs().max())
        df = abs(f_new - f)
        x, f, g = x_new, f_new, g_new
        if dx <= x_tolerance or df <= f_relative_tolerance * max(abs(f), 1e-300):
            break
    return x


def _strong_wolfe(fg, x, f0, g0, d, gtd0, t, max_evals, c1=1e-4, c2=0.9):
    """Bracketing + zoom (Nocedal & Wright Alg. 3.5 / 3.6).  Returns (step, f, g, success)."""
    def phi(a):
        fv, gv = fg(x + d * a)
        return float(fv), gv, float(torch.dot(gv, d))

    evals = 0
    a_prev, f_prev, gtd_prev, g_prev = 0.0, f0, gtd0, g0
    a = t
    lo = hi = None
    while evals < max_evals:
        f_a, g_a, gtd_a = phi(a)
        evals += 1
        if not math.isfinite(f_a) or f_a > f0 + c1 * a * gtd0 or (evals > 1 and f_a >= f_prev):
            lo, hi = (a_prev, f_prev, gtd_prev, g_prev), (a, f_a, gtd_a, g_a)
            break
        if abs(gtd_a) <= -c2 * gtd0:
            return a, f_a, g_a, True
        if gtd_a >= 0:
            lo, hi = (a, f_a, gtd_a, g_a), (a_prev, f_prev, gtd_prev, g_prev)
            break
        a_prev, f_prev, gtd_prev, g_prev = a, f_a, gtd_a, g_a
        a *= 2.0
    if lo is None:
        return a_prev, f_prev, g_prev, a_prev > 0
    while evals < max_evals:
        a = 0.5 * (lo[0] + hi[0])  # bisection inside the bracket
        f_a, g_a, gtd_a = phi(a)
        evals += 1
        if not math.isfinite(f_a) or f_a > f0 + c1 * a * gtd0 or f_a >= lo[1]:
            hi = (a, f_a, gtd_a, g_a)
        else:
            if abs(gtd_a) <= -c2 * gtd0:
                return a, f_a, g_a, True
            if gtd_a * (hi[0] - lo[0]) >= 0:
                hi = lo
            lo = (a, f_a, gtd_a, g_a)
        if abs(hi[0] - lo[0]) < 1e-16 * max(1.0, abs(lo[0])):
            break
    if lo[0] > 0 and lo[1] < f0:
        return lo[0], lo[1], lo[3], True
    return 0.0, f0, g0, False
