"""Optimizers. ``Adam`` and ``AdamW`` keep the reference's update rules and hyper-parameter
defaults (neunet/optim.py:4-37, 39-73). On ``device="cuda"`` ONE multi-tensor kernel updates every
parameter (``neunet.b200.FusedAdam``), replacing the reference's per-tensor loop of ~10 NumPy/CuPy
temporaries; on ``"cpu"`` the update is the NumPy loop. ``grad_scale`` (used by data-parallel
training: 1/world_size after a sum all-reduce) is folded into the same kernel."""
from __future__ import annotations

import numpy as np


class _AdamBase:
    _mode = 0  # b200.OPT_ADAM_L2

    def __init__(self, params, lr, betas, eps, weight_decay):
        self.params = list(params)
        self.lr = lr
        self.betas = betas
        self.eps = eps
        self.weight_decay = weight_decay
        self.m = [p.xp.zeros_like(p.data) for p in self.params]
        self.v = [p.xp.zeros_like(p.data) for p in self.params]
        self.t = 0
        self.grad_scale = 1.0
        self._fused = None
        self._fused_key = None

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    # -- device path ---------------------------------------------------------------------------
    def _device_step(self, ranges=None):
        """ranges: None = one launch over every parameter; else an iterable of (first, count, before) -- the update is
        issued range by range, `before()` (e.g. the wait for that gradient chunk's all-reduce) called ahead of each."""
        from . import b200
        key = tuple(p.data.data_ptr() for p in self.params)
        if self._fused is None or self._fused_key != key:
            for i, p in enumerate(self.params):  # the kernel updates in place: needs owned, dense storage
                if not p.data.is_contiguous():
                    p.data = p.data.contiguous()
            key = tuple(p.data.data_ptr() for p in self.params)
            self._fused = b200.FusedAdam([p.data for p in self.params], self.m, self.v)
            self._fused_key = key
        grads = []
        for p in self.params:
            g = p.grad
            if g is not None:
                if tuple(g.shape) != tuple(p.data.shape):
                    g = g.reshape(tuple(p.data.shape))
                if not g.is_contiguous() or g.dtype != p.data.dtype:
                    g = g.contiguous().to(p.data.dtype)
            grads.append(g)
        self._fused.sync_staging(self.params)  # the kernel also writes the bf16 planes Linear layers read
        if ranges is None:
            self._fused.step(grads, self.lr, self.betas, self.eps, self.weight_decay, self.t, self._mode, self.grad_scale)
        else:
            self._fused.set_grads(grads)
            first_launch = True
            for first, count, before in ranges:
                if before is not None:
                    before()
                self._fused.step_range(first, count, self.lr, self.betas, self.eps, self.weight_decay, self.t, self._mode,
                                       self.grad_scale, advance=first_launch)
                first_launch = False
            b200.weights_changed()
        self._fused.stamp_staging(self.params, grads)

    def step_ranges(self, ranges):
        """One optimizer step issued as several launches over disjoint parameter ranges that together cover every
        parameter (``neunet.distributed.GradBucket.all_reduce_and_step``); same result as ``step()``."""
        self.t += 1
        return self._device_step(ranges)

    def step(self):
        self.t += 1
        if self.params and self.params[0].device == "cuda":
            return self._device_step()
        b1, b2 = self.betas
        for i, p in enumerate(self.params):
            g = p.grad
            if g is None:
                continue
            if self.grad_scale != 1.0:
                g = g * np.float32(self.grad_scale)
            g = self._pre(p, g)
            self.m[i] = b1 * self.m[i] + (1 - b1) * g
            self.v[i] = b2 * self.v[i] + (1 - b2) * g ** 2
            m_hat = self.m[i] / (1 - b1 ** self.t)
            v_hat = self.v[i] / (1 - b2 ** self.t)
            p.data -= self.lr * m_hat / (np.sqrt(v_hat) + self.eps)


class Adam(_AdamBase):
    """L2 regularisation folded into the gradient (optim.py:24-25)."""
    _mode = 0

    def __init__(self, params, lr: float = 0.01, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay=0):
        super().__init__(params, lr, betas, eps, weight_decay)

    def _pre(self, p, g):
        return g + self.weight_decay * p.data if self.weight_decay != 0 else g


class AdamW(_AdamBase):
    """Decoupled weight decay applied to the parameter first (optim.py:59-60)."""
    _mode = 1

    def __init__(self, params, lr: float = 0.01, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01):
        super().__init__(params, lr, betas, eps, weight_decay)

    def _pre(self, p, g):
        if self.weight_decay != 0:
            p.data -= self.lr * self.weight_decay * p.data
        return g


class SGD:
    """Plain SGD (optim.py:76-88); array back-end only."""

    def __init__(self, params, lr: float = 0.01):
        self.params = list(params)
        self.lr = lr

    def step(self):
        for p in self.params:
            if p.grad is None:
                continue
            p.data -= self.lr * p.grad
        if self.params and self.params[0].device == "cuda":
            from . import b200
            b200.weights_changed()

    def zero_grad(self):
        for p in self.params:
            p.grad = None


_OUT_OF_SCOPE = frozenset({"Momentum", "RMSprop", "Adagrad", "Adadelta", "Adamax", "NAdam"})


class _OutOfScope(NotImplementedError, AttributeError):
    """Also an AttributeError, so ``hasattr`` / ``getattr(..., default)`` keep working."""


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise _OutOfScope(
            f"neunet.optim.{name} exists in the reference (neunet/optim.py) but is outside the hot path this package implements "
            "(SGD, Adam, AdamW; DESIGN.md section 8).")
    raise AttributeError(f"module 'neunet.optim' has no attribute {name!r}")
