"""Activations. Swish and Softmax are the two named epilogue ops: on "cuda" they run as
vectorised sm_100a kernels (and Swish can be fused into the Linear GEMM via ``nn.LinearSwish``);
semantics: neunet/nn/activations.py:208-233 (Swish), 437-459 (Softmax), 462-491 (LogSoftmax).
The remaining element-wise activations stay on the array back-end through differentiable ops."""
from __future__ import annotations

import numpy as np

from ..autograd import Tensor
from .modules import Module


class _ActTensor(Tensor):
    def __init__(self, data, args, op, device, grad_fn):
        t = Tensor._wrap(data, args, op, True, device)
        self.__dict__.update(t.__dict__)
        self.grad_fn = grad_fn


def _sigmoid(xp, x):
    return 1 / (1 + xp.exp(-x))


# ---- Swish -----------------------------------------------------------------------------------
def _swish_grad(t: Tensor, f_x, beta, grad):
    if t.device == "cuda":
        from .. import b200
        t.apply_grad(b200.swish_backward(t.data, grad, beta))
    else:
        s = _sigmoid(np, beta * t.data)
        t.apply_grad(grad * (beta * f_x + s * (1 - beta * f_x)))


class Swish(Module):
    def __init__(self, beta=1):
        self.beta = beta

    def forward(self, x: Tensor):
        if x.device == "cuda":
            from .. import b200
            from ..autograd import _pending
            if _pending(x, "linear") and not x._f_act:
                # Swish directly on a not-yet-launched nn.Linear: it becomes the GEMM's epilogue (one kernel, and
                # swish' is folded into the dO staging pass of the backward), exactly like nn.LinearSwish
                return x._f_fuse_swish(float(self.beta))
            f_x = b200.swish_forward(x.data, self.beta)
        else:
            f_x = x.data * _sigmoid(np, self.beta * x.data)
        return _ActTensor(f_x, [x, f_x, self.beta], "swish", x.device, _swish_grad)

    def __call__(self, x):
        return self.forward(x)


# ---- Softmax / LogSoftmax -------------------------------------------------------------------------
def _softmax_grad(t: Tensor, f_x, axis, grad):
    if t.device == "cuda":
        from .. import b200
        t.apply_grad(b200.softmax_backward(f_x, grad, axis))
    else:
        t.apply_grad((grad - (grad * f_x).sum(axis, keepdims=True)) * f_x)


class Softmax(Module):
    def __init__(self, axis=1):
        self.axis = axis

    def forward(self, x: Tensor):
        if x.device == "cuda":
            from .. import b200
            from ..autograd import _Deferred, _pending
            if _pending(x) and x._f_kind in ("where", "div", "matmul") and self.axis in (-1, x.ndim - 1):
                # softmax over pending attention scores stays pending (absorbed by the fused attention kernel)
                axis, held = self.axis, {}

                def thunk():
                    held["y"] = b200.softmax_forward(x.data, axis)
                    return held["y"]

                def grad_fn(t, grad):
                    t.apply_grad(b200.softmax_backward(held["y"] if "y" in held else out.data, grad, axis))
                out = _Deferred.make(thunk, x.shape, [x], "softmax", True, _f_kind="softmax", _f_src=x)
                out.grad_fn = grad_fn
                return out
            f_x = b200.softmax_forward(x.data, self.axis)
        else:
            e = np.exp(x.data - np.max(x.data, axis=self.axis, keepdims=True))
            f_x = e / np.sum(e, axis=self.axis, keepdims=True)
        return _ActTensor(f_x, [x, f_x, self.axis], "softmax", x.device, _softmax_grad)

    def __call__(self, x):
        return self.forward(x)


def _log_softmax_grad(t: Tensor, f_x, axis, grad):
    xp = t.xp
    t.apply_grad(grad - xp.exp(f_x) * xp.sum(grad, axis=axis, keepdims=True))


class LogSoftmax(Module):
    def __init__(self, axis=1):
        self.axis = axis

    def forward(self, x: Tensor):
        xp = x.xp
        m = xp.max(x.data, axis=self.axis, keepdims=True)
        f_x = x.data - m - xp.log(xp.sum(xp.exp(x.data - m), axis=self.axis, keepdims=True))
        return _ActTensor(f_x, [x, f_x, self.axis], "log_softmax", x.device, _log_softmax_grad)

    def __call__(self, x):
        return self.forward(x)


# ---- simple element-wise activations (array back-end) ----------------------------------------------
class _Elementwise(Module):
    def __call__(self, x):
        return self.forward(x)


class Sigmoid(_Elementwise):
    def forward(self, x: Tensor):
        f = _sigmoid(x.xp, x.data)
        return _ActTensor(f, [x, f], "sigmoid", x.device, lambda t, f_x, grad: t.apply_grad(grad * f_x * (1 - f_x)))


class Tanh(_Elementwise):
    def forward(self, x: Tensor):
        f = x.xp.tanh(x.data)
        return _ActTensor(f, [x, f], "tanh", x.device, lambda t, f_x, grad: t.apply_grad(grad * (1 - f_x ** 2)))


class ReLU(_Elementwise):
    def forward(self, x: Tensor):
        f = x.xp.maximum(x.data, 0)
        return _ActTensor(f, [x], "relu", x.device, lambda t, grad: t.apply_grad(grad * (t.data > 0)))


class LeakyReLU(_Elementwise):
    def __init__(self, alpha=0.01):
        self.alpha = alpha

    def forward(self, x: Tensor):
        xp, a = x.xp, self.alpha
        grad_fn = lambda t, alpha, grad: t.apply_grad(grad * xp.where(t.data <= 0, alpha, 1))  # noqa: E731
        if x.device == "cuda" and x.ndim == 4:
            from ..autograd import _Deferred, fusion_enabled
            if fusion_enabled():
                # pending: a BatchNorm2d that follows absorbs the activation into its kernels (DDPM ResBlock)
                out = _Deferred.make(lambda: xp.where(x.data <= 0, a * x.data, x.data), x.shape, [x, a], "leaky_relu", True,
                                     _f_kind="leaky_relu", _f_src=x, _f_alpha=float(a))
                out.grad_fn = grad_fn
                return out
        f = xp.where(x.data <= 0, a * x.data, x.data)
        return _ActTensor(f, [x, a], "leaky_relu", x.device, grad_fn)


class GELU(_Elementwise):
    def forward(self, x: Tensor):
        c = float(np.sqrt(2 / np.pi))
        inner = (x + x ** 3 * 0.044715) * c
        return x * 0.5 * (inner.tanh() + 1)


class Softplus(_Elementwise):
    def forward(self, x: Tensor):
        return (x.exp() + 1).log()
