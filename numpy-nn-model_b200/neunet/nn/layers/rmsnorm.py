"""nn.RMSNorm: y = x / sqrt(mean(x^2) + eps) * w (+ b)   (reference: neunet/nn/layers/rmsnorm.py:39-94).
On "cuda" forward and backward are one-pass warp-per-row kernels (``neunet.b200``); dw/db are
two-stage column reductions instead of the reference's serial per-column loops (rmsnorm.cu:248-278)."""
from __future__ import annotations

import numpy as np

from ... import ones as _ones, zeros as _zeros
from ...autograd import Tensor
from ..modules import Module
from ..parameter import Parameter


class _RMSNormTensor(Tensor):
    def __init__(self, data, args, op, device):
        t = Tensor._wrap(data, args, op, True, device)
        self.__dict__.update(t.__dict__)
        self.grad_fn = _rmsnorm_grad_fn


def _rmsnorm_grad_fn(X: Tensor, weight: Tensor, bias, X_std, grad):
    if X.device == "cuda":
        from ... import b200
        # the gradient X already holds (residual branch) is accumulated by the same kernel: no separate add
        prev = X.grad if (X.requires_grad and X.grad is not None and tuple(X.grad.shape) == tuple(X.data.shape)) else None
        dx, dw, db, accumulated = b200.rmsnorm_backward(grad, X.data, weight.data, X_std, need_db=bias is not None,
                                                        dx_add=prev)
        if accumulated:
            X.grad = dx
            weight.apply_grad(dw)
            if bias is not None:
                bias.apply_grad(db)
            return
    else:
        n = X.data.shape[-1]
        x = X.data
        dxh = weight.data * grad
        dx = (dxh * X_std - x * np.sum(dxh * x / X_std, axis=-1, keepdims=True) / n) / X_std ** 2
        lead = tuple(range(grad.ndim - 1))
        dw = np.sum(grad * (x / X_std), axis=lead)
        db = np.sum(grad, axis=lead) if bias is not None else None
    X.apply_grad(dx)
    weight.apply_grad(dw)
    if bias is not None:
        bias.apply_grad(db)


class RMSNorm(Module):
    def __init__(self, dim: int, eps: float = 1e-6, device="cpu", bias=False):
        super().__init__()
        self.eps = eps
        self.weight = Parameter(_ones(dim))
        self.bias = Parameter(_zeros(dim)) if bias else None
        self.to(device)

    def forward(self, X: Tensor) -> Tensor:
        b = self.bias
        planes = None
        if X.device == "cuda":
            from ... import b200
            from ...autograd import _pending
            bd = b.data if b is not None else None
            if _pending(X, "add_dropout"):
                # `x = x + dropout(a); norm(x)`: residual add, dropout and norm in one pass; the sum is delivered
                # to the pending add node as a by-product
                O, std, S, planes = b200.rmsnorm_forward(X._f_x.data, self.weight.data, bd, self.eps,
                                                         add_dropout=(X._f_a.data, X._f_p, X._f_ticket), want_planes=True)
                X.data = S
            else:
                O, std, _, planes = b200.rmsnorm_forward(X.data, self.weight.data, bd, self.eps, want_planes=True)
        else:
            std = np.sqrt(np.mean(X.data ** 2, -1, keepdims=True) + self.eps)
            O = X.data / std * self.weight.data
            if b is not None:
                O = O + b.data
        out = _RMSNormTensor(O, (X, self.weight, b, std), "rmsnorm", X.device)
        if planes is not None:
            out._b200_xst = planes  # bf16 operand planes for the nn.Linear layers that read this output
        return out
