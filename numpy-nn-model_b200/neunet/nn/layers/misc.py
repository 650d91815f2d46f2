"""Layers that are NOT on the dense hot path but that the named examples need so they run
unchanged: Embedding, Dropout, BatchNorm2d, MaxPool2d, Flatten, LayerNorm.
They stay on the array back-end (NumPy / torch element-wise ops), as the reference keeps them on
NumPy / CuPy. Semantics follow neunet/nn/layers/{embedding,dropout,batchnorm2d,maxpool2d,flatten,
layernorm}.py, including the inverted BatchNorm momentum (batchnorm2d.py:87-88) and Embedding's
assignment-style backward through ``__getitem__`` (autograd.py:909-910)."""
from __future__ import annotations

import numpy as np

from ... import tensor as _tensor
from ...autograd import Tensor
from ..modules import Module
from ..parameter import Parameter


class Embedding(Module):
    def __init__(self, num_embeddings: int, embedding_dim: int, device="cpu"):
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.weight = Parameter(_tensor(np.random.randn(num_embeddings, embedding_dim), dtype=np.float32))
        self.to(device)

    def forward(self, X: Tensor) -> Tensor:
        idx = X.data if isinstance(X, Tensor) else X
        if isinstance(idx, np.ndarray):
            idx = idx.astype(np.int32)
        if self.weight.device == "cuda":
            from ... import b200
            from ...backend import is_device_array
            ids = idx if is_device_array(idx) else self.weight.xp.array(np.asarray(idx), dtype=np.int64)
            if ids.dtype in (b200.torch.int32, b200.torch.int64) and self.weight.data.ndim == 2:
                # gather kernel forward; backward keeps the reference's assignment semantics of weight[idx]
                # (autograd.py:909-910: the LAST duplicate id wins) with a deterministic two-pass scatter
                out = _StaticTensor(b200.embedding_forward(self.weight.data, ids), (self.weight, ids), "embedding", "cuda",
                                    _embedding_grad)
                return out
        return self.weight[idx]

    def __call__(self, X):
        return self.forward(X)


def _embedding_grad(weight: Tensor, ids, grad):
    from ... import b200
    if not weight.requires_grad:
        return
    buf = getattr(weight, "_grad_buffer", None) if weight.grad is None else None  # data-parallel bucket slice
    weight.apply_grad(b200.embedding_backward(ids, grad, weight.data.shape[0], out=buf))


class _StaticTensor(Tensor):
    """Generic result tensor carrying a hand-written grad_fn."""

    def __init__(self, data, args, op, device, grad_fn):
        t = Tensor._wrap(data, args, op, True, device)
        self.__dict__.update(t.__dict__)
        self.grad_fn = grad_fn


def _dropout_grad(X: Tensor, mask, grad):
    if isinstance(mask, _DeviceMask):
        from ...autograd import _apply_dropped_grad
        _apply_dropped_grad(X, grad, mask.p, mask.ticket)
        return
    X.apply_grad(grad * mask)


class _DeviceMask:
    """Stands in for the saved mask array of dropout.py:31 on "cuda": the mask is regenerated from
    its Philox ticket in backward (neunet/b200: dropout_ticket), never stored."""
    __slots__ = ("p", "ticket")

    def __init__(self, p, ticket):
        self.p, self.ticket = p, ticket


class Dropout(Module):
    def __init__(self, p: float = 0.5):
        self.p = p
        self.scale = 1 / (1 - p)
        self.training = True

    def forward(self, X: Tensor) -> Tensor:
        if not isinstance(X, Tensor):
            raise TypeError("Input must be a tensor")
        if self.training and X.device == "cuda" and 0 <= self.p < 1:
            from ... import b200
            from ...autograd import _Deferred, _pending, fusion_enabled
            mask = _DeviceMask(self.p, b200.dropout_ticket())  # the ticket is taken NOW: call order fixes the masks
            if fusion_enabled():
                # pending: `x + dropout(a)`, `norm(x + dropout(a))` and attention absorb it into their kernel
                p, ticket, planes_ok = self.p, mask.ticket, X.ndim in (2, 3)

                def thunk():
                    if not planes_ok:
                        return b200.dropout_apply(X.data, p, ticket)
                    if _pending(X, "linear_swish"):
                        # dropout(swish(linear(x))): one pass over the GEMM's pre-activation (nn/layers/linear.py: fuse_swish)
                        y, planes = X._f_fuse_dropout(p, ticket)
                    else:
                        y, planes = b200.dropout_apply(X.data, p, ticket, want_planes=True)
                    out._b200_xst = planes  # bf16 operand planes for a following nn.Linear
                    return y
                out = _Deferred.make(thunk, X.shape, (X, mask), "dropout", True, _f_kind="dropout", _f_src=X, _f_p=p,
                                     _f_ticket=ticket)
                out.grad_fn = _dropout_grad
                return out
            return _StaticTensor(b200.dropout_apply(X.data, self.p, mask.ticket), (X, mask), "dropout", X.device,
                                 _dropout_grad)
        if not self.training and X.device == "cuda":
            from ...autograd import _Deferred, _pending
            if _pending(X):
                # eval mode: identity; a pending input (attention probabilities) stays pending so the fused kernel still applies
                out = _Deferred.make(lambda: X.data, X.shape, (X, 1), "dropout", True, _f_kind="dropout", _f_src=X, _f_p=0.0,
                                     _f_ticket=None)
                out.grad_fn = _dropout_grad
                return out
        if self.training:
            mask = X.xp.random.binomial(1, 1 - self.p, size=tuple(X.data.shape))
            mask = (mask.astype(np.float32) if isinstance(mask, np.ndarray) else mask) * self.scale
        else:
            mask = 1
        return _StaticTensor(X.data * mask, (X, mask), "dropout", X.device, _dropout_grad)

    def __call__(self, X):
        return self.forward(X)

    def train(self, mode=True):
        self.training = mode

    def eval(self):
        self.training = False


def _bn2d_grad(X: Tensor, weight, bias, xc, inv, affine, grad):
    xp = X.xp
    n = X.data.shape[0] * X.data.shape[2] * X.data.shape[3]
    ax = (0, 2, 3)
    inv4 = inv.reshape(1, -1, 1, 1)
    xhat = xc * inv4
    w4 = weight.data.reshape(1, -1, 1, 1) if affine else 1
    dxh = w4 * grad
    # d/dx of (x - mean) * inv with batch statistics (batchnorm2d.py:20-43)
    s1 = xp.sum(dxh, axis=ax, keepdims=True)
    s2 = xp.sum(dxh * xhat, axis=ax, keepdims=True)
    X.apply_grad(inv4 * (dxh - s1 / n - xhat * s2 / n))
    if affine:
        weight.apply_grad(xp.sum(grad * xhat, axis=ax).reshape(tuple(weight.data.shape)))
        bias.apply_grad(xp.sum(grad, axis=ax).reshape(tuple(bias.data.shape)))


class BatchNorm2d(Module):
    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1, affine: bool = True, device="cpu"):
        self.num_features = num_features
        self.eps = eps
        self.momentum = momentum
        self.affine = affine
        self.running_mean = Parameter(_tensor(np.zeros((1, num_features)), dtype=np.float32), requires_grad=False)
        self.running_var = Parameter(_tensor(np.ones((1, num_features)), dtype=np.float32), requires_grad=False)
        self.weight = Parameter(_tensor(np.ones((1, num_features)), dtype=np.float32)) if affine else None
        self.bias = Parameter(_tensor(np.zeros((1, num_features)), dtype=np.float32)) if affine else None
        self.training = True
        self.to(device)

    def forward(self, X: Tensor) -> Tensor:
        if not isinstance(X, Tensor):
            raise TypeError("Input must be a tensor")
        if X.device != self.device:
            raise ValueError("Tensors must be on the same device")
        if self.device == "cuda" and X.ndim == 4:
            return self._forward_native(X)
        xp = X.xp
        if self.training:
            mean = xp.mean(X.data, axis=(0, 2, 3))
            var = xp.var(X.data, axis=(0, 2, 3))
            # reference convention: momentum weights the OLD value (batchnorm2d.py:87-88)
            self.running_mean.data = self.momentum * self.running_mean.data + (1 - self.momentum) * mean
            self.running_var.data = self.momentum * self.running_var.data + (1 - self.momentum) * var
        else:
            mean = self.running_mean.data.reshape(-1)
            var = self.running_var.data.reshape(-1)
        xc = X.data - mean.reshape(1, -1, 1, 1)
        inv = 1 / xp.sqrt(var + self.eps)
        O = xc * inv.reshape(1, -1, 1, 1)
        if self.affine:
            O = self.weight.data.reshape(1, -1, 1, 1) * O + self.bias.data.reshape(1, -1, 1, 1)
        return _StaticTensor(O, (X, self.weight, self.bias, xc, inv, self.affine), "batchnorm2d", self.device, _bn2d_grad)

    def _forward_native(self, X: Tensor) -> Tensor:
        """"cuda": fused kernels (neunet.b200.bn_forward). A LeakyReLU whose result is still pending is absorbed:
        the conv -> LeakyReLU -> BatchNorm2d group of the DDPM ResBlock is statistics + one normalising pass."""
        from ... import b200
        from ...autograd import _pending
        src, alpha = X, 1.0
        if _pending(X, "leaky_relu"):
            src, alpha = X._f_src, X._f_alpha
        w = self.weight.data.reshape(-1) if self.affine else None
        b = self.bias.data.reshape(-1) if self.affine else None
        if self.training:
            rm, rv = self.running_mean.data, self.running_var.data
            if not (rm.is_contiguous() and rv.is_contiguous()):
                rm, rv = rm.contiguous(), rv.contiguous()
                self.running_mean.data, self.running_var.data = rm, rv
            O, mean, inv = b200.bn_forward(src.data, w, b, alpha, self.eps, self.momentum, rm, rv)
        else:
            mean = self.running_mean.data.reshape(-1).contiguous()
            inv = 1 / (self.running_var.data.reshape(-1) + self.eps).sqrt()
            O, mean, inv = b200.bn_forward(src.data, w, b, alpha, self.eps, self.momentum, stats=(mean, inv))
        return _StaticTensor(O, (src, self.weight, self.bias, mean, inv, self.affine, alpha, self.training), "batchnorm2d",
                             self.device, _bn2d_native_grad)

    def __call__(self, X):
        return self.forward(X)

    def train(self, mode=True):
        self.training = mode

    def eval(self):
        self.training = False


def _bn2d_native_grad(X: Tensor, weight, bias, mean, inv, affine, alpha, training, grad):
    from ... import b200
    # NB: like the reference (batchnorm2d.py:11-55) the SAME backward formula -- with its batch-sum terms -- is applied in
    # eval mode, where mean / inv come from the running statistics
    w = weight.data.reshape(-1) if affine else None
    dx, dw, db = b200.bn_backward(X.data, grad, mean, inv, w, alpha, need_dx=X.requires_grad, need_dw=affine)
    if dx is not None:
        X.apply_grad(dx)
    if affine:
        weight.apply_grad(dw.reshape(tuple(weight.data.shape)))
        bias.apply_grad(db.reshape(tuple(bias.data.shape)))


def _pair(v):
    return v if isinstance(v, tuple) else (v, v)


def _maxpool_grad(X: Tensor, idx, out_shape, grad):
    xp = X.xp
    B, C, H, W = X.data.shape
    flat = xp.zeros((B * C, H * W), dtype=np.float32)
    g = grad.reshape(B * C, -1)
    if X.device == "cuda":
        flat.scatter_add_(1, idx.reshape(B * C, -1), g)
    else:
        np.add.at(flat, (np.arange(B * C)[:, None], idx.reshape(B * C, -1)), g)
    X.apply_grad(flat.reshape(B, C, H, W))


class MaxPool2d(Module):
    def __init__(self, kernel_size, stride=None, padding=0, dilation=1):
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride) if stride else self.kernel_size
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)

    def forward(self, X: Tensor) -> Tensor:
        B, C, H, W = X.shape
        kh, kw = self.kernel_size
        s, p, d = self.stride, self.padding, self.dilation
        ho = (H + 2 * p[0] - d[0] * (kh - 1) - 1) // s[0] + 1
        wo = (W + 2 * p[1] - d[1] * (kw - 1) - 1) // s[1] + 1
        if X.device == "cuda":
            import torch.nn.functional as F
            out, idx = F.max_pool2d(X.data, self.kernel_size, s, p, d, return_indices=True)
        else:
            xpad = np.pad(X.data, ((0, 0), (0, 0), (p[0], p[0]), (p[1], p[1])), constant_values=-np.inf)
            out = np.full((B, C, ho, wo), -np.inf, dtype=np.float32)
            idx = np.zeros((B, C, ho, wo), dtype=np.int64)
            oy, ox = np.arange(ho)[:, None] * s[0], np.arange(wo)[None, :] * s[1]
            for k in range(kh):
                for l in range(kw):
                    win = xpad[:, :, k * d[0]: k * d[0] + (ho - 1) * s[0] + 1: s[0], l * d[1]: l * d[1] + (wo - 1) * s[1] + 1: s[1]]
                    better = win > out
                    out = np.where(better, win, out)
                    flat = (oy + k * d[0] - p[0]) * W + (ox + l * d[1] - p[1])
                    idx = np.where(better, flat[None, None], idx)
        return _StaticTensor(out, (X, idx, (ho, wo)), "maxpool2d", X.device, _maxpool_grad)

    def __call__(self, X):
        return self.forward(X)


class Flatten(Module):
    def __init__(self, start_dim=1, end_dim=-1):
        self.start_dim, self.end_dim = start_dim, end_dim

    def forward(self, X: Tensor) -> Tensor:
        shape = X.shape
        end = self.end_dim % len(shape)
        return X.reshape(*shape[: self.start_dim], int(np.prod(shape[self.start_dim: end + 1])), *shape[end + 1:])

    def __call__(self, X):
        return self.forward(X)


class LayerNorm(Module):
    """y = (x - mean) / sqrt(var + eps) * w + b over the trailing `normalized_shape` dims, written
    with differentiable Tensor ops (dynamic backward)."""

    def __init__(self, normalized_shape, eps: float = 1e-5, elementwise_affine: bool = True, bias: bool = True, device="cpu"):
        self.normalized_shape = (normalized_shape,) if isinstance(normalized_shape, int) else tuple(normalized_shape)
        self.eps = eps
        self.weight = Parameter(_tensor(np.ones(self.normalized_shape), dtype=np.float32)) if elementwise_affine else None
        self.bias = Parameter(_tensor(np.zeros(self.normalized_shape), dtype=np.float32)) if elementwise_affine and bias else None
        self.to(device)

    def forward(self, X: Tensor) -> Tensor:
        axes = tuple(range(-len(self.normalized_shape), 0))
        mean = X.mean(axis=axes, keepdims=True)
        var = X.var(axis=axes, keepdims=True)
        out = (X - mean) / (var + self.eps).sqrt()
        if self.weight is not None:
            out = out * self.weight
        if self.bias is not None:
            out = out + self.bias
        return out

    def __call__(self, X):
        return self.forward(X)
