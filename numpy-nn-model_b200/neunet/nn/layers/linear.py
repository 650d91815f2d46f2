"""nn.Linear: Y = X . W^T + b  (reference: neunet/nn/layers/linear.py:13-61).

On ``device="cuda"`` forward, dgrad, wgrad and the bias gradient run on the tcgen05 GEMM through
``neunet.b200`` (the seam the reference's ``CUDALinear`` uses, experimental/linear/linear.py:122-215);
``LinearSwish`` is the fused Linear+Swish layer (reference: ``CUDALinearSwish``,
experimental/linear_swish/linear_swish_cutlass.py:198-278): Swish is the GEMM epilogue and
swish' is folded into the staging of dO in backward.
"""
from __future__ import annotations

import numpy as np

from ... import tensor as _tensor
from ...autograd import Tensor
from ..modules import Module
from ..parameter import Parameter


class _LinearTensor(Tensor):
    """Result tensor with the static backward of linear.py:17-24."""

    def __init__(self, data, args, op, device):
        t = Tensor._wrap(data, args, op, True, device)
        self.__dict__.update(t.__dict__)
        self.grad_fn = _linear_grad_fn


def _linear_grad_fn(X: Tensor, weight: Tensor, bias, Z, act, beta, x_staged, grad):
    if X.device == "cuda":
        from ... import b200
        # data-parallel training: the first (normally only) gradient contribution of a parameter is written
        # straight into its slice of the all-reduce bucket (neunet.distributed.GradBucket.overlap_backward)
        dw_out = getattr(weight, "_grad_buffer", None) if weight.grad is None else None
        db_out = getattr(bias, "_grad_buffer", None) if (bias is not None and bias.grad is None) else None
        dx, dw, db = b200.linear_backward(X.data, weight.data, grad, z=Z, act=act, beta=beta,
                                          need_dx=X.requires_grad, need_db=bias is not None, owner=weight,
                                          x_staged=x_staged, dw_out=dw_out, db_out=db_out)
        if dx is not None:
            X.apply_grad(dx)
        weight.apply_grad(dw)
        if bias is not None:
            bias.apply_grad(db)
        return
    if act:  # Swish backward on the saved pre-activation
        s = 1 / (1 + np.exp(-beta * Z))
        f = Z * s
        grad = grad * (beta * f + s * (1 - beta * f))
    X.apply_grad(np.matmul(grad, weight.data))
    weight.apply_grad(np.swapaxes(np.matmul(np.swapaxes(X.data, -1, -2), grad), -1, -2))
    if bias is not None:
        bias.apply_grad(np.sum(grad, axis=0, keepdims=True))


def _deferred_linear(layer, X: Tensor) -> Tensor:
    """``Linear.forward`` on "cuda" with fusion on: the result is *pending* (neunet/autograd.py, deferred
    evaluation). The GEMM is launched when the output is first read -- together with every other pending
    Linear on the same input (the q/k/v projections of one RMSNorm output run back to back on the same staged
    operand) -- unless an ``nn.Swish`` consumes it first, in which case Swish becomes the GEMM's epilogue."""
    from ... import b200
    from ...autograd import _Deferred
    W, b = layer.weight, layer.bias
    act0, beta0, training = layer._act, layer._beta, layer.training_mode()

    def run(act, beta):
        return b200.linear_forward(X.data, W.data, b.data if b is not None else None, act=act, beta=beta,
                                   save_z=bool(act), owner=W, keep_x_staged=training, x_owner=X)

    def thunk():
        sibs = X.__dict__.get("_b200_pending_lin")
        if sibs:  # launch the siblings right behind this one: same X planes, still hot in L2
            X.__dict__["_b200_pending_lin"] = None
        O, Z, xst = run(act0, beta0)
        out.args = (X, W, b, Z, act0, beta0, xst)
        out.data = O
        if sibs:
            for sib in sibs:
                if sib is not out and sib._data is None:
                    sib.data  # noqa: B018 -- forces the sibling's GEMM
        return O

    def fuse_swish(beta):
        O, Z, xst = run(1, beta)
        out.args = (X, W, b, None, 0, 1.0, xst)
        out.data = Z  # the pre-activation is the GEMM's side output: the Linear result itself is delivered for free
        return _LinearTensor(O, (X, W, b, Z, 1, beta, xst), "linear_swish", "cuda")

    out = _Deferred.make(thunk, tuple(X.shape[:-1]) + (layer.out_features,), (X, W, b, None, act0, beta0, None), "linear", True,
                         _f_kind="linear", _f_act=act0, _f_fuse_swish=fuse_swish)
    out.grad_fn = _linear_grad_fn
    try:
        lst = X.__dict__.get("_b200_pending_lin")
        if lst is None:
            lst = X.__dict__["_b200_pending_lin"] = []
        lst.append(out)
    except AttributeError:
        pass
    return out


class Linear(Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, device="cpu"):
        self.in_features = in_features
        self.out_features = out_features
        stdv = 1.0 / np.sqrt(in_features)
        # same draws, same order, from the global np.random (linear.py:34-43)
        self.weight = Parameter(_tensor(np.random.uniform(-stdv, stdv, (out_features, in_features)), dtype=np.float32))
        self.bias = (Parameter(_tensor(np.random.uniform(-stdv, stdv, (1, out_features)), dtype=np.float32))
                     if bias else None)
        self.to(device)

    _act, _beta = 0, 1.0

    def forward(self, X: Tensor) -> Tensor:
        if not isinstance(X, Tensor):
            raise TypeError("Input must be a tensor")
        if X.device != self.device:
            raise ValueError("Tensors must be on the same device")
        b = self.bias
        if self.device == "cuda":
            from ... import b200
            from ...autograd import fusion_enabled
            if fusion_enabled():
                return _deferred_linear(self, X)
            # the bf16 planes of X made for the forward GEMM are kept for wgrad (dW = dZ^T . X)
            O, Z, xst = b200.linear_forward(X.data, self.weight.data, b.data if b is not None else None,
                                            act=self._act, beta=self._beta, save_z=bool(self._act), owner=self.weight,
                                            keep_x_staged=self.training_mode(), x_owner=X)
        else:
            xst = None
            Z = np.matmul(X.data, self.weight.data.T)
            if b is not None:
                Z = Z + b.data
            O = Z / (1 + np.exp(-self._beta * Z)) if self._act else Z
            if not self._act:
                Z = None
        return _LinearTensor(O, (X, self.weight, b, Z, self._act, self._beta, xst), "linear", self.device)

    def training_mode(self):
        return getattr(self, "training", True)

    def __call__(self, X):
        return self.forward(X)


class LinearSwish(Linear):
    """Linear followed by Swish(beta) in one pass: ``swish(X . W^T + b)``."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, beta: float = 1.0, device="cpu"):
        self._act, self._beta = 1, float(beta)
        super().__init__(in_features, out_features, bias=bias, device=device)
