"""Losses (not contractions; they stay on the array back-end as differentiable Tensor ops).
Semantics: neunet/nn/losses.py -- MSELoss 8-22, BCELoss 25-56, CrossEntropyLoss = LogSoftmax(axis=1)
+ NLLLoss 59-126 (ignore_index, class weights, mean over non-ignored targets), L1Loss 129-148."""
from __future__ import annotations

import numpy as np

from ..autograd import Tensor
from .activations import LogSoftmax
from .modules import Module


def _check(y_pred, y_true):
    if not isinstance(y_pred, Tensor) or not isinstance(y_true, Tensor):
        raise TypeError("Input values must be tensors")
    if y_pred.device != y_true.device:
        raise ValueError("Tensors must be on the same device")


class MSELoss(Module):
    def __init__(self):
        pass

    def forward(self, y_pred: Tensor, y_true: Tensor) -> Tensor:
        _check(y_pred, y_true)
        return y_pred.sub(y_true).power(2).sum().div(int(np.prod(y_pred.shape)))


class L1Loss(Module):
    def __init__(self, reduction="mean"):
        self.reduction = reduction

    def forward(self, y_pred, y_true):
        _check(y_pred, y_true)
        loss = y_pred.sub(y_true).abs()
        return loss.mean() if self.reduction == "mean" else loss.sum() if self.reduction == "sum" else loss


class BCELoss(Module):
    def __init__(self, weight=None, reduction="mean"):
        self.weight = weight
        self.reduction = reduction

    def forward(self, y_pred, y_true):
        _check(y_pred, y_true)
        loss = y_true.mul(y_pred.log()).add((1.0 - y_true).mul((1.0 - y_pred).log()))
        if self.weight is not None:
            loss = loss.mul(self.weight)
        loss = loss.mul(-1)
        return loss.mean() if self.reduction == "mean" else loss.sum() if self.reduction == "sum" else loss


class NLLLoss(Module):
    def __init__(self, weight=None, ignore_index=-100, reduction="mean"):
        self.weight = weight
        self.ignore_index = ignore_index
        self.reduction = reduction

    def forward(self, y_pred: Tensor, y_true: Tensor) -> Tensor:
        _check(y_pred, y_true)
        if y_true.dtype not in (np.int16, np.int32, np.int64):
            raise TypeError("Target must be of int dtype")
        xp = y_pred.xp
        n_cls = y_pred.shape[1]
        w = self.weight if self.weight is not None else xp.ones((n_cls,), dtype=np.float32)
        if isinstance(w, Tensor):
            w = w.data
        if tuple(w.shape) != (n_cls,):
            raise ValueError("Weight shape must be equal to number of classes")
        if y_pred.ndim == 2:
            y_pred = y_pred[..., None]
        tgt = y_true.data
        if tgt.ndim == 1:
            tgt = tgt[..., None]
        if y_pred.device == "cuda":
            tgt = tgt.to(dtype=__import__("torch").int64)
        keep = tgt != self.ignore_index
        safe = xp.where(keep, tgt, xp.zeros_like(tgt))
        # pick y_pred[b, tgt[b, ...], ...] (losses.py:107-110)
        lead = xp.arange(tgt.shape[0], dtype=np.int64).reshape((-1,) + (1,) * (tgt.ndim - 1))
        rest = [xp.arange(s, dtype=np.int64).reshape((1,) * (i + 1) + (-1,) + (1,) * (tgt.ndim - i - 2))
                for i, s in enumerate(tgt.shape[1:])]
        picked = y_pred[(lead, safe, *rest)]
        wk = w[safe] * keep
        loss = -picked * Tensor._wrap(wk, None, None, False, y_pred.device)
        if self.reduction == "mean":
            return (loss / float(xp.sum(wk))).sum() if y_pred.device == "cpu" else (loss / Tensor._wrap(xp.sum(wk), None, None, False, "cuda")).sum()
        if self.reduction == "sum":
            return loss.sum()
        return loss


class CrossEntropyLoss(Module):
    def __init__(self, weight=None, ignore_index=-100, reduction="mean"):
        self.weight = weight
        self.ignore_index = ignore_index
        self.reduction = reduction
        self.log_softmax = LogSoftmax(axis=1)
        self.nll_loss = NLLLoss(weight, ignore_index, reduction)

    def forward(self, y_pred: Tensor, y_true: Tensor) -> Tensor:
        _check(y_pred, y_true)
        if y_pred.device == "cuda" and self.weight is None and y_pred.ndim == 2 and y_true.ndim == 1:
            # fused LogSoftmax + NLL on the device (three kernels instead of ~40 element-wise launches)
            if y_true.dtype not in (np.int16, np.int32, np.int64):
                raise TypeError("Target must be of int dtype")
            from .. import b200
            lin = _pending_linear_source(y_pred) if self.reduction in ("mean", "sum") else None
            loss, saved = b200.cross_entropy_forward(y_pred.data, y_true.data, self.ignore_index, self.reduction)
            if lin is not None and lin.args[3] is None and not lin.args[4] and lin.__dict__.get("_b200_gcall") is None:
                # the logits are the output of an nn.Linear nobody else has read (lm_head): the backward goes from the
                # loss straight to the Linear's inputs -- dlogits only ever exist as bf16 operand planes (row N3)
                X, W, b, _z, _act, _beta, xst = lin.args
                out = _FusedCETensor(loss, (X, W, b, saved, xst), "cross_entropy_linear", y_pred.device)
                out.grad_fn = _fused_ce_linear_grad
                return out
            return _FusedCETensor(loss, (y_pred, saved), "cross_entropy", y_pred.device)
        return self.nll_loss(self.log_softmax(y_pred), y_true)


class _FusedCETensor(Tensor):
    def __init__(self, data, args, op, device):
        t = Tensor._wrap(data, args, op, True, device)
        self.__dict__.update(t.__dict__)
        self.grad_fn = _fused_ce_grad


def _pending_linear_source(t):
    """The still-pending nn.Linear result that `t` is a (reshape) view of, or None. Pending = nobody has read the logits
    yet, so routing the gradient around them cannot starve another consumer."""
    from ..autograd import _pending
    if not _pending(t):
        return None
    node, hops = t, 0
    while _pending(node, "view") and node.op == "reshape" and hops < 4:
        node, hops = node._f_src, hops + 1
    # only wide heads: the staged loss backward walks 256 classes per block, a 10-class head would idle 98 % of it
    if _pending(node, "linear") and node.__dict__.get("_b200_gcall") is None and len(node.shape) >= 2 \
            and int(np.prod(node.shape[:-1])) == t.shape[0] and node.shape[-1] == t.shape[1] and t.shape[1] >= 512:
        return node
    return None


def _fused_ce_linear_grad(X: Tensor, W: Tensor, b, saved, xst, grad):
    from .. import b200
    dw_out = getattr(W, "_grad_buffer", None) if W.grad is None else None
    db_out = getattr(b, "_grad_buffer", None) if (b is not None and b.grad is None) else None
    dx, dw, db = b200.cross_entropy_linear_backward(saved, grad, X.data, W.data, need_dx=X.requires_grad, need_db=b is not None,
                                                    owner=W, x_staged=xst, dw_out=dw_out, db_out=db_out)
    if dx is not None:
        X.apply_grad(dx)
    W.apply_grad(dw)
    if b is not None:
        b.apply_grad(db)


def _fused_ce_grad(y_pred: Tensor, saved, grad):
    from .. import b200
    y_pred.apply_grad(b200.cross_entropy_backward(saved, grad))
