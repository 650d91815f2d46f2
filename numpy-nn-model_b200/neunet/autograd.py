"""Tape-based autograd ``Tensor`` with the reference's public contract (neunet/autograd.py):

* ``Tensor(data, args, op, requires_grad, dtype, device)`` -- data is copied and cast to float32
  unless a dtype is given (autograd.py:16-19); ``device`` is ``"cpu"`` (NumPy) or ``"cuda"``
  (torch CUDA storage + the sm_100a kernels of ``neunet.b200``).
* every differentiable result carries ``.args`` and a ``grad_fn(*args, grad=...)`` that ends in
  ``arg.apply_grad(array)`` (autograd.py:85-93); ``backward()`` seeds ones, topologically sorts the
  tape and calls the ``grad_fn``s in reverse (autograd.py:965-1002). Layers plug into the same tape
  by returning Tensor subclasses with a hand-written ``grad_fn`` (``_LinearTensor`` etc.).

The implementation is new: ops are built from two small factories instead of ~45 hand-written
methods, the tape sort is iterative (no recursion limit on deep GPT tapes), results are wrapped
without the reference's extra copy, and ``matmul`` on ``"cuda"`` calls the tcgen05 GEMM
(forward and both backward contractions) with no array-library fallback.
"""
from __future__ import annotations

import os
from typing import Any

import numpy as np

from . import backend as _be
from .backend import get_xp


def _is_arr(x):
    return isinstance(x, np.ndarray) or _be.is_device_array(x)


# ---- deferred evaluation on "cuda" (epilogue / prologue fusion across the example's separate op calls) ------------
# The examples call q.k^T, "/ scale", where(mask), Softmax, Dropout, ".v", Linear -> Swish, "x + dropout(a)" -> RMSNorm
# as separate Tensor ops (examples/gpt.ipynb cells 2-5). To run them as the fused kernels of neunet.b200 WITHOUT
# changing the example code, a few op results on "cuda" are *deferred*: the Tensor exists at once (shape, tape
# links), its array is produced on first read of ``.data``. A consumer that recognises a still-pending producer
# launches ONE fused kernel instead (and never forces the producer); anything else that reads ``.data`` simply
# runs the producer's own kernel, so results are identical either way up to fp32 round-off.
# NEUNET_B200_FUSE=0 / set_fusion(False) restores strictly eager execution.
_FUSION = {"on": os.environ.get("NEUNET_B200_FUSE", "1") != "0"}


def set_fusion(on: bool) -> bool:
    prev = _FUSION["on"]
    _FUSION["on"] = bool(on)
    return prev


def fusion_enabled() -> bool:
    return _FUSION["on"]


class Tensor:
    __array_priority__ = 1000  # NumPy scalars/arrays defer to Tensor's reflected operators

    def __init__(self, data: Any, args=None, op=None, requires_grad: bool = True, dtype=None, device="cpu"):
        if device not in ("cpu", "cuda"):
            raise ValueError("Device must be 'cpu' or 'cuda'")
        self.xp = get_xp(device)
        if isinstance(data, Tensor):
            data = data.data
        elif isinstance(data, (int, float)) and not isinstance(data, bool):
            self._pyscalar = data  # fused kernels take constants by value (no device read-back)
        want = np.float32 if dtype is None else dtype
        if device == "cpu":
            if _be.is_device_array(data):
                data = _be.to_host(data)
            self.data = np.array(data, dtype=_be.to_numpy_dtype(want))
        else:
            self.data = self.xp.array(data, dtype=want)
        self.grad = None
        self.op = op
        self.args = args
        self.requires_grad = requires_grad
        self.device = device
        self.grad_fn = _no_grad_fn

    # ---------------------------------------------------------------------------------------
    @classmethod
    def _wrap(cls, data, args, op, requires_grad, device, cast=True):
        """Internal constructor for op results: no defensive copy; float32 cast like the
        reference's ``Tensor(result, ...)`` (autograd.py:16-19) unless ``cast`` is False."""
        t = cls.__new__(cls)
        t.xp = get_xp(device)
        if cast:
            if device == "cpu":
                data = np.asarray(data)
                if data.dtype != np.float32:
                    data = data.astype(np.float32)
            elif data.dtype != _be.torch.float32:
                data = data.to(_be.torch.float32)
        t.data = data
        t.grad = None
        t.op = op
        t.args = args
        t.requires_grad = requires_grad
        t.device = device
        t.grad_fn = _no_grad_fn
        return t

    def ensure_tensor(self, t, requires_grad=False) -> "Tensor":
        if isinstance(t, Tensor):
            if t.device != self.device:
                raise ValueError("Tensors must be on the same device")
            return t
        return Tensor(t, requires_grad=requires_grad, device=self.device, dtype=self.dtype)

    # ---- host / device movement -------------------------------------------------------------
    def numpy(self) -> np.ndarray:
        if self.device != "cpu":
            raise ValueError("Tensor must be on the CPU")
        if self.requires_grad:
            raise ValueError("Tensor must not require gradient")
        return self.data

    def to(self, device) -> "Tensor":
        if device == self.device:
            return self
        if device not in ("cpu", "cuda"):
            raise ValueError("Device must be 'cpu' or 'cuda'")
        return Tensor(self.data, requires_grad=self.requires_grad, dtype=self.dtype, device=device)

    def cpu(self):
        return self.to("cpu")

    def cuda(self):
        return self.to("cuda")

    def detach(self) -> "Tensor":
        return Tensor(self.data, args=None, op=self.op, requires_grad=False, dtype=self.dtype, device=self.device)

    def item(self):
        return self.data.item()

    def contiguous(self) -> "Tensor":
        # the reference's contiguous() is value-preserving and returns self (autograd.py:81-83)
        return self

    # ---- gradient accumulation ---------------------------------------------------------------
    def apply_grad(self, grad):
        if not self.requires_grad:
            return
        grad = self._reverse_broadcast(grad)
        self.grad = grad if self.grad is None else self.grad + grad

    def _reverse_broadcast(self, grad):
        """Sum a broadcast gradient back to this tensor's shape (autograd.py:948-962)."""
        gshape, sshape = tuple(grad.shape), tuple(self.shape)  # .shape: a pending result is not forced for its shape
        if gshape == sshape:
            return grad
        xp = self.xp
        if len(sshape) == len(gshape):
            axes = tuple(i for i, (a, b) in enumerate(zip(sshape, gshape)) if a != b)
            grad = xp.sum(grad, axis=axes, keepdims=True)
        else:
            padded = (1,) * (len(gshape) - len(sshape)) + sshape
            axes = tuple(i for i, (a, b) in enumerate(zip(padded, gshape)) if a != b)
            grad = xp.sum(grad, axis=axes)
        return grad.reshape(sshape)

    # ---- op factories ---------------------------------------------------------------------------
    def _binary(self, t, op, fwd, bwd_self, bwd_t):
        t = self.ensure_tensor(t)
        rg = self.requires_grad or t.requires_grad
        out = Tensor._wrap(fwd(self.data, t.data), (self, t) if rg else None, op, rg, self.device)
        if rg:
            def grad_fn(a: "Tensor", b: "Tensor", grad):
                if a.requires_grad:
                    a.apply_grad(bwd_self(a, b, grad))
                if b.requires_grad:
                    b.apply_grad(bwd_t(a, b, grad))
            out.grad_fn = grad_fn
        return out

    def _unary(self, op, value, bwd):
        out = Tensor._wrap(value, (self,), op, self.requires_grad, self.device)

        def grad_fn(a: "Tensor", grad):
            if a.requires_grad:
                a.apply_grad(bwd(a, grad))
        out.grad_fn = grad_fn
        return out

    # ---- arithmetic -------------------------------------------------------------------------------
    def add(self, t):
        if self.device == "cuda" and _FUSION["on"] and isinstance(t, Tensor):
            fused = _defer_add_dropout(self, t)
            if fused is not None:
                return fused
        return self._binary(t, "add", lambda a, b: a + b, lambda a, b, g: g, lambda a, b, g: g)

    def sub(self, t):
        return self._binary(t, "sub", lambda a, b: a - b, lambda a, b, g: g, lambda a, b, g: -g)

    def mul(self, t):
        return self._binary(t, "mul", lambda a, b: a * b, lambda a, b, g: g * b.data, lambda a, b, g: g * a.data)

    def div(self, t):
        if _pending(self, "matmul") and isinstance(t, (int, float)) and not isinstance(t, bool) and t != 0:
            # `scores / scale` on a not-yet-launched q.k^T: stays pending so the attention kernel can absorb it
            src, c = self, float(t)
            out = _Deferred.make(lambda: src.data / c, self.shape, (self, c), "div", self.requires_grad,
                                 _f_kind="div", _f_src=self, _f_c=c)
            out.grad_fn = _div_scalar_grad
            return out
        return self._binary(t, "div", lambda a, b: a / b, lambda a, b, g: g / b.data,
                            lambda a, b, g: -g * a.data / (b.data ** 2))

    def power(self, t):
        xp = self.xp
        return self._binary(t, "power", lambda a, b: a ** b,
                            lambda a, b, g: g * b.data * a.data ** (b.data - 1),
                            lambda a, b, g: g * a.data ** b.data * xp.log(a.data))

    def maximum(self, t):
        xp = self.xp
        return self._binary(t, "maximum", lambda a, b: xp.maximum(a, b),
                            lambda a, b, g: g * (a.data >= b.data), lambda a, b, g: g * (a.data <= b.data))

    def minimum(self, t):
        xp = self.xp
        return self._binary(t, "minimum", lambda a, b: xp.minimum(a, b),
                            lambda a, b, g: g * (a.data <= b.data), lambda a, b, g: g * (a.data >= b.data))

    def matmul(self, t):
        """``xp.matmul`` forward; backward follows the four rank cases of autograd.py:206-226.
        On "cuda" all three contractions run on the tcgen05 GEMM (neunet.b200)."""
        t = self.ensure_tensor(t)
        rg = self.requires_grad or t.requires_grad
        xp = self.xp
        staged = None
        held = {}
        if self.device == "cuda" and _FUSION["on"] and self.ndim >= 2 and t.ndim >= 2:
            from . import b200
            fused = _try_fused_attention(self, t)
            if fused is not None:
                return fused
            if self.shape[-1] != t.shape[-2]:
                raise ValueError(f"matmul: shapes {self.shape} and {t.shape} not aligned")
            a_t, b_t = self, t

            def thunk():
                if rg:
                    # keep the bf16 planes of both operands: backward (dA = G.B^T, dB = A^T.G) reads the same planes
                    data, held["staged"] = b200.matmul(a_t.data, b_t.data, keep_staged=True)
                    return data
                return b200.matmul(a_t.data, b_t.data)
            oshape = tuple(np.broadcast_shapes(self.shape[:-2], t.shape[:-2])) + (self.shape[-2], t.shape[-1])
            out = _Deferred.make(thunk, oshape, (self, t) if rg else None, "matmul", rg, _f_kind="matmul", _f_a=self, _f_b=t)
        else:
            if rg and self.device == "cuda" and self.data.ndim >= 2 and t.data.ndim >= 2:
                from . import b200
                data, staged = b200.matmul(self.data, t.data, keep_staged=True)
                held["staged"] = staged
            else:
                data = xp.matmul(self.data, t.data)
            out = Tensor._wrap(data, (self, t) if rg else None, "matmul", rg, self.device)
        if not rg:
            return out

        if self.device == "cuda":
            from . import b200

            def grad_fn(a: "Tensor", b: "Tensor", grad):
                staged = held.get("staged")
                ad, bd = a.data, b.data
                if ad.ndim == 1 and bd.ndim == 1:  # vector . vector: no contraction left
                    if a.requires_grad:
                        a.apply_grad(grad * bd)
                    if b.requires_grad:
                        b.apply_grad(grad * ad)
                    return
                # lift vectors to matrices, contract on the device, drop the lifted axis again
                a2 = ad.unsqueeze(0) if ad.ndim == 1 else ad
                b2 = bd.unsqueeze(1) if bd.ndim == 1 else bd
                g2 = grad
                if ad.ndim == 1:
                    g2 = g2.unsqueeze(-2)
                if bd.ndim == 1:
                    g2 = g2.unsqueeze(-1)
                da, db = b200.matmul_backward(a2, b2, g2, a.requires_grad, b.requires_grad, staged=staged)
                if da is not None:
                    a.apply_grad(da.squeeze(-2) if ad.ndim == 1 else da)
                if db is not None:
                    b.apply_grad(db.squeeze(-1) if bd.ndim == 1 else db)
        else:
            def grad_fn(a: "Tensor", b: "Tensor", grad):
                ad, bd = a.data, b.data
                if ad.ndim > 1 and bd.ndim > 1:
                    if a.requires_grad:
                        a.apply_grad(np.matmul(grad, np.swapaxes(bd, -1, -2)))
                    if b.requires_grad:
                        b.apply_grad(np.matmul(np.swapaxes(ad, -1, -2), grad))
                elif ad.ndim == 1 and bd.ndim == 1:
                    if a.requires_grad:
                        a.apply_grad(grad * bd)
                    if b.requires_grad:
                        b.apply_grad(grad * ad)
                elif ad.ndim == 1:
                    if a.requires_grad:
                        a.apply_grad(np.matmul(grad, np.swapaxes(bd, -1, -2)))
                    if b.requires_grad:
                        b.apply_grad(np.outer(ad, grad))
                else:
                    if a.requires_grad:
                        a.apply_grad(np.outer(grad, bd))
                    if b.requires_grad:
                        b.apply_grad(np.matmul(np.swapaxes(ad, -1, -2), grad))
        out.grad_fn = grad_fn
        return out

    # ---- reductions ------------------------------------------------------------------------------
    @staticmethod
    def _axis_of(args, kwargs):
        return kwargs.get("axis", None) if len(args) == 0 else args[0]

    def _reduce(self, op, args, kwargs, scale_fn):
        axis = self._axis_of(args, kwargs)
        keepdims = kwargs.get("keepdims", args[1] if len(args) > 1 else False)
        xp = self.xp
        value = getattr(xp, op)(self.data, axis=axis, keepdims=keepdims)
        out = Tensor._wrap(value, (self, axis), op, self.requires_grad, self.device)

        def grad_fn(a: "Tensor", axis, grad):
            if not a.requires_grad:
                return
            if grad.ndim != a.data.ndim and axis is not None:
                grad = xp.expand_dims(grad, axis)
            a.apply_grad(scale_fn(a, axis, xp.ones_like(a.data) * grad))
        out.grad_fn = grad_fn
        return out

    @staticmethod
    def _count(a, axis):
        shape = a.data.shape
        if axis is None:
            return int(np.prod(shape))
        axes = axis if isinstance(axis, (tuple, list)) else (axis,)
        return int(np.prod([shape[i] for i in axes]))

    def sum(self, *args, **kwargs):
        return self._reduce("sum", args, kwargs, lambda a, axis, g: g)

    def mean(self, *args, **kwargs):
        return self._reduce("mean", args, kwargs, lambda a, axis, g: g / Tensor._count(a, axis))

    def var(self, *args, **kwargs):  # ddof = 0
        xp = self.xp
        return self._reduce("var", args, kwargs,
                            lambda a, axis, g: g * 2 * (a.data - xp.mean(a.data, axis=axis, keepdims=True))
                            / Tensor._count(a, axis))

    def _extremum(self, op, axis, keepdims):
        xp = self.xp
        fn = getattr(xp, op)
        out = Tensor._wrap(fn(self.data, axis=axis, keepdims=keepdims), (self, axis), op, self.requires_grad, self.device)

        def grad_fn(a: "Tensor", axis, grad):
            if not a.requires_grad:
                return
            if grad.ndim != a.data.ndim and axis is not None:
                grad = xp.expand_dims(grad, axis)
            a.apply_grad(grad * (a.data == fn(a.data, axis=axis, keepdims=True)))
        out.grad_fn = grad_fn
        return out

    def max(self, axis=None, keepdims=False):
        return self._extremum("max", axis, keepdims)

    def min(self, axis=None, keepdims=False):
        return self._extremum("min", axis, keepdims)

    # ---- element-wise math -------------------------------------------------------------------------
    def sqrt(self):
        return self._unary("sqrt", self.xp.sqrt(self.data), lambda a, g: g * 0.5 * a.data ** -0.5)

    def log(self):
        return self._unary("log", self.xp.log(self.data), lambda a, g: g * 1 / a.data)

    def exp(self):
        xp = self.xp
        return self._unary("exp", xp.exp(self.data), lambda a, g: g * xp.exp(a.data))

    def tanh(self):
        xp = self.xp
        return self._unary("tanh", xp.tanh(self.data), lambda a, g: g * (1 - xp.tanh(a.data) ** 2))

    def sin(self):
        xp = self.xp
        return self._unary("sin", xp.sin(self.data), lambda a, g: g * xp.cos(a.data))

    def cos(self):
        xp = self.xp
        return self._unary("cos", xp.cos(self.data), lambda a, g: g * -xp.sin(a.data))

    def abs(self):
        xp = self.xp
        return self._unary("abs", xp.abs(self.data), lambda a, g: g * xp.sign(a.data))

    def __neg__(self):
        return self._unary("neg", -self.data, lambda a, g: -g)

    def __pos__(self):
        return self._unary("pos", self.data, lambda a, g: g)

    # ---- shape ops -------------------------------------------------------------------------------------
    def concatenate(self, *tensors, axis=0):
        tensors = tuple(self.ensure_tensor(t) for t in tensors)
        parts = (self,) + tensors
        xp = self.xp
        rg = any(p.requires_grad for p in parts)
        out = Tensor._wrap(xp.concatenate([p.data for p in parts], axis=axis), parts + (axis,), "concatenate", rg,
                           self.device)

        def grad_fn(*args, grad):
            *ins, ax = args
            start = 0
            for p in ins:
                n = p.data.shape[ax]
                if p.requires_grad:
                    idx = [slice(None)] * grad.ndim
                    idx[ax] = slice(start, start + n)
                    p.apply_grad(grad[tuple(idx)])
                start += n
        out.grad_fn = grad_fn
        return out

    def reshape(self, *shape):
        shape = shape[0] if len(shape) == 1 else shape
        if isinstance(shape, int):
            shape = (shape,)
        if _pending(self):
            # a view of a pending result stays pending (q/k/v head splits must not force their Linear one by one)
            src, shp = self, _resolve_shape(tuple(shape), self.size)
            out = _Deferred.make(lambda: src.data.reshape(shp), shp, (self,), "reshape",
                                 self.requires_grad, _f_kind="view", _f_src=self)
        else:
            out = Tensor._wrap(self.data.reshape(tuple(shape)), (self,), "reshape", self.requires_grad, self.device, cast=False)
            if self.device == "cuda" and getattr(self, "_b200_xst", None) is not None:
                out._b200_xst = self._b200_xst

        def grad_fn(a: "Tensor", grad):
            if a.requires_grad:
                a.apply_grad(grad.reshape(tuple(a.data.shape)))
        out.grad_fn = grad_fn
        return out

    def transpose(self, *axes):
        axes = axes[0] if len(axes) == 1 else axes
        if axes is None or (hasattr(axes, "__len__") and len(axes) == 0):
            axes = tuple(range(self.data.ndim))[::-1]
        axes = tuple(axes)
        xp = self.xp
        if _pending(self):
            src = self
            out = _Deferred.make(lambda: xp.transpose(src.data, axes), tuple(self.shape[a] for a in axes),
                                 (self, axes), "transpose", self.requires_grad, _f_kind="view", _f_src=self)
        else:
            out = Tensor._wrap(xp.transpose(self.data, axes), (self, axes), "transpose", self.requires_grad, self.device)
            if self.device == "cuda" and getattr(self, "_b200_xst", None) is not None:
                out._b200_xst = self._b200_xst

        def grad_fn(a: "Tensor", axes, grad):
            # NB: like the reference (autograd.py:618-620) the gradient is permuted by `axes` again,
            # which is the inverse permutation only for involutions such as (0,2,1,3)/(0,1,3,2).
            if a.requires_grad:
                a.apply_grad(xp.transpose(grad, axes))
        out.grad_fn = grad_fn
        return out

    def swapaxes(self, axis1, axis2):
        xp = self.xp
        out = Tensor._wrap(xp.swapaxes(self.data, axis1, axis2), (self, axis1, axis2), "swapaxes", self.requires_grad,
                           self.device)

        def grad_fn(a: "Tensor", axis1, axis2, grad):
            if a.requires_grad:
                a.apply_grad(xp.swapaxes(grad, axis1, axis2))
        out.grad_fn = grad_fn
        return out

    def flip(self, axis):
        if axis is None:
            axis = tuple(range(self.data.ndim))
        xp = self.xp
        out = Tensor._wrap(xp.flip(self.data, axis), (self, axis), "flip", self.requires_grad, self.device)

        def grad_fn(a: "Tensor", axis, grad):
            if a.requires_grad:
                a.apply_grad(xp.flip(grad, axis))
        out.grad_fn = grad_fn
        return out

    def where(self, condition, t):
        """``where(condition, self, t)`` (autograd.py:658-684)."""
        condition = self.ensure_tensor(condition)
        t = self.ensure_tensor(t)
        rg = self.requires_grad or t.requires_grad
        xp = self.xp
        if (self.device == "cuda" and (_pending(t, "div") or _pending(t, "matmul")) and not self.requires_grad
                and not condition.requires_grad and getattr(self, "_pyscalar", None) is not None and self.ndim == 0):
            # where(mask, constant, scores) on pending attention scores: stays pending (absorbed by the attention kernel)
            a_t, c_t, b_t = self, condition, t
            out = _Deferred.make(lambda: xp.where(c_t.data != 0, a_t.data, b_t.data),
                                 tuple(np.broadcast_shapes(condition.shape, t.shape)), (self, condition, t) if rg else None,
                                 "where", rg, _f_kind="where", _f_fill=float(self._pyscalar), _f_cond=condition, _f_src=t)
        else:
            cond = condition.data != 0 if self.device == "cuda" else condition.data
            out = Tensor._wrap(xp.where(cond, self.data, t.data), (self, condition, t) if rg else None, "where", rg,
                               self.device)
        if rg:
            def grad_fn(a: "Tensor", condition: "Tensor", b: "Tensor", grad):
                c = condition.data != 0 if a.device == "cuda" else condition.data
                if a.requires_grad:
                    a.apply_grad(xp.where(c, grad, xp.zeros_like(grad)))
                if b.requires_grad:
                    b.apply_grad(xp.where(c, xp.zeros_like(grad), grad))
            out.grad_fn = grad_fn
        return out

    # ---- comparisons (non-differentiable, float32 0/1 results like the reference) -----------------------
    def _compare(self, t, op, fn):
        if (op == "equal" and self.device == "cuda" and _FUSION["on"] and isinstance(t, (int, float))
                and not isinstance(t, bool) and not _pending(self)
                and self.data.dtype in (_be.torch.int32, _be.torch.float32)):
            # `mask == 0`: pending, so the attention kernel can test the mask itself instead of reading a 0/1 tensor
            src, c = self, t
            return _Deferred.make(lambda: (src.data == c).to(_be.torch.float32), self.shape, None, op, False,
                                  _f_kind="eq_scalar", _f_base=self, _f_value=float(t))
        t = self.ensure_tensor(t)
        return Tensor._wrap(fn(self.data, t.data), None, op, False, self.device)

    def equal(self, t): return self._compare(t, "equal", lambda a, b: a == b)
    def not_equal(self, t): return self._compare(t, "not_equal", lambda a, b: a != b)
    def greater(self, t): return self._compare(t, "greater", lambda a, b: a > b)
    def greater_equal(self, t): return self._compare(t, "greater_equal", lambda a, b: a >= b)
    def less(self, t): return self._compare(t, "less", lambda a, b: a < b)
    def less_equal(self, t): return self._compare(t, "less_equal", lambda a, b: a <= b)
    def logical_and(self, t): return self._compare(t, "logical_and", lambda a, b: self.xp.logical_and(a, b))
    def logical_or(self, t): return self._compare(t, "logical_or", lambda a, b: self.xp.logical_or(a, b))
    def logical_not(self): return Tensor._wrap(self.xp.logical_not(self.data), None, "logical_not", False, self.device)

    __eq__ = equal  # type: ignore[assignment]
    __ne__ = not_equal  # type: ignore[assignment]
    __gt__ = greater
    __ge__ = greater_equal
    __lt__ = less
    __le__ = less_equal
    __and__ = logical_and
    __or__ = logical_or
    __hash__ = object.__hash__

    def __invert__(self): return self.logical_not()
    def __abs__(self): return self.abs()
    def __add__(self, t): return self.add(t)
    def __sub__(self, t): return self.sub(t)
    def __mul__(self, t): return self.mul(t)
    def __truediv__(self, t): return self.div(t)
    def __matmul__(self, t): return self.matmul(t)
    def __pow__(self, t): return self.power(t)
    def __radd__(self, t): return self.ensure_tensor(t).add(self)
    def __rsub__(self, t): return self.ensure_tensor(t).sub(self)
    def __rmul__(self, t): return self.ensure_tensor(t).mul(self)
    def __rtruediv__(self, t): return self.ensure_tensor(t).div(self)
    def __rmatmul__(self, t): return self.ensure_tensor(t).matmul(self)
    def __rpow__(self, t): return self.ensure_tensor(t).power(self)

    def __repr__(self):
        return f"Tensor({self.data}, requires_grad={self.requires_grad}, dtype={self.dtype}, device={self.device})"

    # ---- indexing ------------------------------------------------------------------------------------------
    def _index(self, index):
        """Indices may contain Tensors / NumPy arrays; on "cuda" array indices are moved to the device."""
        def conv(i):
            if isinstance(i, Tensor):
                i = i.data
            if self.device == "cuda" and isinstance(i, np.ndarray):
                i = self.xp.array(i, dtype=np.int64 if i.dtype != np.bool_ else np.bool_)
            elif self.device == "cuda" and _be.is_device_array(i) and i.dtype not in (_be.torch.int64, _be.torch.bool):
                i = i.to(_be.torch.int64)
            elif self.device == "cpu" and isinstance(i, np.ndarray) and i.dtype.kind == "f":
                i = i.astype(np.int64)
            return i
        if isinstance(index, tuple):
            return tuple(conv(i) for i in index)
        return conv(index)

    def __getitem__(self, index):
        index = self._index(index)
        xp = self.xp
        out = Tensor._wrap(self.data[index], (self, index), "getitem", self.requires_grad, self.device, cast=False)

        def grad_fn(a: "Tensor", index, grad):
            if a.requires_grad:
                # assignment, not accumulation: duplicate indices keep the last write (autograd.py:909-910)
                full = xp.zeros_like(a.data)
                rows = None
                if a.device == "cuda":
                    # an integer index tensor over the first axis, bare or followed only by new axes
                    # (ddpm cell 4: coef[timesteps, None, None, None]): a device index_put with duplicate
                    # indices is non-deterministic, the reference keeps the LAST write
                    first = index[0] if (isinstance(index, tuple) and index and all(i is None for i in index[1:])) else index
                    if _be.is_device_array(first) and first.dtype == _be.torch.int64:
                        rows = first
                if rows is not None:
                    _assign_last_wins(full, rows, grad)
                else:
                    full[index] = grad
                a.apply_grad(full)
        out.grad_fn = grad_fn
        return out

    def __setitem__(self, key, value):
        if self.requires_grad:
            raise RuntimeError("Cannot assign values to a tensor with requires_grad=True")
        value = self.ensure_tensor(value)
        self.data[self._index(key)] = value.data

    def __array__(self, dtype=None, copy=None):
        host = _be.to_host(self.data)
        return host.astype(dtype, copy=False) if dtype is not None else host

    def __len__(self):
        return self.data.shape[0]

    @property
    def shape(self) -> tuple:
        return tuple(self.data.shape)

    @property
    def T(self):
        return self.transpose()

    @property
    def dtype(self):
        d = self.data.dtype
        return d if isinstance(d, np.dtype) else _be.to_numpy_dtype(d)

    @property
    def ndim(self) -> int:
        return self.data.ndim

    @property
    def size(self) -> int:
        return int(np.prod(self.data.shape))

    # ---- backward --------------------------------------------------------------------------------------------
    def backward(self, grad=None):
        if not self.requires_grad:
            return
        xp = self.xp
        self.data  # noqa: B018 -- a pending root (deferred evaluation) is produced now: its producer fills in what backward needs
        if grad is None:
            grad = xp.ones_like(self.data)
        elif isinstance(grad, Tensor):
            grad = grad.data
        if not _is_arr(grad) or (self.device == "cuda" and isinstance(grad, np.ndarray)):
            grad = xp.array(grad, dtype=self.dtype)
        elif self.device == "cpu":
            grad = np.array(grad, dtype=self.dtype)
        self.apply_grad(grad)

        # iterative post-order DFS over `.args` (the reference recurses, autograd.py:982-999)
        tape, seen = [], {id(self)}
        stack = [(self, 0)]
        while stack:
            node, i = stack.pop()
            children = node.args if node.args is not None else ()
            if node.args is None:
                continue  # leaves never enter the tape
            advanced = False
            while i < len(children):
                c = children[i]
                i += 1
                if isinstance(c, Tensor) and c.requires_grad and id(c) not in seen:
                    seen.add(id(c))
                    stack.append((node, i))
                    stack.append((c, 0))
                    advanced = True
                    break
            if not advanced:
                tape.append(node)
        # leaves with a `_grad_ready` hook (neunet.distributed.GradBucket.overlap_backward): the hook
        # fires right after the LAST tape node that feeds the leaf has run, i.e. when its gradient is final
        # fused sibling Linear layers (nn/layers/linear.py: _GroupCall) run ONE backward when the last member that takes
        # part in this tape has received its gradient: count the members first
        gcalls = {}
        for v in tape:
            gc = v.__dict__.get("_b200_gcall")
            if gc is not None:
                gcalls[id(gc)] = gc
                gc.expected, gc.arrived = 0, 0
        for v in tape:
            gc = v.__dict__.get("_b200_gcall")
            if gc is not None:
                gc.expected += 1
        uses = {}
        for v in tape:
            for a in v.args:
                if isinstance(a, Tensor) and a.args is None and a.requires_grad and getattr(a, "_grad_ready", None) is not None:
                    ent = uses.get(id(a))
                    if ent is None:
                        uses[id(a)] = [a, 1]
                    else:
                        ent[1] += 1
        for v in reversed(tape):
            g = v.__dict__.get("_grad") if isinstance(v, _Deferred) else v.grad  # raw: a parked _MaskedGrad goes to its consumer as is
            v.grad_fn(*v.args, grad=g)
            if uses:
                for a in v.args:
                    ent = uses.get(id(a)) if isinstance(a, Tensor) else None
                    if ent is not None:
                        ent[1] -= 1
                        if ent[1] == 0:
                            del uses[id(a)]
                            a._grad_ready(a)


class _MaskedGrad:
    """An upstream gradient that still has to pass the backward of an nn.Dropout (``raw * mask(ticket) / (1 - p)``), parked
    as the ``.grad`` of an nn.Linear result: the Linear's backward applies the mask inside the pass that converts its
    upstream gradient to bf16 planes (neunet.b200.linear_backward: grad_drop), so the stand-alone mask kernel and its
    read + write of the whole gradient disappear. Anything else that reads ``.grad`` gets the masked array
    (``_Deferred.grad`` resolves it)."""
    __slots__ = ("raw", "p", "ticket")

    def __init__(self, raw, p, ticket):
        self.raw, self.p, self.ticket = raw, p, ticket

    def resolve(self):
        from . import b200
        return b200.dropout_apply(self.raw, self.p, self.ticket)


_MASKED_GRAD_CONSUMERS = set()  # grad_fns that accept a _MaskedGrad (nn/layers/linear.py registers _linear_grad_fn)


def _apply_dropped_grad(a, grad, p, ticket):
    """``a.grad += grad * dropout_mask`` -- parked unmasked when ``a`` is an nn.Linear result with no gradient yet."""
    if (_FUSION["on"] and isinstance(a, _Deferred) and a.requires_grad and a.__dict__.get("_grad") is None
            and a.grad_fn in _MASKED_GRAD_CONSUMERS and tuple(grad.shape) == tuple(a.shape) and grad.shape[-1] % 4 == 0):
        a.__dict__["_grad"] = _MaskedGrad(grad, p, ticket)
        return
    from . import b200
    a.apply_grad(b200.dropout_apply(grad, p, ticket))


class _Deferred(Tensor):
    """A "cuda" result whose array is produced on first read of ``.data`` (see the note on deferred
    evaluation at the top of this file). ``_f_kind`` and the other ``_f_*`` attributes describe the
    producer to consumers that can absorb it into a fused kernel."""

    @classmethod
    def make(cls, thunk, shape, args, op, requires_grad, **meta):
        t = cls.__new__(cls)
        d = t.__dict__
        d["_data"] = None
        d["_thunk"] = thunk
        d["_shape"] = tuple(int(v) for v in shape)
        d["xp"] = get_xp("cuda")
        d["_grad"] = None
        d["op"] = op
        d["args"] = args
        d["requires_grad"] = requires_grad
        d["device"] = "cuda"
        d["grad_fn"] = _no_grad_fn
        d.update(meta)
        return t

    @property
    def data(self):
        d = self._data
        if d is None:
            thunk = self._thunk
            d = thunk()
            if self._data is None:  # a fused consumer may have delivered the array meanwhile
                self._data = d
            else:
                d = self._data
            self._thunk = None
            if self.__dict__.get("_f_kind") == "view":
                # views inherit the ready-made bf16 operand planes of their source (validated by
                # pointer / version / shape at the consuming nn.Linear, so a mismatching entry is ignored)
                planes = self._f_src.__dict__.get("_b200_xst")
                if planes is not None:
                    self.__dict__["_b200_xst"] = planes
        return d

    @data.setter
    def data(self, value):
        self._data = value
        self._thunk = None

    @property
    def pending(self):
        return self._data is None

    @property
    def grad(self):
        g = self.__dict__.get("_grad")
        if type(g) is _MaskedGrad:
            g = self.__dict__["_grad"] = g.resolve()
        return g

    @grad.setter
    def grad(self, value):
        self.__dict__["_grad"] = value

    @property
    def shape(self):
        return self._shape if self._data is None else tuple(self._data.shape)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def dtype(self):
        if self._data is None:
            return np.dtype(np.float32)
        return _be.to_numpy_dtype(self._data.dtype)

    def __len__(self):
        return self.shape[0]


def _pending(t, kind=None):
    return (isinstance(t, _Deferred) and t._data is None and _FUSION["on"]
            and (kind is None or t.__dict__.get("_f_kind") == kind))


def _resolve_shape(shape, size):
    shape = tuple(int(v) for v in shape)
    if -1 in shape:
        known = int(np.prod([v for v in shape if v != -1])) or 1
        shape = tuple(size // known if v == -1 else v for v in shape)
    return shape


def _div_scalar_grad(a, c, grad):
    if a.requires_grad:
        a.apply_grad(grad / c)


def _defer_add_dropout(x, t):
    """`x + dropout(a)` (either order) with the dropout still pending: one kernel, and the sum itself stays
    pending so that a following RMSNorm can absorb it as its prologue."""
    if _pending(t, "dropout") and not _pending(x, "dropout"):
        res, drop = x, t
    elif _pending(x, "dropout") and not _pending(t, "dropout"):
        res, drop = t, x
    else:
        return None
    if tuple(res.shape) != tuple(drop.shape) or res.device != "cuda" or drop._f_ticket is None:
        return None
    from . import b200
    a_t, p, ticket = drop._f_src, drop._f_p, drop._f_ticket
    rg = res.requires_grad or a_t.requires_grad
    def thunk():
        if len(res.shape) not in (2, 3):
            return b200.dropout_apply(a_t.data, p, ticket, residual=res.data)
        y, planes = b200.dropout_apply(a_t.data, p, ticket, residual=res.data, want_planes=True)
        out._b200_xst = planes  # the last residual sum of the stack feeds fc_out directly
        return y
    out = _Deferred.make(thunk, res.shape, (res, a_t, drop.args[1]) if rg else None, "add_dropout", rg,
                         _f_kind="add_dropout", _f_x=res, _f_a=a_t, _f_p=p, _f_ticket=ticket)

    def grad_fn(xr, a, mask, grad):
        if xr.requires_grad:
            xr.apply_grad(grad)
        if a.requires_grad:
            _apply_dropped_grad(a, grad, mask.p, mask.ticket)
    out.grad_fn = grad_fn
    return out


def _try_fused_attention(attn, v):
    """`matmul(dropout(softmax(where(mask, fill, (q @ kT) / scale))), v)` with every intermediate still pending
    -> ONE kernel (neunet.b200.attention_forward); None when the pattern or the size limits do not match."""
    if not _pending(attn):
        return None
    p, ticket, sm = 0.0, None, attn
    if attn._f_kind == "dropout":
        p, ticket, sm = attn._f_p, attn._f_ticket, attn._f_src
    if not _pending(sm, "softmax"):
        return None
    src = sm._f_src
    mask, fill = None, 0.0
    if _pending(src, "where"):
        cond, fill = src._f_cond, src._f_fill
        src = src._f_src
        if _pending(cond, "eq_scalar"):
            base = cond._f_base.data
            mask = (base, 2 if base.dtype == _be.torch.int32 else 3, cond._f_value)
        else:
            mask = (cond.data, 1, 0.0)
    scale = 1.0
    if _pending(src, "div"):
        scale, src = src._f_c, src._f_src
    if not _pending(src, "matmul"):
        return None
    q_t, kt_t = src._f_a, src._f_b
    if not (q_t.ndim == 4 and kt_t.ndim == 4 and v.ndim == 4):
        return None
    B, H, Tq, D = q_t.shape
    Tk = kt_t.shape[3]
    if kt_t.shape != (B, H, D, Tk) or v.shape != (B, H, Tk, D) or attn.shape != (B, H, Tq, Tk):
        return None
    from . import b200
    if not b200.attention_supported(Tq, Tk, D):
        return None
    if mask is not None:
        ms = tuple(mask[0].shape)
        if len(ms) > 4 or any(a != 1 and a != b for a, b in zip(ms[::-1], (B, H, Tq, Tk)[::-1])):
            return None
    out, attn_data, planes = b200.attention_forward(q_t.data, kt_t.data, v.data, mask, fill, scale, p, ticket,
                                                    want_planes=True)
    attn.data = attn_data  # what the example returns as `attn`: delivered by the fused kernel
    rg = q_t.requires_grad or kt_t.requires_grad or v.requires_grad
    node = Tensor._wrap(out, (q_t, kt_t, v) if rg else None, "attention", rg, "cuda", cast=False)
    node._b200_xst = planes
    if rg:
        def grad_fn(q, kt, vv, grad):
            dq, dkt, dv = b200.attention_backward(q.data, kt.data, vv.data, mask, fill, scale, p, ticket, grad)
            if q.requires_grad:
                q.apply_grad(dq)
            if kt.requires_grad:
                kt.apply_grad(dkt)
            if vv.requires_grad:
                vv.apply_grad(dv)
        node.grad_fn = grad_fn
    return node


def _assign_last_wins(full, index, grad):
    """``full[index] = grad`` for one integer index tensor over the first axis, with NumPy's
    deterministic semantics for duplicate indices (the LAST occurrence wins). A plain device
    index_put with duplicates is non-deterministic; the reference's CPU path keeps the last write (autograd.py:909-910)."""
    torch = _be.torch
    flat = index.reshape(-1)
    n = flat.numel()
    rows = full.shape[0]
    flat = torch.where(flat < 0, flat + rows, flat)
    order = torch.arange(n, device=flat.device)
    last = torch.full((rows,), -1, dtype=torch.int64, device=flat.device)
    last.scatter_reduce_(0, flat, order, reduce="amax", include_self=True)
    hit = (last >= 0).reshape((rows,) + (1,) * (full.ndim - 1))
    g2 = grad.reshape((n,) + tuple(full.shape[1:]))
    # gather form (static shapes, no host sync: legal inside CUDA-graph capture)
    full.copy_(torch.where(hit, g2[last.clamp(min=0)], torch.zeros((), dtype=full.dtype, device=full.device)))


def _no_grad_fn(*args, **kwargs):
    return None
