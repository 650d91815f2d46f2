"""neunet -- B200-native drop-in for the dense hot path of AkiRusProd/numpy-nn-model.

Same public surface as the reference package (``neunet/__init__.py``): torch-like factories and
functional wrappers around ``Tensor``, ``neunet.nn`` layers and ``neunet.optim`` optimizers, with
``device="cpu"`` (NumPy) and ``device="cuda"`` (torch CUDA storage; ``matmul`` / ``nn.Linear`` /
``nn.Conv2d`` / Swish / Softmax / RMSNorm / Adam(W) run on hand-written sm_100a kernels through the
ctypes C-ABI in ``neunet.b200``; there is no fallback on that path).
"""
from __future__ import annotations

import pickle
from pathlib import Path

import numpy as np

from . import backend as _be
from .autograd import Tensor
from .backend import get_xp as _get_xp

int16 = np.int16
int32 = np.int32
int64 = np.int64
float16 = np.float16
float32 = np.float32
float64 = np.float64


def _shape(shape):
    return tuple(shape[0]) if len(shape) == 1 and isinstance(shape[0], (list, tuple)) else tuple(shape)


def save(obj, f, pickle_protocol: int = 2):
    with Path(f).open("wb") as fh:
        pickle.dump(obj, fh, protocol=pickle_protocol)


def load(f):
    with Path(f).open("rb") as fh:
        return pickle.load(fh)


def tensor(data, requires_grad=False, dtype=float32, device=None):
    return Tensor(data, requires_grad=requires_grad, dtype=dtype, device=device or "cpu")


def _filled(kind, shape, dtype, requires_grad, device):
    device = device or "cpu"
    dt = float32 if dtype is None else dtype
    xp = _get_xp(device)
    return Tensor._wrap(getattr(xp, kind)(_shape(shape), dtype=dt), None, None, requires_grad, device, cast=False)


def ones(*shape, dtype=None, requires_grad=False, device=None):
    return _filled("ones", shape, dtype, requires_grad, device)


def zeros(*shape, dtype=None, requires_grad=False, device=None):
    return _filled("zeros", shape, dtype, requires_grad, device)


def rand(*shape, dtype=None, requires_grad=False, device=None):
    dt = float32 if dtype is None else dtype
    return Tensor(np.random.rand(*_shape(shape)), requires_grad=requires_grad, dtype=dt, device=device or "cpu")


def randn(*shape, dtype=None, requires_grad=False, device=None):
    dt = float32 if dtype is None else dtype
    return Tensor(np.random.randn(*_shape(shape)), requires_grad=requires_grad, dtype=dt, device=device or "cpu")


def arange(start=0, end=None, step=1, dtype=None, requires_grad=False, device=None):
    if end is None:
        start, end = 0, start
    dt = float32 if dtype is None else dtype
    return Tensor(np.arange(start, end, step), requires_grad=requires_grad, dtype=dt, device=device or "cpu")


def ones_like(tensor, dtype=None, requires_grad=False, device=None):
    return ones(*tensor.shape, dtype=tensor.dtype if dtype is None else dtype, requires_grad=requires_grad,
                device=tensor.device if device is None else device)


def zeros_like(tensor, dtype=None, requires_grad=False, device=None):
    return zeros(*tensor.shape, dtype=tensor.dtype if dtype is None else dtype, requires_grad=requires_grad,
                 device=tensor.device if device is None else device)


def argmax(x, axis=None, keepdims=False):
    return Tensor(x.xp.argmax(x.data, axis=axis, keepdims=keepdims), requires_grad=False, device=x.device, dtype=int32)


def argmin(x, axis=None, keepdims=False):
    return Tensor(x.xp.argmin(x.data, axis=axis, keepdims=keepdims), requires_grad=False, device=x.device, dtype=int32)


def add(x, y): return x.add(y)
def sub(x, y): return x.sub(y)
def mul(x, y): return x.mul(y)
def div(x, y): return x.div(y)
def matmul(x, y): return x.matmul(y)
def sum(x, axis=None, keepdims=False): return x.sum(axis=axis, keepdims=keepdims)
def mean(x, axis=None, keepdims=False): return x.mean(axis=axis, keepdims=keepdims)
def var(x, axis=None, keepdims=False): return x.var(axis=axis, keepdims=keepdims)
def power(x, y): return x.power(y)
def sqrt(x): return x.sqrt()
def log(x): return x.log()
def exp(x): return x.exp()
def tanh(x): return x.tanh()
def sin(x): return x.sin()
def cos(x): return x.cos()
def maximum(x, y): return x.maximum(y)
def minimum(x, y): return x.minimum(y)
def max(x, axis=None, keepdims=False): return x.max(axis=axis, keepdims=keepdims)
def min(x, axis=None, keepdims=False): return x.min(axis=axis, keepdims=keepdims)


def concatenate(*tensors, axis=0):
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tuple(tensors[0])
    return Tensor.concatenate(*tensors, axis=axis)


def reshape(x, *shape): return x.reshape(*shape)
def abs(x): return x.abs()
def transpose(x, *axes): return x.transpose(*axes)
def swapaxes(x, axis1, axis2): return x.swapaxes(axis1, axis2)
def flip(x, axis): return x.flip(axis=axis)


def where(condition, x, y):
    if not isinstance(x, Tensor):
        x = tensor(x, device=condition.device)
    return x.where(condition, y)


def equal(x, y): return x.equal(y)
def not_equal(x, y): return x.not_equal(y)
def greater(x, y): return x.greater(y)
def greater_equal(x, y): return x.greater_equal(y)
def less(x, y): return x.less(y)
def less_equal(x, y): return x.less_equal(y)
def logical_and(x, y): return x.logical_and(y)
def logical_or(x, y): return x.logical_or(y)
def logical_not(x): return x.logical_not()


def copy(x: Tensor) -> Tensor:
    return Tensor(x.data, requires_grad=x.requires_grad, device=x.device, dtype=x.dtype)


clone = copy

from . import nn, optim  # noqa: E402,F401
