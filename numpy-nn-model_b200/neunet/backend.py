"""Array back-ends behind ``Tensor.xp``.

The reference picks ``numpy`` for ``device="cpu"`` and ``cupy`` for ``device="cuda"``
(neunet/autograd.py:9-14). Here ``"cpu"`` is NumPy and ``"cuda"`` is a NumPy-flavoured facade over
torch CUDA tensors (allocator, streams, element-wise / reduce / index ops). The dense contractions
never go through this facade's torch ops: ``matmul`` routes to the hand-written sm_100a kernels in
``neunet.b200`` and there is no CPU or torch fallback for it on ``"cuda"``.
"""
from __future__ import annotations

import numpy as np

try:  # torch is only needed for device="cuda"
    import torch
except Exception:  # pragma: no cover
    torch = None


_NP2T = {}
if torch is not None:
    _NP2T = {
        np.dtype(np.float32): torch.float32,
        np.dtype(np.float64): torch.float64,
        np.dtype(np.float16): torch.float16,
        np.dtype(np.int64): torch.int64,
        np.dtype(np.int32): torch.int32,
        np.dtype(np.int16): torch.int16,
        np.dtype(np.int8): torch.int8,
        np.dtype(np.uint8): torch.uint8,
        np.dtype(np.bool_): torch.bool,
    }
    _T2NP = {v: k for k, v in _NP2T.items()}


def to_torch_dtype(dtype):
    if dtype is None:
        return None
    if torch is not None and isinstance(dtype, torch.dtype):
        return dtype
    if dtype is float:
        return torch.float32
    if dtype is int:
        return torch.int64
    if dtype is bool:
        return torch.bool
    return _NP2T[np.dtype(dtype)]


def to_numpy_dtype(dtype):
    if torch is not None and isinstance(dtype, torch.dtype):
        return _T2NP[dtype]
    return np.dtype(dtype)


def _axis_kw(axis):
    if isinstance(axis, list):
        axis = tuple(axis)
    return axis


class _TorchRandom:
    """``xp.random`` used by Dropout etc. Draws on the host with NumPy's global generator (so seeds
    behave like the reference's CuPy/NumPy calls) and uploads."""

    def __init__(self, xp):
        self._xp = xp

    def rand(self, *shape):
        return self._xp.array(np.random.rand(*shape), dtype=np.float64)

    def randn(self, *shape):
        return self._xp.array(np.random.randn(*shape), dtype=np.float64)

    def uniform(self, low=0.0, high=1.0, size=None):
        return self._xp.array(np.random.uniform(low, high, size), dtype=np.float64)

    def normal(self, loc=0.0, scale=1.0, size=None):
        return self._xp.array(np.random.normal(loc, scale, size), dtype=np.float64)

    def binomial(self, n, p, size=None):
        # device-side Bernoulli: the reference's NumPy binomial costs 0.17 s/step on the GPT example
        if n == 1:
            # torch's default CUDA generator: seeded by torch.manual_seed and safe under graph capture
            return (torch.rand(tuple(size), device=self._xp.device) < p).to(torch.float32)
        return self._xp.array(np.random.binomial(n, p, size), dtype=np.float32)

    def randint(self, low, high=None, size=None):
        return self._xp.array(np.random.randint(low, high, size), dtype=np.int64)


class TorchXP:
    """Subset of the NumPy API the framework uses, implemented on torch CUDA tensors."""

    name = "torch-cuda"
    float32 = np.float32
    int32 = np.int32
    newaxis = None
    pi = np.pi

    def __init__(self):
        self._gen = None
        self.random = _TorchRandom(self)
        self.ndarray = torch.Tensor if torch is not None else type(None)

    # -- device ------------------------------------------------------------------------------
    @property
    def device(self):
        if torch is None or not torch.cuda.is_available():
            raise RuntimeError(
                'device="cuda" needs a CUDA device (B200, sm_100a) and torch with CUDA; '
                "there is no CPU fallback for the cuda path")
        return torch.device("cuda", torch.cuda.current_device())

    def generator(self):
        if self._gen is None or self._gen.device != self.device:
            self._gen = torch.Generator(device=self.device)
            self._gen.manual_seed(int(np.random.randint(0, 2 ** 31 - 1)))
        return self._gen

    # -- creation ----------------------------------------------------------------------------
    def array(self, data, dtype=None, copy=True):
        td = to_torch_dtype(dtype)
        if isinstance(data, torch.Tensor):
            out = data.to(device=self.device, dtype=td if td is not None else data.dtype)
            if copy and out.data_ptr() == data.data_ptr():
                out = out.clone()
            return out
        if isinstance(data, (bool, int, float, np.generic)) or (isinstance(data, np.ndarray) and data.ndim == 0):
            # scalars become a device fill (no host->device copy, so it is legal inside graph capture)
            val = data.item() if hasattr(data, "item") else data
            if td is None:
                td = torch.float64 if isinstance(val, float) else (torch.bool if isinstance(val, bool) else torch.int64)
            return torch.full((), val, dtype=td, device=self.device)
        if isinstance(data, np.ndarray):
            arr = data
        else:
            if isinstance(data, (list, tuple)) and any(isinstance(d, torch.Tensor) for d in data):
                return torch.stack([self.array(d, dtype) for d in data])
            arr = np.array(data)
        if dtype is not None:
            arr = arr.astype(to_numpy_dtype(dtype), copy=False)
        elif arr.dtype == np.float64:
            pass
        arr = np.ascontiguousarray(arr)
        return torch.from_numpy(arr).to(self.device)

    def asarray(self, data, dtype=None):
        return self.array(data, dtype=dtype, copy=False)

    def ascontiguousarray(self, a):
        return a.contiguous()

    def zeros(self, shape, dtype=np.float32):
        return torch.zeros(shape if isinstance(shape, (tuple, list)) else (shape,), dtype=to_torch_dtype(dtype), device=self.device)

    def ones(self, shape, dtype=np.float32):
        return torch.ones(shape if isinstance(shape, (tuple, list)) else (shape,), dtype=to_torch_dtype(dtype), device=self.device)

    def empty(self, shape, dtype=np.float32):
        return torch.empty(shape if isinstance(shape, (tuple, list)) else (shape,), dtype=to_torch_dtype(dtype), device=self.device)

    def full(self, shape, fill, dtype=np.float32):
        return torch.full(shape if isinstance(shape, (tuple, list)) else (shape,), fill, dtype=to_torch_dtype(dtype), device=self.device)

    def zeros_like(self, a, dtype=None):
        return torch.zeros_like(a, dtype=to_torch_dtype(dtype))

    def ones_like(self, a, dtype=None):
        return torch.ones_like(a, dtype=to_torch_dtype(dtype))

    def empty_like(self, a, dtype=None):
        return torch.empty_like(a, dtype=to_torch_dtype(dtype))

    def arange(self, start, end=None, step=1, dtype=None):
        if end is None:
            start, end = 0, start
        return torch.arange(start, end, step, dtype=to_torch_dtype(dtype), device=self.device)

    def eye(self, n, dtype=np.float32):
        return torch.eye(n, dtype=to_torch_dtype(dtype), device=self.device)

    # -- element-wise ------------------------------------------------------------------------
    def _t(self, a, like=None):
        if isinstance(a, torch.Tensor):
            return a
        dt = like.dtype if isinstance(like, torch.Tensor) else torch.float32
        if isinstance(a, (bool, int, float)):
            return torch.full((), a, dtype=dt, device=self.device)
        return torch.as_tensor(a, dtype=dt, device=self.device)

    def exp(self, a): return torch.exp(a)
    def log(self, a): return torch.log(a)
    def sqrt(self, a): return torch.sqrt(self._t(a))
    def tanh(self, a): return torch.tanh(a)
    def sin(self, a): return torch.sin(a)
    def cos(self, a): return torch.cos(a)
    def abs(self, a): return torch.abs(a)
    def sign(self, a): return torch.sign(a)
    def square(self, a): return a * a
    def power(self, a, b): return torch.pow(self._t(a, b), self._t(b, a))
    def maximum(self, a, b): return torch.maximum(self._t(a, b), self._t(b, a))
    def minimum(self, a, b): return torch.minimum(self._t(a, b), self._t(b, a))
    def clip(self, a, lo, hi): return torch.clamp(a, lo, hi)
    def logical_and(self, a, b): return torch.logical_and(a, b)
    def logical_or(self, a, b): return torch.logical_or(a, b)
    def logical_not(self, a): return torch.logical_not(a)
    def isnan(self, a): return torch.isnan(a)

    def where(self, cond, a, b):
        cond = self._t(cond).to(torch.bool) if not (isinstance(cond, torch.Tensor) and cond.dtype == torch.bool) else cond
        ref = a if isinstance(a, torch.Tensor) else b
        return torch.where(cond, self._t(a, ref), self._t(b, ref))

    # -- reductions --------------------------------------------------------------------------
    def sum(self, a, axis=None, keepdims=False):
        return a.sum() if axis is None and not keepdims else torch.sum(a, dim=_axis_kw(axis) if axis is not None else tuple(range(a.ndim)), keepdim=keepdims)

    def mean(self, a, axis=None, keepdims=False):
        return a.mean() if axis is None and not keepdims else torch.mean(a, dim=_axis_kw(axis) if axis is not None else tuple(range(a.ndim)), keepdim=keepdims)

    def var(self, a, axis=None, keepdims=False):
        dim = _axis_kw(axis) if axis is not None else tuple(range(a.ndim))
        return torch.var(a, dim=dim, keepdim=keepdims, unbiased=False)

    def max(self, a, axis=None, keepdims=False):
        if axis is None and not keepdims:
            return a.max()
        return torch.amax(a, dim=_axis_kw(axis) if axis is not None else tuple(range(a.ndim)), keepdim=keepdims)

    def min(self, a, axis=None, keepdims=False):
        if axis is None and not keepdims:
            return a.min()
        return torch.amin(a, dim=_axis_kw(axis) if axis is not None else tuple(range(a.ndim)), keepdim=keepdims)

    def argmax(self, a, axis=None, keepdims=False):
        return torch.argmax(a, dim=axis, keepdim=keepdims)

    def argmin(self, a, axis=None, keepdims=False):
        return torch.argmin(a, dim=axis, keepdim=keepdims)

    def prod(self, a, axis=None):
        return torch.prod(a) if axis is None else torch.prod(a, dim=axis)

    # -- shape -------------------------------------------------------------------------------
    def reshape(self, a, shape): return a.reshape(shape)
    def transpose(self, a, axes=None):
        if axes is None:
            axes = tuple(reversed(range(a.ndim)))
        return a.permute(*axes)
    def swapaxes(self, a, i, j): return a.swapaxes(i, j)
    def expand_dims(self, a, axis):
        if isinstance(axis, (tuple, list)):
            for ax in sorted(ax % (a.ndim + len(axis)) for ax in axis):
                a = a.unsqueeze(ax)
            return a
        return a.unsqueeze(axis)
    def squeeze(self, a, axis=None): return a.squeeze() if axis is None else a.squeeze(axis)
    def broadcast_to(self, a, shape): return a.expand(shape)
    def concatenate(self, arrays, axis=0): return torch.cat(list(arrays), dim=axis)
    def stack(self, arrays, axis=0): return torch.stack(list(arrays), dim=axis)
    def flip(self, a, axis):
        return torch.flip(a, dims=(axis,) if isinstance(axis, int) else tuple(axis))
    def rot90(self, a, k, axes): return torch.rot90(a, k, dims=axes)
    def triu(self, a, k=0): return torch.triu(a, diagonal=k)
    def tril(self, a, k=0): return torch.tril(a, diagonal=k)
    def outer(self, a, b): return torch.outer(a.reshape(-1), b.reshape(-1))
    def pad(self, a, pad_width, constant_values=0):
        flat = []
        for lo, hi in reversed(list(pad_width)):
            flat += [int(lo), int(hi)]
        return torch.nn.functional.pad(a, flat, value=constant_values)
    def copy(self, a): return a.clone()
    def astype(self, a, dtype): return a.to(to_torch_dtype(dtype))

    # -- the hot op ---------------------------------------------------------------------------
    def matmul(self, a, b):
        from . import b200
        return b200.matmul(a, b)

    def dot(self, a, b):
        return self.matmul(a, b)


_TORCH_XP = None


def get_xp(device):
    """``np`` for "cpu", the torch facade for "cuda" (mirrors neunet/autograd.py:9-14)."""
    global _TORCH_XP
    if device == "cpu":
        return np
    if device == "cuda":
        if torch is None:
            raise RuntimeError('device="cuda" needs torch')
        if _TORCH_XP is None:
            _TORCH_XP = TorchXP()
        return _TORCH_XP
    raise ValueError("Device must be 'cpu' or 'cuda'")


def is_device_array(a):
    return torch is not None and isinstance(a, torch.Tensor)


def astype(xp, a, dtype):
    return a.astype(dtype) if xp is np else xp.astype(a, dtype)


def to_host(a):
    """device array -> NumPy (used by .cpu(), state_dict pickles, .item())."""
    if is_device_array(a):
        return a.detach().cpu().numpy()
    return np.asarray(a)
