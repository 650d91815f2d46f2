"""nn.Conv2d (NCHW cross-correlation with stride, dilation and asymmetric padding).

Reference: neunet/nn/layers/conv2d.py (forward 297-355, backward 16-117, geometry 193-295). The
reference pads a copy, dilates the kernel in place and runs three einsums over as_strided windows;
here the geometry is resolved once and the three contractions go to ``neunet.b200`` (implicit-GEMM
style on the tcgen05 kernel, direct kernels for tiny channel counts) on ``"cuda"``, or to a
tap-wise NumPy formulation on ``"cpu"``.
"""
from __future__ import annotations

import numpy as np

from ... import tensor as _tensor
from ...autograd import Tensor
from ..modules import Module
from ..parameter import Parameter


def _pair(v):
    return v if isinstance(v, tuple) else (v, v)


def resolve_padding(padding, kernel_size, stride, dilation, in_hw):
    """``padding`` -> (top, bottom, left, right), following Conv2d.build (conv2d.py:197-243)."""
    kh, kw = kernel_size
    if padding == "valid":
        return (0, 0, 0, 0)
    if padding in ("same", "real same"):
        if padding == "same":
            ud = dilation[0] * (kh - 1) - stride[0] + 1
            lr = dilation[1] * (kw - 1) - stride[1] + 1
        else:
            ud = (stride[0] - 1) * (in_hw[0] - 1) + dilation[0] * (kh - 1)
            lr = (stride[1] - 1) * (in_hw[1] - 1) + dilation[1] * (kw - 1)
        return (abs(ud // 2), abs(ud - ud // 2), abs(lr // 2), abs(lr - lr // 2))
    padding = tuple(padding)
    if len(padding) == 2:
        return (padding[0], padding[0], padding[1], padding[1])
    return padding


def out_hw(in_hw, kernel_size, stride, pad4, dilation):
    """conv2d.py:245-260."""
    return ((in_hw[0] + pad4[0] + pad4[1] - dilation[0] * (kernel_size[0] - 1) - 1) // stride[0] + 1,
            (in_hw[1] + pad4[2] + pad4[3] - dilation[1] * (kernel_size[1] - 1) - 1) // stride[1] + 1)


# ---- index-only helpers (bit-exact; conv2d.py:361-401) ---------------------------------------
def set_padding(array, padding):
    xp = np if isinstance(array, np.ndarray) else _xp_of(array)
    return xp.pad(array, ((0, 0), (0, 0), (padding[0], padding[1]), (padding[2], padding[3])), constant_values=0)


def remove_padding(array, padding):
    return array[:, :, padding[0]: array.shape[2] - padding[1], padding[2]: array.shape[3] - padding[3]]


def set_stride(array, stride):
    xp = np if isinstance(array, np.ndarray) else _xp_of(array)
    out = xp.zeros((array.shape[0], array.shape[1], stride[0] * array.shape[2] - (stride[0] - 1),
                    stride[1] * array.shape[3] - (stride[1] - 1)), dtype=array.dtype if isinstance(array, np.ndarray) else np.float32)
    out[:, :, :: stride[0], :: stride[1]] = array
    return out


def remove_stride(array, stride):
    return array[:, :, :: stride[0], :: stride[1]]


def _xp_of(array):
    from ...backend import get_xp
    return get_xp("cuda")


# ---- CPU formulation: sum over kernel taps of strided slices ----------------------------------
def _taps(kh, kw, stride, dilation, ho, wo):
    for k in range(kh):
        for l in range(kw):
            ys, xs = k * dilation[0], l * dilation[1]
            yield k, l, (slice(None), slice(None), slice(ys, ys + (ho - 1) * stride[0] + 1, stride[0]),
                         slice(xs, xs + (wo - 1) * stride[1] + 1, stride[1]))


def _cpu_forward(x, w, b, stride, pad4, dilation):
    kh, kw = w.shape[2:]
    ho, wo = out_hw(x.shape[2:], (kh, kw), stride, pad4, dilation)
    xp_ = set_padding(x, pad4)
    o = np.zeros((x.shape[0], w.shape[0], ho, wo), dtype=np.float32)
    for k, l, sl in _taps(kh, kw, stride, dilation, ho, wo):
        o += np.tensordot(xp_[sl], w[:, :, k, l], axes=([1], [1])).transpose(0, 3, 1, 2)
    if b is not None:
        o += b[None, :, None, None]
    return o


def _cpu_backward(x, w, g, stride, pad4, dilation, need_dx):
    kh, kw = w.shape[2:]
    ho, wo = g.shape[2:]
    xp_ = set_padding(x, pad4)
    dw = np.zeros_like(w)
    dxp = np.zeros_like(xp_) if need_dx else None
    for k, l, sl in _taps(kh, kw, stride, dilation, ho, wo):
        dw[:, :, k, l] = np.tensordot(g, xp_[sl], axes=([0, 2, 3], [0, 2, 3]))
        if need_dx:
            dxp[sl] += np.tensordot(g, w[:, :, k, l], axes=([1], [0])).transpose(0, 3, 1, 2)
    dx = None
    if need_dx:
        dx = np.ascontiguousarray(dxp[:, :, pad4[0]: pad4[0] + x.shape[2], pad4[2]: pad4[2] + x.shape[3]])
    return dx, dw


class _Conv2dTensor(Tensor):
    def __init__(self, data, args, op, device):
        t = Tensor._wrap(data, args, op, True, device)
        self.__dict__.update(t.__dict__)
        self.grad_fn = _conv2d_grad_fn


def _conv2d_grad_fn(X: Tensor, weight: Tensor, bias, stride, pad4, dilation, grad, x_planes=None):
    if X.device == "cuda":
        from ... import b200
        dx, dw, db = b200.conv2d_backward(X.data, weight.data, grad, stride, pad4, dilation,
                                          need_dx=X.requires_grad, need_db=bias is not None, x_planes=x_planes)
    else:
        dx, dw = _cpu_backward(X.data, weight.data, grad, stride, pad4, dilation, X.requires_grad)
        db = np.sum(grad, axis=(0, 2, 3)) if bias is not None else None
    if dx is not None:
        X.apply_grad(dx)
    weight.apply_grad(dw)
    if bias is not None:
        bias.apply_grad(db)


class Conv2d(Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=(1, 1), padding=(0, 0), dilation=(1, 1),
                 bias: bool = True, device="cpu"):
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        # NB: the reference wraps ANY non-tuple padding as (p, p) (conv2d.py:162), which makes its
        # string modes unreachable; here a string is kept as a string so "same"/"valid" work.
        self.padding = padding if isinstance(padding, (tuple, str)) else (padding, padding)
        self.stride = _pair(stride)
        self.dilation = _pair(dilation)
        stdv = 1.0 / np.sqrt(in_channels * self.kernel_size[0] * self.kernel_size[1])
        self.weight = Parameter(_tensor(
            np.random.uniform(-stdv, stdv, (out_channels, in_channels, *self.kernel_size)), dtype=np.float32))
        self.bias = Parameter(_tensor(np.zeros(out_channels), dtype=np.float32)) if bias else None
        self.input_size = None
        self.to(device)

    def forward(self, X: Tensor) -> Tensor:
        if not isinstance(X, Tensor):
            raise TypeError("Input must be a tensor")
        if X.device != self.device:
            raise ValueError("Tensors must be on the same device")
        self.input_size = X.shape
        pad4 = resolve_padding(self.padding, self.kernel_size, self.stride, self.dilation, X.shape[2:])
        b = self.bias
        planes = None
        if self.device == "cuda":
            from ... import b200
            # the channels-last bf16 planes of X made for the forward contraction are kept for wgrad
            O, planes = b200.conv2d_forward(X.data, self.weight.data, b.data if b is not None else None,
                                            self.stride, pad4, self.dilation, keep_planes=True)
        else:
            O = _cpu_forward(X.data, self.weight.data, b.data if b is not None else None,
                             self.stride, pad4, self.dilation)
        out = _Conv2dTensor(O, (X, self.weight, b, self.stride, pad4, self.dilation), "conv2d", self.device)
        if planes is not None:
            out.grad_fn = lambda X_, w_, b_, s_, p_, d_, grad: _conv2d_grad_fn(X_, w_, b_, s_, p_, d_, grad, x_planes=planes)
        return out

    def __call__(self, X):
        return self.forward(X)
