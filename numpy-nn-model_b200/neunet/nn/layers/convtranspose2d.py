"""nn.ConvTranspose2d with the reference's unusual conventions (convtranspose2d.py:115-387):
weights are (out, in, kh, kw) and are NOT flipped; the layer is defined as a stride-1
cross-correlation over the zero-stuffed, (k-1)-padded, padding-cropped input. The correlation
itself is the Conv2d hot path; only the index-only input preparation lives here.
"""
from __future__ import annotations

import numpy as np

from ... import tensor as _tensor
from ...autograd import Tensor
from ..modules import Module
from ..parameter import Parameter
from .conv2d import _conv2d_grad_fn, _cpu_forward, _pair


class _ConvTranspose2dTensor(Tensor):
    def __init__(self, data, args, op, device):
        t = Tensor._wrap(data, args, op, True, device)
        self.__dict__.update(t.__dict__)
        self.grad_fn = _convT_grad_fn


def _convT_native_grad_fn(X: Tensor, weight, bias, stride, pad4, dilation, out_pad, planes, grad):
    from ... import b200
    dx, dw, db = b200.conv_transpose2d_backward(X.data, weight.data, grad, stride, pad4, dilation, out_pad,
                                                need_dx=X.requires_grad, need_db=bias is not None, x_planes=planes)
    if dx is not None:
        X.apply_grad(dx)
    weight.apply_grad(dw)
    if bias is not None:
        bias.apply_grad(db)


def _convT_grad_fn(X: Tensor, Xprep: Tensor, weight, bias, dilation, unprepare, grad):
    # gradient w.r.t. the prepared input / weight / bias via the Conv2d backward ...
    _conv2d_grad_fn(Xprep, weight, bias, (1, 1), (0, 0, 0, 0), dilation, grad)
    # ... then undo the index-only preparation (convtranspose2d.py:16-33)
    if X.requires_grad and Xprep.grad is not None:
        X.apply_grad(unprepare(Xprep.grad))


class ConvTranspose2d(Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=(1, 1), padding=(0, 0), dilation=(1, 1),
                 output_padding=(0, 0), bias: bool = True, device="cpu"):
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.padding = _pair(padding)
        self.stride = _pair(stride)
        self.dilation = _pair(dilation)
        self.output_padding = _pair(output_padding)
        stdv = 1.0 / np.sqrt(in_channels * self.kernel_size[0] * self.kernel_size[1])
        self.weight = Parameter(_tensor(
            np.random.uniform(-stdv, stdv, (out_channels, in_channels, *self.kernel_size)), dtype=np.float32))
        self.bias = Parameter(_tensor(np.zeros(out_channels), dtype=np.float32)) if bias else None
        self.input_size = None
        self.to(device)

    def _pad4(self):
        p = self.padding
        return (p[0], p[0], p[1], p[1]) if len(p) == 2 else tuple(p)

    def forward(self, X: Tensor) -> Tensor:
        if not isinstance(X, Tensor):
            raise TypeError("Input must be a tensor")
        if X.device != self.device:
            raise ValueError("Tensors must be on the same device")
        self.input_size = X.shape
        xp = X.xp
        B, C, H, W = X.shape
        s, op, pad4 = self.stride, self.output_padding, self._pad4()
        if X.device == "cuda":
            from ... import b200
            if b200.conv_transpose2d_supported(X.shape, self.weight.shape, s, pad4, self.dilation, op):
                # gather form, real taps only (no zero-stuffed copy, 1/(s0*s1) of the multiplies)
                bd = self.bias.data if self.bias is not None else None
                O, planes = b200.conv_transpose2d_forward(X.data, self.weight.data, bd, s, pad4, self.dilation, op)
                out = _ConvTranspose2dTensor(O, (X, self.weight, self.bias, s, pad4, self.dilation, op, planes),
                                             "convtranspose2d", self.device)
                out.grad_fn = _convT_native_grad_fn
                return out
        dk = (self.dilation[0] * (self.kernel_size[0] - 1) + 1, self.dilation[1] * (self.kernel_size[1] - 1) + 1)
        hs, ws = s[0] * H - (s[0] - 1) + op[0], s[1] * W - (s[1] - 1) + op[1]
        full = xp.zeros((B, C, hs + 2 * (dk[0] - 1), ws + 2 * (dk[1] - 1)), dtype=np.float32)
        full[:, :, dk[0] - 1: dk[0] - 1 + s[0] * (H - 1) + 1: s[0], dk[1] - 1: dk[1] - 1 + s[1] * (W - 1) + 1: s[1]] = X.data
        y0, y1 = pad4[0], full.shape[2] - pad4[1]
        x0, x1 = pad4[2], full.shape[3] - pad4[3]
        prep = full[:, :, y0:y1, x0:x1]
        if X.device == "cuda":
            prep = prep.contiguous()
        else:
            prep = np.ascontiguousarray(prep)
        Xprep = Tensor._wrap(prep, None, "convT_prepare", X.requires_grad, X.device)

        def unprepare(gprep):
            gfull = xp.zeros(tuple(full.shape), dtype=np.float32)
            gfull[:, :, y0:y1, x0:x1] = gprep
            return gfull[:, :, dk[0] - 1: dk[0] - 1 + s[0] * (H - 1) + 1: s[0],
                         dk[1] - 1: dk[1] - 1 + s[1] * (W - 1) + 1: s[1]]

        b = self.bias
        if self.device == "cuda":
            from ... import b200
            O = b200.conv2d_forward(prep, self.weight.data, b.data if b is not None else None, (1, 1), (0, 0, 0, 0),
                                    self.dilation)
        else:
            O = _cpu_forward(prep, self.weight.data, b.data if b is not None else None, (1, 1), (0, 0, 0, 0),
                             self.dilation)
        return _ConvTranspose2dTensor(O, (X, Xprep, self.weight, b, self.dilation, unprepare), "convtranspose2d",
                                      self.device)

    def __call__(self, X):
        return self.forward(X)
