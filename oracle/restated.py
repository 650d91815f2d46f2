"""CPU restatement (plain NumPy, fp32) of the reference's dense hot path.

TEST INFRASTRUCTURE ONLY. Nothing under ``numpy-nn-model_b200/`` imports this module; it is used
by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs as the checker and as the timed CPU baseline ("port").

Parity pinning: every function below is checked in ``tests/test_oracle_golden.py`` against
``tests/golden/*.npz``, which ``oracle/make_golden.py`` produced by importing the UNMODIFIED
reference tree (``/root/reference``, with a ``cupy`` stub because the package imports CuPy
unconditionally) and running its own ``Tensor.matmul`` / ``nn.Linear`` / ``nn.Conv2d`` /
``optim.Adam(W)`` / ``nn.Swish`` / ``nn.Softmax`` / ``nn.RMSNorm`` on seeded inputs, and against the
README autograd known-answer vector (README.md:144-170).

Each function cites the reference lines it follows. The contraction itself is NumPy's
``matmul`` (OpenBLAS), the third-party routine the reference calls (pinned numpy 1.24.0 in
pyproject.toml:10; numpy 2.3 here -- BLAS blocking differences are ~1e-7, far below the 1e-4 bar).
The convolution is restated as a sum over kernel taps of strided slices -- the same arithmetic as
the reference's as_strided+einsum, in a form that does not share its indexing code.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


# ----------------------------------------------------------------------------------------------
# autograd plumbing
# ----------------------------------------------------------------------------------------------
def reverse_broadcast(grad: np.ndarray, shape: tuple) -> np.ndarray:
    """Sum ``grad`` down to ``shape`` (neunet/autograd.py:948-962)."""
    if grad.shape == tuple(shape):
        return grad
    if len(shape) == grad.ndim:
        axes = tuple(i for i, (a, b) in enumerate(zip(shape, grad.shape)) if a != b)
        grad = grad.sum(axes, keepdims=True)
    else:
        padded = (1,) * (grad.ndim - len(shape)) + tuple(shape)
        axes = tuple(i for i, (a, b) in enumerate(zip(padded, grad.shape)) if a != b)
        grad = grad.sum(axis=axes)
    return grad.reshape(shape)


# ----------------------------------------------------------------------------------------------
# Tensor.matmul  (neunet/autograd.py:192-230)
# ----------------------------------------------------------------------------------------------
def matmul_forward(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.matmul(a, b)  # autograd.py:199


def matmul_backward(a: np.ndarray, b: np.ndarray, g: np.ndarray):
    """Returns (dA, dB) already reduced to the operand shapes (apply_grad, autograd.py:85-93)."""
    if a.ndim > 1 and b.ndim > 1:  # autograd.py:207-211
        da = np.matmul(g, np.swapaxes(b, -1, -2))
        db = np.matmul(np.swapaxes(a, -1, -2), g)
    elif a.ndim == 1 and b.ndim == 1:  # 212-216
        da, db = g * b, g * a
    elif a.ndim == 1:  # vector x matrix, 217-221
        da = np.matmul(g, np.swapaxes(b, -1, -2))
        db = np.outer(a, g)
    else:  # matrix x vector, 222-226
        da = np.outer(g, b)
        db = np.matmul(np.swapaxes(a, -1, -2), g)
    return reverse_broadcast(da, a.shape), reverse_broadcast(db, b.shape)


# ----------------------------------------------------------------------------------------------
# nn.Linear  (neunet/nn/layers/linear.py)
# ----------------------------------------------------------------------------------------------
def linear_init(in_features: int, out_features: int, bias: bool = True):
    """Same draws, in the same order, from the global np.random (linear.py:34-43)."""
    stdv = 1.0 / np.sqrt(in_features)
    w = np.random.uniform(-stdv, stdv, (out_features, in_features)).astype(f32)
    b = np.random.uniform(-stdv, stdv, (1, out_features)).astype(f32) if bias else None
    return w, b


def linear_forward(x, w, b=None):
    o = np.matmul(x, w.T)  # linear.py:54
    if b is not None:
        o = o + b  # linear.py:56
    return o.astype(f32)


def linear_backward(x, w, b, g):
    """linear.py:17-24 followed by apply_grad's un-broadcast (3-D inputs give a batched dW that is
    summed over the batch axis afterwards, autograd.py:955-958)."""
    dx = np.matmul(g, w)
    dw = np.swapaxes(np.matmul(np.swapaxes(x, -1, -2), g), -1, -2)
    dw = reverse_broadcast(dw, w.shape)
    db = None
    if b is not None:
        db = reverse_broadcast(np.sum(g, axis=0, keepdims=True), b.shape)
    return dx, dw, db


# ----------------------------------------------------------------------------------------------
# Swish / Softmax / RMSNorm epilogues
# ----------------------------------------------------------------------------------------------
def _sigmoid(x):
    return 1 / (1 + np.exp(-x))


def swish_forward(x, beta=1.0):
    return x * _sigmoid(beta * x)  # activations.py:225-228


def swish_backward(x, g, beta=1.0):
    f = x * _sigmoid(beta * x)
    return g * (beta * f + _sigmoid(beta * x) * (1 - beta * f))  # activations.py:212-216


def softmax_forward(x, axis=-1):
    e = np.exp(x - np.max(x, axis=axis, keepdims=True))  # activations.py:453
    return e / np.sum(e, axis=axis, keepdims=True)  # 455


def softmax_backward(f, g, axis=-1):
    return (g - (g * f).sum(axis, keepdims=True)) * f  # activations.py:442


def rmsnorm_forward(x, w, b=None, eps=1e-6):
    """rmsnorm.py:84-94. Returns (y, x_norm, x_std)."""
    std = np.sqrt(np.mean(x ** 2, -1, keepdims=True) + eps)
    xn = x / std
    y = xn * w
    if b is not None:
        y = y + b
    return y, xn, std


def rmsnorm_backward(x, w, b, xn, std, g):
    """rmsnorm.py:43-59 + apply_grad un-broadcast. Returns (dx, dw, db)."""
    n = x.shape[-1]
    dxh = w * g
    dx = (dxh * std - x * np.sum(dxh * x / std, axis=-1, keepdims=True) / n) / std ** 2
    dw = reverse_broadcast(np.sum(g * xn, axis=0), w.shape)
    db = reverse_broadcast(np.sum(g, axis=0), b.shape) if b is not None else None
    return dx, dw, db


# ----------------------------------------------------------------------------------------------
# nn.Conv2d  (neunet/nn/layers/conv2d.py)
# ----------------------------------------------------------------------------------------------
def conv2d_init(cin, cout, kernel_size, bias=True):
    """conv2d.py:166-186: weight U(+-1/sqrt(cin*kh*kw)) from the global np.random, bias zeros."""
    kh, kw = kernel_size if isinstance(kernel_size, tuple) else (kernel_size, kernel_size)
    stdv = 1.0 / np.sqrt(cin * kh * kw)
    w = np.random.uniform(-stdv, stdv, (cout, cin, kh, kw)).astype(f32)
    b = np.zeros(cout, dtype=f32) if bias else None
    return w, b


def conv2d_padding(padding, kernel_size, stride, dilation, in_hw):
    """Resolve Conv2d's `padding` argument to (top, bottom, left, right) (conv2d.py:197-243)."""
    kh, kw = kernel_size
    h, w = in_hw
    if padding == "valid":
        return (0, 0, 0, 0)
    if padding in ("same", "real same"):
        if padding == "same":
            ud = dilation[0] * (kh - 1) - stride[0] + 1
            lr = dilation[1] * (kw - 1) - stride[1] + 1
        else:
            ud = (stride[0] - 1) * (h - 1) + dilation[0] * (kh - 1)
            lr = (stride[1] - 1) * (w - 1) + dilation[1] * (kw - 1)
        up, down = ud // 2, ud - ud // 2
        left, right = lr // 2, lr - lr // 2
        return (abs(up), abs(down), abs(left), abs(right))
    if len(padding) == 2:
        return (padding[0], padding[0], padding[1], padding[1])
    return tuple(padding)


def conv2d_out_hw(in_hw, kernel_size, stride, pad4, dilation):
    """conv2d.py:245-260."""
    ho = (in_hw[0] + pad4[0] + pad4[1] - dilation[0] * (kernel_size[0] - 1) - 1) // stride[0] + 1
    wo = (in_hw[1] + pad4[2] + pad4[3] - dilation[1] * (kernel_size[1] - 1) - 1) // stride[1] + 1
    return ho, wo


def set_padding(a, pad4):
    """conv2d.py:361-368."""
    return np.pad(a, ((0, 0), (0, 0), (pad4[0], pad4[1]), (pad4[2], pad4[3])), constant_values=0)


def remove_padding(a, pad4):
    """conv2d.py:371-378."""
    return a[:, :, pad4[0]: a.shape[2] - pad4[1], pad4[2]: a.shape[3] - pad4[3]]


def set_stride(a, s):
    """Zero-stuffing (conv2d.py:381-396)."""
    out = np.zeros((a.shape[0], a.shape[1], s[0] * a.shape[2] - (s[0] - 1), s[1] * a.shape[3] - (s[1] - 1)), a.dtype)
    out[:, :, :: s[0], :: s[1]] = a
    return out


def remove_stride(a, s):
    """conv2d.py:399-401."""
    return a[:, :, :: s[0], :: s[1]]


def conv2d_forward(x, w, b, stride=(1, 1), pad4=(0, 0, 0, 0), dilation=(1, 1)):
    """O[b,o,h,w] = sum_{i,k,l} Xpad[b,i,h*s+k*d,w*s+l*d] * W[o,i,k,l] + bias[o]
    (conv2d.py:306-335: pad, dilate the kernel, window, einsum 'bihwkl,oikl->bohw', add bias)."""
    bsz, cin, h, wd = x.shape
    cout, _, kh, kw = w.shape
    ho, wo = conv2d_out_hw((h, wd), (kh, kw), stride, pad4, dilation)
    xp = set_padding(x, pad4)
    o = np.zeros((bsz, cout, ho, wo), dtype=f32)
    for k in range(kh):
        for l in range(kw):
            ys, xs = k * dilation[0], l * dilation[1]
            win = xp[:, :, ys: ys + (ho - 1) * stride[0] + 1: stride[0], xs: xs + (wo - 1) * stride[1] + 1: stride[1]]
            # (b, i, ho, wo) x (o, i) -> (b, o, ho, wo)
            o += np.einsum("bihw,oi->bohw", win, w[:, :, k, l], optimize=True).astype(f32)
    if b is not None:
        o += b[None, :, None, None]
    return o


def conv2d_backward(x, w, b, g, stride=(1, 1), pad4=(0, 0, 0, 0), dilation=(1, 1)):
    """dW (conv2d.py:93), db (94), dX (35-106). Returns (dx, dw, db).

    dX is the scatter of every window contribution back onto the padded input, cropped to the
    input; positions the strided windows never touch keep gradient 0 -- what the reference's
    set_padding(.., input - stride_compared) / remove_padding fix-ups (97-106) produce."""
    bsz, cin, h, wd = x.shape
    cout, _, kh, kw = w.shape
    ho, wo = g.shape[2], g.shape[3]
    xp = set_padding(x, pad4)
    dxp = np.zeros_like(xp)
    dw = np.zeros_like(w)
    for k in range(kh):
        for l in range(kw):
            ys, xs = k * dilation[0], l * dilation[1]
            sl = (slice(None), slice(None), slice(ys, ys + (ho - 1) * stride[0] + 1, stride[0]),
                  slice(xs, xs + (wo - 1) * stride[1] + 1, stride[1]))
            dw[:, :, k, l] = np.einsum("bihw,bohw->oi", xp[sl], g, optimize=True)
            dxp[sl] += np.einsum("bohw,oi->bihw", g, w[:, :, k, l], optimize=True).astype(f32)
    dx = dxp[:, :, pad4[0]: pad4[0] + h, pad4[2]: pad4[2] + wd]
    db = np.sum(g, axis=(0, 2, 3)) if b is not None else None
    return np.ascontiguousarray(dx), dw, db


# ----------------------------------------------------------------------------------------------
# optimizers  (neunet/optim.py)
# ----------------------------------------------------------------------------------------------
def adam_step(p, g, m, v, t, lr=0.01, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """One Adam update of one tensor (optim.py:17-33); t is the 1-based step. Returns (p, m, v).
    Hyper-parameters are Python floats as in the reference (NumPy scalars would promote to fp64)."""
    lr, eps, weight_decay, betas = float(lr), float(eps), float(weight_decay), (float(betas[0]), float(betas[1]))
    if weight_decay != 0:
        g = g + weight_decay * p
    m = betas[0] * m + (1 - betas[0]) * g
    v = betas[1] * v + (1 - betas[1]) * g ** 2
    m_hat = m / (1 - betas[0] ** t)
    v_hat = v / (1 - betas[1] ** t)
    p = p - lr * m_hat / (np.sqrt(v_hat) + eps)
    return p.astype(f32), m.astype(f32), v.astype(f32)


def adamw_step(p, g, m, v, t, lr=0.01, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
    """One AdamW update (optim.py:52-69): decoupled decay first, then the Adam update."""
    lr, eps, weight_decay, betas = float(lr), float(eps), float(weight_decay), (float(betas[0]), float(betas[1]))
    if weight_decay != 0:
        p = p - lr * weight_decay * p
    m = betas[0] * m + (1 - betas[0]) * g
    v = betas[1] * v + (1 - betas[1]) * g ** 2
    m_hat = m / (1 - betas[0] ** t)
    v_hat = v / (1 - betas[1] ** t)
    p = p - lr * m_hat / (np.sqrt(v_hat) + eps)
    return p.astype(f32), m.astype(f32), v.astype(f32)


# ----------------------------------------------------------------------------------------------
# losses used by the benchmark steps (not on the hot path; neunet/nn/losses.py)
# ----------------------------------------------------------------------------------------------
def cross_entropy(logits, labels, ignore_index=-100):
    """CrossEntropyLoss = LogSoftmax(axis=1) + NLLLoss(mean) (losses.py:59-126, activations.py:480-489).
    Returns (loss, dlogits)."""
    mx = np.max(logits, axis=1, keepdims=True)
    e = np.exp(logits - mx)
    logp = logits - mx - np.log(np.sum(e, axis=1, keepdims=True))
    keep = labels != ignore_index
    n = max(int(keep.sum()), 1)
    idx = np.arange(logits.shape[0])
    safe = np.where(keep, labels, 0)
    loss = -(logp[idx, safe] * keep).sum() / n
    gl = np.zeros_like(logits)
    gl[idx, safe] = -keep.astype(f32) / n
    dlogits = gl - np.exp(logp) * gl.sum(axis=1, keepdims=True)
    return f32(loss), dlogits.astype(f32)


def mse(pred, target):
    """MSELoss mean (losses.py:8-22). Returns (loss, dpred)."""
    d = pred - target
    return f32(np.mean(d ** 2)), (2.0 * d / d.size).astype(f32)
