"""TEST-ONLY stand-in for the native library so the HOST-side logic of ``device="cuda"`` -- deferred
evaluation, fusion pattern matching, tape wiring, operand-plane caches (neunet/autograd.py, nn/layers) --
can be exercised by the ``-m "not gpu"`` suite in a container without a GPU.

``install()`` monkey-patches ``neunet.b200``'s Python entry points with plain torch-CPU formulas of the same
ops and points the "cuda" array back-end at CPU tensors. It is never imported by product code; the kernels
themselves are only ever validated on a real B200 (``-m gpu`` tests). ``uninstall()`` restores everything.
"""
import contextlib

import numpy as np
import torch

_saved = {}
calls = []  # names of the (mocked) native entry points, in call order: lets tests assert WHICH kernels a model would launch


def _mask(shape, p, ticket):
    seed, call_id, epoch, _dev = ticket
    g = torch.Generator().manual_seed((int(seed) * 1000003 + int(call_id) * 7919 + int(epoch)) % (2 ** 31))
    return (torch.rand(tuple(shape), generator=g) >= p).to(torch.float32) / (1.0 - p)


def _planes(t2d):
    from neunet import b200
    return (b200.planes_key(t2d), torch.empty(1, dtype=torch.uint8))


def install():
    import neunet.backend as be
    from neunet import b200

    if _saved:
        return
    names = ["require_device", "linear_forward", "linear_backward", "matmul", "matmul_backward", "softmax_forward",
             "softmax_backward", "swish_forward", "swish_backward", "rmsnorm_forward", "rmsnorm_backward", "dropout_apply",
             "attention_supported", "attention_forward", "attention_backward", "cross_entropy_forward",
             "cross_entropy_backward", "dropout_ticket", "_cache_scope", "conv2d_forward", "conv2d_backward",
             "conv_transpose2d_supported", "bn_forward", "bn_backward", "embedding_forward", "embedding_backward",
             "cross_entropy_linear_backward", "swish_dropout_apply"]
    for n in names:
        _saved[n] = getattr(b200, n, None)
    _saved["device_prop"] = be.TorchXP.device
    _saved["fused_adam"] = b200.FusedAdam
    be.TorchXP.device = property(lambda self: torch.device("cpu"))
    b200._cache_scope = lambda: (False, 0)
    b200.require_device = lambda: 0
    counter = {"epoch": 0}

    def dropout_ticket():
        counter["epoch"] += 1
        return (b200._rng["seed"], 0, counter["epoch"], None)
    b200.dropout_ticket = dropout_ticket
    b200._mock_reset_rng = lambda: counter.update(epoch=0)

    def linear_forward(x, w, bias=None, act=0, beta=1.0, save_z=False, owner=None, keep_x_staged=False, x_owner=None):
        cached = getattr(x_owner, "_b200_xst", None) if x_owner is not None else None
        x2 = x.reshape(-1, w.shape[1])
        hit = cached is not None and cached[0] == b200.planes_key(x2)
        calls.append("linear_forward_staged" if hit else "linear_forward")
        z = x @ w.T
        if bias is not None:
            z = z + bias.reshape(-1)
        o = z * torch.sigmoid(beta * z) if act else z
        return o, (z if save_z else None), None

    def linear_backward(x, w, grad, z=None, act=0, beta=1.0, need_dx=True, need_db=True, owner=None, x_staged=None,
                        dw_out=None, db_out=None, grad_drop=None):
        calls.append("linear_backward_dropped" if grad_drop is not None else "linear_backward")
        if grad._base is not None and grad.shape[-1] == w.shape[0] and w.shape[0] > 3 * 8 and grad.is_contiguous():
            calls.append("linear_backward_zero_copy_candidate")
        g = grad
        if grad_drop is not None:
            g = g * _mask(g.shape, grad_drop[0], grad_drop[1])
        if act:
            s = torch.sigmoid(beta * z)
            f = z * s
            g = g * (beta * f + s * (1 - beta * f))
        g2, x2 = g.reshape(-1, w.shape[0]), x.reshape(-1, w.shape[1])
        dx = (g2 @ w).reshape(x.shape) if need_dx else None
        dw = g2.T @ x2
        db = g2.sum(0, keepdim=True) if need_db else None
        return dx, dw, db

    def matmul(a, b, alpha=1.0, keep_staged=False):
        calls.append("matmul")
        out = torch.matmul(a.to(torch.float32), b.to(torch.float32)) * alpha
        return (out, None) if keep_staged else out

    def matmul_backward(a, b, grad, need_da=True, need_db=True, alpha=1.0, staged=None):
        calls.append("matmul_backward")
        da = torch.matmul(grad, b.transpose(-1, -2)) if need_da else None
        db = torch.matmul(a.transpose(-1, -2), grad) if need_db else None
        return da, db

    def softmax_forward(x, axis=-1):
        calls.append("softmax_forward")
        return torch.softmax(x, dim=axis)

    def softmax_backward(y, grad, axis=-1):
        calls.append("softmax_backward")
        return (grad - (grad * y).sum(axis, keepdim=True)) * y

    def swish_forward(x, beta=1.0):
        calls.append("swish_forward")
        return x * torch.sigmoid(beta * x)

    def swish_backward(x, grad, beta=1.0):
        calls.append("swish_backward")
        s = torch.sigmoid(beta * x)
        f = x * s
        return grad * (beta * f + s * (1 - beta * f))

    def dropout_apply(x, p, ticket, residual=None, want_planes=False):
        calls.append("dropout_fused" if residual is not None else "dropout")
        y = x * _mask(x.shape, p, ticket)
        if residual is not None:
            y = y + residual
        if want_planes:
            return y, (_planes(y.reshape(-1, y.shape[-1])) if y.shape[-1] % 8 == 0 else None)
        return y

    def swish_dropout_apply(z, beta, p, ticket, want_planes=True):
        calls.append("swish_dropout")
        y = z * torch.sigmoid(beta * z) * _mask(z.shape, p, ticket)
        return y, (_planes(y.reshape(-1, y.shape[-1])) if (want_planes and y.shape[-1] % 8 == 0) else None)

    def rmsnorm_forward(x, w, b=None, eps=1e-6, add_dropout=None, want_planes=False):
        calls.append("rmsnorm_forward_fused" if add_dropout is not None else "rmsnorm_forward")
        s_out = None
        if add_dropout is not None:
            a, p, ticket = add_dropout
            x = x + a * _mask(a.shape, p, ticket)
            s_out = x
        std = torch.sqrt((x * x).mean(-1, keepdim=True) + eps)
        y = x / std * w
        if b is not None:
            y = y + b
        planes = _planes(y.reshape(-1, y.shape[-1])) if (want_planes and y.shape[-1] % 8 == 0) else None
        return y, std, s_out, planes

    def rmsnorm_backward(grad, x, w, std, need_db=False, dx_add=None):
        calls.append("rmsnorm_backward_acc" if dx_add is not None else "rmsnorm_backward")
        n = x.shape[-1]
        dxh = w * grad
        dx = (dxh * std - x * (dxh * x / std).sum(-1, keepdim=True) / n) / std ** 2
        lead = tuple(range(grad.ndim - 1))
        dw = (grad * (x / std)).sum(lead)
        db = grad.sum(lead) if need_db else None
        if dx_add is not None:
            dx = dx + dx_add
        return dx, dw, db, dx_add is not None

    def attention_supported(Tq, Tk, D):
        return Tq <= 64 and Tk <= 64 and D <= 64 and Tk % 4 == 0 and D % 4 == 0

    def _masked(mask, shape4):
        if mask is None:
            return None
        t, kind, cmp = mask
        m = (t != 0) if kind == 1 else (t.to(torch.float32) == cmp)
        return m.expand(shape4)

    def _attn_probs(q, kT, mask, fill, scale):
        s = torch.matmul(q, kT) / scale
        m = _masked(mask, s.shape)
        if m is not None:
            s = torch.where(m, torch.full((), fill, dtype=s.dtype), s)
        return torch.softmax(s, dim=-1), m

    def attention_forward(q, kT, v, mask, fill, scale, p, ticket, want_planes=False):
        calls.append("attention_forward")
        B, H, Tq, D = q.shape
        pr, _ = _attn_probs(q, kT, mask, fill, scale)
        attn = pr * _mask(pr.shape, p, ticket) if p > 0 else pr
        out = torch.matmul(attn, v).permute(0, 2, 1, 3).contiguous()  # (B, Tq, H, D) memory order
        planes = _planes(out.reshape(B * Tq, H * D)) if want_planes and (H * D) % 8 == 0 else None
        return out.permute(0, 2, 1, 3), attn, planes

    def attention_backward(q, kT, v, mask, fill, scale, p, ticket, grad):
        calls.append("attention_backward")
        pr, m = _attn_probs(q, kT, mask, fill, scale)
        dm = _mask(pr.shape, p, ticket) if p > 0 else torch.ones_like(pr)
        pd = pr * dm
        dv = torch.matmul(pd.transpose(-1, -2), grad)
        dp = torch.matmul(grad, v.transpose(-1, -2)) * dm
        ds = (dp - (dp * pr).sum(-1, keepdim=True)) * pr
        if m is not None:
            ds = torch.where(m, torch.zeros((), dtype=ds.dtype), ds)
        ds = ds / scale
        dq = torch.matmul(ds, kT.transpose(-1, -2))
        dkT = torch.matmul(q.transpose(-1, -2), ds)
        B, H, Tq, D = q.shape
        if Tq == kT.shape[3]:  # packed dq | dk | dv like the real binding
            packed = torch.empty((B, Tq, 3, H, D), dtype=torch.float32)
            packed[:, :, 0] = dq.permute(0, 2, 1, 3)
            packed[:, :, 1] = dkT.permute(0, 3, 1, 2)
            packed[:, :, 2] = dv.permute(0, 2, 1, 3)
            calls.append("attention_backward_packed")
            return packed[:, :, 0].permute(0, 2, 1, 3), packed[:, :, 1].permute(0, 2, 3, 1), packed[:, :, 2].permute(0, 2, 1, 3)
        return dq, dkT, dv

    def cross_entropy_forward(logits, targets, ignore_index=-100, reduction="mean"):
        calls.append("cross_entropy_forward")
        tgt = targets.reshape(-1).to(torch.int64)
        lse = torch.logsumexp(logits, dim=1)
        keep = tgt != ignore_index
        safe = torch.where(keep, tgt, torch.zeros_like(tgt))
        row = torch.where(keep, lse - logits.gather(1, safe[:, None])[:, 0], torch.zeros_like(lse))
        if reduction == "none":
            return row, (logits, tgt, lse, None, int(ignore_index))
        inv = (1.0 / keep.sum().to(torch.float32)) if reduction == "mean" else torch.ones(())
        return row.sum() * inv, (logits, tgt, lse, inv.reshape(1), int(ignore_index))

    def cross_entropy_backward(saved, upstream):
        calls.append("cross_entropy_backward")
        logits, tgt, lse, inv, ignore_index = saved
        keep = (tgt != ignore_index)
        sm = torch.exp(logits - lse[:, None])
        onehot = torch.zeros_like(sm)
        onehot[torch.arange(len(tgt)), torch.where(keep, tgt, torch.zeros_like(tgt))] = 1.0
        d = (sm - onehot) * keep[:, None]
        up = upstream.reshape(-1)
        if inv is None and up.numel() == len(tgt) and len(tgt) > 1:
            return d * up[:, None]
        return d * (up[0] * (inv[0] if inv is not None else 1.0))

    import torch.nn.functional as F

    def conv2d_forward(x, w, bias, stride, pad4, dil, keep_planes=False):
        calls.append("conv2d_forward")
        xp = F.pad(x, (pad4[2], pad4[3], pad4[0], pad4[1]))
        o = F.conv2d(xp, w, bias.reshape(-1) if bias is not None else None, stride=tuple(stride), dilation=tuple(dil))
        return (o, None) if keep_planes else o

    def conv2d_backward(x, w, grad, stride, pad4, dil, need_dx=True, need_db=True, x_planes=None):
        calls.append("conv2d_backward")
        xl, wl = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        conv2d_forward(xl, wl, None, stride, pad4, dil).backward(grad)
        calls.pop()
        return (xl.grad if need_dx else None), wl.grad, (grad.sum((0, 2, 3)) if need_db else None)

    def bn_forward(x, w, b, alpha, eps, momentum, running_mean=None, running_var=None, stats=None):
        calls.append("bn_forward_lrelu" if alpha != 1.0 else "bn_forward")
        a = torch.where(x <= 0, alpha * x, x)
        if stats is None:
            mean, var = a.mean((0, 2, 3)), a.var((0, 2, 3), unbiased=False)
            inv = 1 / torch.sqrt(var + eps)
            if running_mean is not None:
                running_mean.mul_(momentum).add_((1 - momentum) * mean.reshape(running_mean.shape))
                running_var.mul_(momentum).add_((1 - momentum) * var.reshape(running_var.shape))
        else:
            mean, inv = stats
        y = (a - mean.reshape(1, -1, 1, 1)) * inv.reshape(1, -1, 1, 1)
        if w is not None:
            y = y * w.reshape(1, -1, 1, 1) + b.reshape(1, -1, 1, 1)
        return y, mean, inv

    def bn_backward(x, grad, mean, inv, w, alpha, need_dx=True, need_dw=True):
        calls.append("bn_backward")
        a = torch.where(x <= 0, alpha * x, x)
        xh = (a - mean.reshape(1, -1, 1, 1)) * inv.reshape(1, -1, 1, 1)
        n = x.shape[0] * x.shape[2] * x.shape[3]
        wg = grad * (w.reshape(1, -1, 1, 1) if w is not None else 1)
        s1, s2 = wg.sum((0, 2, 3), keepdim=True), (wg * xh).sum((0, 2, 3), keepdim=True)
        dx = torch.where(x <= 0, alpha, 1.0) * inv.reshape(1, -1, 1, 1) * (wg - s1 / n - xh * s2 / n)
        return dx, (grad * xh).sum((0, 2, 3)), grad.sum((0, 2, 3))

    def cross_entropy_linear_backward(saved, upstream, x, w, need_dx=True, need_db=True, owner=None, x_staged=None,
                                      dw_out=None, db_out=None):
        calls.append("cross_entropy_linear_backward")
        d = cross_entropy_backward(saved, upstream)
        calls.pop()
        out = linear_backward(x, w, d, need_dx=need_dx, need_db=need_db)
        calls.pop()
        return out
    b200.cross_entropy_linear_backward = cross_entropy_linear_backward

    def embedding_forward(weight, ids):
        calls.append("embedding_forward")
        return weight[ids.to(torch.int64)]

    def embedding_backward(ids, grad, V, out=None):
        calls.append("embedding_backward")
        flat = ids.reshape(-1).to(torch.int64)
        g2 = grad.reshape(flat.numel(), -1)
        dw = torch.zeros((V, g2.shape[1]), dtype=torch.float32)
        for i in range(flat.numel()):  # assignment semantics: the last duplicate wins
            dw[flat[i]] = g2[i]
        return dw

    b200.embedding_forward, b200.embedding_backward = embedding_forward, embedding_backward
    b200.conv2d_forward, b200.conv2d_backward = conv2d_forward, conv2d_backward
    b200.conv_transpose2d_supported = lambda *a: False
    b200.bn_forward, b200.bn_backward = bn_forward, bn_backward

    for n, f in dict(linear_forward=linear_forward, linear_backward=linear_backward, matmul=matmul,
                     matmul_backward=matmul_backward, softmax_forward=softmax_forward, softmax_backward=softmax_backward,
                     swish_forward=swish_forward, swish_backward=swish_backward, dropout_apply=dropout_apply,
                     swish_dropout_apply=swish_dropout_apply,
                     rmsnorm_forward=rmsnorm_forward, rmsnorm_backward=rmsnorm_backward,
                     attention_supported=attention_supported, attention_forward=attention_forward,
                     attention_backward=attention_backward, cross_entropy_forward=cross_entropy_forward,
                     cross_entropy_backward=cross_entropy_backward).items():
        setattr(b200, n, f)


def uninstall():
    import neunet.backend as be
    from neunet import b200
    if not _saved:
        return
    be.TorchXP.device = _saved.pop("device_prop")
    b200.FusedAdam = _saved.pop("fused_adam")
    for n, f in list(_saved.items()):
        if f is not None:
            setattr(b200, n, f)
    _saved.clear()
    calls.clear()


@contextlib.contextmanager
def mocked():
    install()
    try:
        yield
    finally:
        uninstall()


def to_np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
