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
        from ...autograd import _MaskedGrad
        grad_drop = None
        if type(grad) is _MaskedGrad:  # the upstream nn.Dropout left its mask to this layer's staging pass
            grad, grad_drop = grad.raw, (grad.p, grad.ticket)
        # data-parallel training: the first (normally only) gradient contribution of a parameter is written
        # straight into its slice of the all-reduce bucket (neunet.distributed.GradBucket.overlap_backward)
        dw_out = getattr(weight, "_grad_buffer", None) if weight.grad is None else None
        db_out = getattr(bias, "_grad_buffer", None) if (bias is not None and bias.grad is None) else None
        dx, dw, db = b200.linear_backward(X.data, weight.data, grad, z=Z, act=act, beta=beta,
                                          need_dx=X.requires_grad, need_db=bias is not None, owner=weight,
                                          x_staged=x_staged, dw_out=dw_out, db_out=db_out, grad_drop=grad_drop)
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


from ... import autograd as _autograd  # noqa: E402

_autograd._MASKED_GRAD_CONSUMERS.add(_linear_grad_fn)


class _SiblingGroup:
    """Horizontal fusion of Linear layers that read the SAME input (the wq / wk / wv projections of an attention block,
    examples/gpt.ipynb cell 2): their parameters are re-homed into one [sum N_i, K] weight and one [1, sum N_i] bias
    buffer (each layer's ``weight.data`` / ``bias.data`` becomes a view of its rows / columns, so the optimizer, state
    dicts and the un-fused path keep working), and one GEMM with N = sum N_i replaces the separate forward, dgrad and wgrad
    launches. Adoption happens on the first eager call (never inside a CUDA-graph capture) and is re-validated every
    call: if someone re-assigned a parameter's array the layers simply run un-fused again."""

    def __init__(self, layers):
        xp = layers[0].weight.xp
        self.layers = list(layers)
        self.K = layers[0].in_features
        self.Ns = [l.out_features for l in layers]
        self.offs = [0]
        for n in self.Ns:
            self.offs.append(self.offs[-1] + n)
        self.has_bias = layers[0].bias is not None
        self.W = xp.empty((self.offs[-1], self.K), dtype=np.float32)
        self.b = xp.empty((1, self.offs[-1]), dtype=np.float32) if self.has_bias else None
        for i, l in enumerate(self.layers):
            a, e = self.offs[i], self.offs[i + 1]
            self.W[a:e].copy_(l.weight.data)
            l.weight.data = self.W[a:e]
            if self.has_bias:
                self.b[:, a:e].copy_(l.bias.data.reshape(1, -1))
                l.bias.data = self.b[:, a:e]

    class _Holder:  # owner of the staged-weight cache entry of the fused matrix
        pass

    def valid(self):
        for i, l in enumerate(self.layers):
            a = self.offs[i]
            w = l.weight.data
            if w.data_ptr() != self.W.data_ptr() + a * self.K * 4 or tuple(w.shape) != (self.Ns[i], self.K):
                return False
            if self.has_bias and l.bias.data.data_ptr() != self.b.data_ptr() + a * 4:
                return False
        return True

    @staticmethod
    def eligible(layers, X):
        l0 = layers[0]
        return (len(layers) >= 2 and len({id(l) for l in layers}) == len(layers)
                and all(type(l) is type(l0) and l.in_features == l0.in_features and l._act == 0
                        and (l.bias is None) == (l0.bias is None) and l.device == "cuda" for l in layers)
                and X.ndim in (2, 3))

    @classmethod
    def of(cls, layers):
        from ... import b200
        key = tuple(id(l) for l in layers)
        g = getattr(layers[0], "_b200_sibling_group", None)
        if g is not None and g.key == key and g.valid():
            return g
        if b200._cache_scope()[0]:
            return None  # re-homing parameters allocates and copies: not inside a CUDA-graph capture
        g = cls(layers)
        g.key = key
        g.holder = cls._Holder()
        layers[0]._b200_sibling_group = g
        return g


fusion_stats = {"group_calls": 0, "backward_zero_copy": 0, "backward_gathered": 0}  # observability for tests / profiling


class _GroupCall:
    """One fused forward call: what the members' backward passes need to run as ONE fused backward."""
    __slots__ = ("group", "X", "xst", "grads", "expected", "arrived", "shape")

    def __init__(self, group, X, xst, shape):
        self.group, self.X, self.xst, self.shape = group, X, xst, shape
        self.grads = [None] * len(group.layers)
        self.expected, self.arrived = 0, 0


def _group_member_grad_fn(call, index):
    def grad_fn(*args, grad):
        # every member lists ALL weights / biases of the group in its args: a parameter's `_grad_ready` hook
        # (neunet.distributed) then fires only after the LAST member has run, i.e. after the fused backward below
        call.grads[index] = grad
        call.arrived += 1
        if call.arrived < call.expected:
            return
        _fused_sibling_backward(call)
    return grad_fn


def _fused_sibling_backward(call):
    from ... import b200
    torch = b200.torch
    g, X = call.group, call.X
    lead = tuple(call.shape[:-1])
    M = int(np.prod(lead))
    Nt = g.offs[-1]
    # zero-copy when the member gradients are already the column blocks of ONE [M, sum N] buffer (the fused attention
    # backward writes dq | dk | dv that way); otherwise they are gathered (a member without a gradient contributes zeros)
    base = None
    first = next((x for x in call.grads if x is not None), None)
    if first is not None and all(x is not None for x in call.grads):
        want_stride = tuple([Nt * int(np.prod(lead[i + 1:])) for i in range(len(lead))]) + (1,)
        p0 = call.grads[0].data_ptr()
        ok = all(tuple(x.shape) == lead + (g.Ns[i],) and tuple(x.stride()) == want_stride and x.dtype == torch.float32
                 and x.data_ptr() == p0 + g.offs[i] * 4 for i, x in enumerate(call.grads))
        if ok:
            base = torch.as_strided(call.grads[0], (M, Nt), (Nt, 1))
            fusion_stats["backward_zero_copy"] += 1
    if base is None:
        fusion_stats["backward_gathered"] += 1
        base = X.xp.zeros((M, Nt), dtype=np.float32)
        for i, x in enumerate(call.grads):
            if x is not None:
                base[:, g.offs[i]:g.offs[i + 1]].copy_(x.reshape(M, g.Ns[i]))
    dx, dW, db = b200.linear_backward(X.data, g.W, base, need_dx=X.requires_grad, need_db=g.has_bias, owner=g.holder,
                                      x_staged=call.xst)
    if dx is not None:
        X.apply_grad(dx)
    for i, l in enumerate(g.layers):
        a, e = g.offs[i], g.offs[i + 1]
        l.weight.apply_grad(dW[a:e])
        if g.has_bias:
            l.bias.apply_grad(db[:, a:e])
    call.grads = [None] * len(g.layers)
    call.arrived = 0


def _launch_sibling_group(X, sibs):
    """One GEMM for all pending Linear results on X. Returns False when the group form does not apply."""
    from ... import b200
    layers = [s_._f_layer for s_ in sibs]
    if not _SiblingGroup.eligible(layers, X):
        return False
    g = _SiblingGroup.of(layers)
    if g is None:
        return False
    training = layers[0].training_mode()
    O, _, xst = b200.linear_forward(X.data, g.W, g.b, owner=g.holder, keep_x_staged=training, x_owner=X)
    fusion_stats["group_calls"] += 1
    call = _GroupCall(g, X, xst, tuple(O.shape))
    params = tuple(l.weight for l in g.layers) + tuple(l.bias for l in g.layers if l.bias is not None)
    for i, s_ in enumerate(sibs):
        s_.args = (X,) + params
        s_.grad_fn = _group_member_grad_fn(call, i)
        s_.__dict__["_b200_gcall"] = call
        s_.data = O[..., g.offs[i]:g.offs[i + 1]]  # a column block of the fused output (strided view)
    return True


def _deferred_linear(layer, X: Tensor) -> Tensor:
    """``Linear.forward`` on "cuda" with fusion on: the result is *pending* (neunet/autograd.py, deferred
    evaluation). The GEMM is launched when the output is first read -- together with every other pending
    Linear on the same input: the q/k/v projections of one RMSNorm output become ONE GEMM over a shared weight
    buffer (``_SiblingGroup``) -- unless an ``nn.Swish`` consumes it first, in which case Swish becomes the GEMM's
    epilogue."""
    from ... import b200
    from ...autograd import _Deferred
    W, b = layer.weight, layer.bias
    act0, beta0, training = layer._act, layer._beta, layer.training_mode()

    def run(act, beta):
        return b200.linear_forward(X.data, W.data, b.data if b is not None else None, act=act, beta=beta,
                                   save_z=bool(act), owner=W, keep_x_staged=training, x_owner=X)

    def thunk():
        sibs = X.__dict__.get("_b200_pending_lin")
        if sibs:
            X.__dict__["_b200_pending_lin"] = None
            sibs = [s_ for s_ in sibs if s_._data is None]
            if len(sibs) >= 2 and _launch_sibling_group(X, sibs):
                return out._data
        O, Z, xst = run(act0, beta0)
        out.args = (X, W, b, Z, act0, beta0, xst)
        out.data = O
        if sibs:  # not fusable: at least launch the siblings right behind this one (same X planes, hot in L2)
            for sib in sibs:
                if sib is not out and sib._data is None:
                    sib.data  # noqa: B018 -- forces the sibling's GEMM
        return O

    def fuse_swish(beta):
        """``nn.Swish`` on this pending Linear: the result stays pending. Read as it is, Swish is the GEMM's epilogue;
        consumed by ``nn.Dropout`` first (the feed-forward block fc_2(dropout(swish(fc_1(x))))), the GEMM writes the
        pre-activation only and ONE pass turns it into the dropped activations + their bf16 planes for fc_2
        (``nnb_swish_dropout_fused``): the Swish output is never written unless somebody reads it."""
        lst = X.__dict__.get("_b200_pending_lin")
        if lst and out in lst:
            lst.remove(out)

        def deliver(Z, xst):
            out.args = (X, W, b, None, 0, 1.0, xst)
            out.data = Z  # the pre-activation is the GEMM's side output: the Linear result itself is delivered for free
            sw.args = (X, W, b, Z, 1, beta, xst)

        def launch():
            O, Z, xst = run(1, beta)
            deliver(Z, xst)
            return O

        def fuse_dropout(p, ticket):
            Z, _, xst = run(0, 1.0)
            deliver(Z, xst)
            sw.__dict__["_f_kind"] = "swish_lazy"
            sw.__dict__["_thunk"] = lambda: b200.swish_forward(Z, beta)
            return b200.swish_dropout_apply(Z, beta, p, ticket, want_planes=True)

        sw = _Deferred.make(launch, out.shape, (X, W, b, None, 1, beta, None), "linear_swish", True, _f_kind="linear_swish",
                            _f_fuse_dropout=fuse_dropout)
        sw.grad_fn = _linear_grad_fn
        return sw

    out = _Deferred.make(thunk, tuple(X.shape[:-1]) + (layer.out_features,), (X, W, b, None, act0, beta0, None), "linear", True,
                         _f_kind="linear", _f_act=act0, _f_fuse_swish=fuse_swish, _f_layer=layer)
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
