"""Fused kernels of round 2 on the B200, through the C-ABI: fused attention (forward + backward), dropout with
residual / operand planes, RMSNorm with the residual-dropout prologue and operand planes, RMSNorm backward with
gradient accumulation -- each against a plain torch fp32 composition of the reference's separate ops
(examples/gpt.ipynb cell 2 l.25-40; neunet/nn/layers/dropout.py:17-46; rmsnorm.py:39-94) -- and, at model level, the
fused (deferred) execution of the GPT example against the strictly eager one with the SAME dropout masks."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _b200():
    from neunet import b200
    b200.require_device()
    return b200


def _bf16_planes(buf, rows, cols, x3):
    hi = buf[: rows * cols * 2].view(torch.bfloat16).reshape(rows, cols)
    if not x3:
        return hi.float()
    off = ((rows * cols * 2 + 255) // 256) * 256
    lo = buf[off: off + rows * cols * 2].view(torch.bfloat16).reshape(rows, cols)
    return hi.float() + lo.float()


def _ref_attention(q, kT, v, masked, fill, scale, keep):
    s = torch.matmul(q, kT) / scale
    if masked is not None:
        s = torch.where(masked, torch.full((), fill, device=s.device), s)
    p = torch.softmax(s, dim=-1)
    a = p * keep
    return a, torch.matmul(a, v)


@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
@pytest.mark.parametrize("B,H,Tq,Tk,D,mask_kind,p", [
    (2, 3, 64, 64, 64, 2, 0.1),     # GPT shape, int mask tested against 0, dropout
    (1, 2, 7, 12, 16, 1, 0.0),      # ragged tile, float condition tensor, no dropout
    (3, 1, 33, 64, 32, 3, 0.25),    # float mask == value
    (2, 2, 64, 64, 64, 0, 0.0),     # no mask at all
    (2, 2, 61, 60, 20, 2, 0.0),     # nothing a multiple of 8: the MMA k-loops run into the zero padding
])
def test_attention_forward_backward_vs_torch(B, H, Tq, Tk, D, mask_kind, p, prec):
    """The per-head products are TF32 tensor-core MMAs: bf16x3 = hi/lo split, three products (fp32-grade, the parity
    bar); bf16 = one TF32 product (10-bit mantissa; the bar is the library's bf16-mode tolerance)."""
    b200 = _b200()
    torch.backends.cuda.matmul.allow_tf32 = False
    with b200.precision(prec):
        _attention_case(b200, B, H, Tq, Tk, D, mask_kind, p, tight=prec == "bf16x3")


def _attention_case(b200, B, H, Tq, Tk, D, mask_kind, p, tight):
    tol_a, tol_o, tol_g = (1e-5, 1e-5, 2e-5) if tight else (2e-3, 4e-3, 6e-3)
    g = torch.Generator(device="cuda").manual_seed(B * 100 + Tq)
    # operands in the example's memory order: (B, T, H, D) buffers seen through transposed views
    q = torch.randn(B, Tq, H, D, generator=g, device="cuda").permute(0, 2, 1, 3)
    k = torch.randn(B, Tk, H, D, generator=g, device="cuda").permute(0, 2, 1, 3)
    v = torch.randn(B, Tk, H, D, generator=g, device="cuda").permute(0, 2, 1, 3)
    kT = k.permute(0, 1, 3, 2)
    dO = torch.randn(B, Tq, H, D, generator=g, device="cuda").permute(0, 2, 1, 3)
    causal = torch.tril(torch.ones(Tq, Tk, device="cuda"))
    mask, masked = None, None
    if mask_kind == 2:
        base = causal.to(torch.int32).expand(B, 1, Tq, Tk).contiguous()
        mask, masked = (base, 2, 0.0), (base == 0).expand(B, H, Tq, Tk)
    elif mask_kind == 1:
        cond = (causal == 0).float().expand(B, 1, Tq, Tk).contiguous()
        mask, masked = (cond, 1, 0.0), (cond != 0).expand(B, H, Tq, Tk)
    elif mask_kind == 3:
        base = (causal * 5.0).expand(B, 1, Tq, Tk).contiguous()
        mask, masked = (base, 3, 0.0), (base == 0).expand(B, H, Tq, Tk)
    ticket = (1234, 7, 99, None)
    keep = torch.ones(B, H, Tq, Tk, device="cuda")
    if p > 0:
        keep = b200.dropout_apply(keep, p, ticket)  # the same Philox ticket regenerates the same mask
    scale, fill = float(np.sqrt(H * D)), -1e9
    out, attn, planes = b200.attention_forward(q, kT, v, mask, fill, scale, p, ticket if p > 0 else None, want_planes=True)
    a_ref, o_ref = _ref_attention(q, kT, v, masked, fill, scale, keep)
    assert out.shape == (B, H, Tq, D) and out.permute(0, 2, 1, 3).is_contiguous()
    assert (attn - a_ref).abs().max().item() <= tol_a
    assert (out - o_ref).abs().max().item() <= tol_o * max(1.0, o_ref.abs().max().item())
    got = _bf16_planes(planes[1], B * Tq, H * D, b200.get_precision() == "bf16x3")
    want = out.permute(0, 2, 1, 3).reshape(B * Tq, H * D)
    assert (got - want).abs().max().item() <= (1e-4 if b200.get_precision() == "bf16x3" else 2e-2) * max(1.0, want.abs().max().item())
    # backward against torch autograd on the same composition
    ql, kl, vl = (t.detach().clone().requires_grad_(True) for t in (q, kT, v))
    _, o2 = _ref_attention(ql, kl, vl, masked, fill, scale, keep)
    o2.backward(dO)
    dq, dkT, dv = b200.attention_backward(q, kT, v, mask, fill, scale, p, ticket if p > 0 else None, dO)
    for got, want in ((dq, ql.grad), (dkT, kl.grad), (dv, vl.grad)):
        assert got.shape == want.shape
        assert (got - want).abs().max().item() <= tol_g * max(1.0, want.abs().max().item())
    if Tq == Tk:  # dq | dk | dv are the column blocks of one [B*T, 3*H*D] matrix (the upstream gradient of a fused q/k/v GEMM)
        want = (Tq * 3 * H * D, 3 * H * D, D, 1)
        assert dq.permute(0, 2, 1, 3).stride() == want and dkT.permute(0, 3, 1, 2).stride() == want
        assert dkT.data_ptr() - dq.data_ptr() == H * D * 4 and dv.data_ptr() - dq.data_ptr() == 2 * H * D * 4
    else:
        assert dq.permute(0, 2, 1, 3).is_contiguous() and dkT.permute(0, 3, 1, 2).is_contiguous()


def test_attention_rejects_unsupported_sizes():
    b200 = _b200()
    assert b200.attention_supported(64, 64, 64) and not b200.attention_supported(65, 64, 64)
    assert not b200.attention_supported(8, 10, 16)  # Tk % 4 != 0
    q = torch.randn(1, 1, 80, 16, device="cuda")
    with pytest.raises(RuntimeError):
        b200.attention_forward(q, q.permute(0, 1, 3, 2), q, None, 0.0, 1.0, 0.0, None)


@pytest.mark.parametrize("prec", ["bf16", "bf16x3"])
def test_dropout_residual_and_planes(prec):
    b200 = _b200()
    with b200.precision(prec):
        x = torch.randn(6, 40, 64, device="cuda")
        r = torch.randn(6, 40, 64, device="cuda")
        ticket = (5, 2, 17, None)
        plain = b200.dropout_apply(x, 0.3, ticket)
        y, planes = b200.dropout_apply(x, 0.3, ticket, residual=r, want_planes=True)
        assert torch.equal(y, plain + r)
        frac = (plain == 0).float().mean().item()
        assert 0.25 < frac < 0.35
        got = _bf16_planes(planes[1], 240, 64, prec == "bf16x3")
        assert (got - y.reshape(240, 64)).abs().max().item() <= (1e-4 if prec == "bf16x3" else 4e-2)
        assert planes[0] == b200.planes_key(y.reshape(240, 64))


@pytest.mark.parametrize("cols,with_bias", [(512, False), (96, True), (1024, False)])
def test_rmsnorm_fused_forward_backward(cols, with_bias):
    b200 = _b200()
    rows = 300
    x = torch.randn(rows, cols, device="cuda")
    a = torch.randn(rows, cols, device="cuda")
    w = torch.rand(cols, device="cuda") + 0.5
    b = torch.randn(cols, device="cuda") if with_bias else None
    ticket = (77, 1, 3, None)
    s_ref = b200.dropout_apply(a, 0.1, ticket, residual=x)
    y0, std0, _, _ = b200.rmsnorm_forward(s_ref, w, b, 1e-6)
    y, std, s, planes = b200.rmsnorm_forward(x, w, b, 1e-6, add_dropout=(a, 0.1, ticket), want_planes=True)
    assert torch.equal(s, s_ref)
    ref = s_ref / torch.sqrt((s_ref * s_ref).mean(-1, keepdim=True) + 1e-6) * w + (b if b is not None else 0)
    assert (y - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
    assert (y - y0).abs().max().item() <= 1e-6 * max(1.0, ref.abs().max().item())
    assert (std - std0).abs().max().item() <= 1e-6
    got = _bf16_planes(planes[1], rows, cols, b200.get_precision() == "bf16x3")
    assert (got - y).abs().max().item() <= (1e-4 if b200.get_precision() == "bf16x3" else 4e-2) * max(1.0, y.abs().max().item())
    g = torch.randn(rows, cols, device="cuda")
    prev = torch.randn(rows, cols, device="cuda")
    dx0, dw0, db0, acc0 = b200.rmsnorm_backward(g, s, w, std, need_db=with_bias)
    dx1, dw1, db1, acc1 = b200.rmsnorm_backward(g, s, w, std, need_db=with_bias, dx_add=prev)
    assert not acc0 and acc1
    assert (dx1 - (dx0 + prev)).abs().max().item() <= 1e-6 * max(1.0, dx0.abs().max().item())
    assert torch.equal(dw0, dw1)


def _gpt_step(fuse, dropout, prec="bf16x3", sizes=None):
    import models as M
    import neunet
    import neunet.nn as nn
    from neunet import autograd, b200
    sizes = sizes or dict(vocab=120, d_model=64, n_heads=4, d_ff=128, n_layers=2)
    np.random.seed(5)
    model = M.build_gpt(neunet, nn, pad_idx=0, device="cuda", dropout=dropout, **sizes)
    model.train()
    rng = np.random.RandomState(1)
    batch = rng.randint(1, sizes["vocab"], (4, 17))
    batch[1, -3:] = 0
    prev = autograd.set_fusion(fuse)
    try:
        b200.manual_seed(99)
        b200.reset_launch_count()
        with b200.precision(prec):
            loss_fn = nn.CrossEntropyLoss(ignore_index=0)
            out, attn = model.forward(batch[:, :-1])
            logits = out.reshape(out.shape[0] * out.shape[1], out.shape[2])
            loss = loss_fn(logits, neunet.tensor(batch[:, 1:].flatten(), device="cuda", dtype=neunet.int32))
            loss.backward()
        torch.cuda.synchronize()
        return (float(loss.data.item()), out.data.clone(), attn.data.clone(),
                [None if p.grad is None else p.grad.clone() for p in model.parameters()], b200.launch_count())
    finally:
        autograd.set_fusion(prev)


@pytest.mark.parametrize("dropout", [0.0, 0.1])
def test_gpt_fused_execution_equals_eager_with_same_masks(dropout):
    """Same seeds -> same Philox tickets in the same order -> identical masks: the deferred/fused step must reproduce
    the eager step to contraction round-off (bf16x3; max-norm relative error)."""
    eager = _gpt_step(False, dropout)
    fused = _gpt_step(True, dropout)
    assert abs(eager[0] - fused[0]) <= 1e-4 * abs(eager[0])
    for got, want in ((fused[1], eager[1]), (fused[2], eager[2])):
        assert (got - want).abs().max().item() <= 1e-4 * max(want.abs().max().item(), 1e-3)
    for got, want in zip(fused[3], eager[3]):
        assert (got is None) == (want is None)
        if got is not None:
            assert (got - want).abs().max().item() <= 2e-4 * max(want.abs().max().item(), 1e-3)
    assert fused[4] < eager[4]  # fewer native launches, and none of the array back-end's where/div/copy kernels in between


def test_fused_step_captures_into_a_cuda_graph():
    import models as M
    import neunet
    import neunet.nn as nn
    from neunet import b200, optim
    np.random.seed(2)
    model = M.build_gpt(neunet, nn, vocab=64, d_model=64, n_heads=4, d_ff=128, n_layers=1, pad_idx=0, device="cuda", dropout=0.1)
    model.train()
    opt = optim.Adam(model.parameters(), lr=1e-3)
    loss_fn = nn.CrossEntropyLoss(ignore_index=0)
    T = 16
    ids = neunet.tensor(np.random.randint(1, 64, (2, T)), dtype=np.int32, device="cuda")
    tgt = neunet.tensor(np.random.randint(1, 64, (2 * T,)), dtype=np.int32, device="cuda")
    mask = neunet.tensor(np.broadcast_to(np.tril(np.ones((T, T))), (2, T, T)).copy(), dtype=np.int32, device="cuda")

    def step(i, t):
        opt.zero_grad()
        out, _ = model.decoder(i, mask)
        loss = loss_fn(out.reshape(2 * T, 64), t)
        loss.backward()
        opt.step()
        return loss
    with b200.precision("bf16"):
        for _ in range(2):
            step(ids, tgt)
        g = b200.GraphedStep(step, [ids, tgt], optimizer=opt, warmup=1)
        losses = [float(g.replay().item()) for _ in range(5)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


@pytest.mark.parametrize("B,C,H,W,alpha,affine", [(8, 128, 16, 16, 0.01, True), (5, 16, 7, 7, 1.0, True), (64, 256, 4, 4, 0.01, False),
                                                  (64, 32, 32, 32, 0.01, True), (3, 520, 2, 2, 0.2, True), (40, 64, 8, 8, 0.01, True)])
def test_leaky_relu_batchnorm_fused_vs_torch(B, C, H, W, alpha, affine):
    """Fused LeakyReLU + BatchNorm2d kernels vs torch fp32 (F.leaky_relu + F.batch_norm in training mode, the semantics of
    neunet/nn/layers/batchnorm2d.py:57-115 with ddof = 0 batch statistics), forward, running stats, dx / dw / db."""
    import torch.nn.functional as F
    b200 = _b200()
    g = torch.Generator(device="cuda").manual_seed(C)
    x = torch.randn(B, C, H, W, generator=g, device="cuda") * 1.5 + 0.3
    w = (torch.rand(C, generator=g, device="cuda") + 0.5) if affine else None
    b = torch.randn(C, generator=g, device="cuda") if affine else None
    rm, rv = torch.zeros(1, C, device="cuda"), torch.ones(1, C, device="cuda")
    y, mean, inv = b200.bn_forward(x, w, b, alpha, 1e-5, 0.1, rm, rv)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True) if affine else None
    br = b.clone().requires_grad_(True) if affine else None
    a = F.leaky_relu(xr, alpha) if alpha != 1.0 else xr
    ref = F.batch_norm(a, None, None, wr, br, training=True, eps=1e-5)
    assert (y - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    am = a.detach().mean((0, 2, 3))
    av = a.detach().var((0, 2, 3), unbiased=False)
    assert (mean - am).abs().max().item() <= 1e-5
    # the reference's momentum convention: running = momentum * running + (1 - momentum) * batch statistic
    assert (rm.reshape(-1) - 0.9 * am).abs().max().item() <= 1e-5
    assert (rv.reshape(-1) - (0.1 + 0.9 * av)).abs().max().item() <= 1e-4
    go = torch.randn(B, C, H, W, generator=g, device="cuda")
    ref.backward(go)
    dx, dw, db = b200.bn_backward(x, go, mean, inv, w, alpha, need_dx=True, need_dw=affine)
    assert (dx - xr.grad).abs().max().item() <= 5e-5 * max(1.0, xr.grad.abs().max().item())
    if affine:
        assert (dw - wr.grad).abs().max().item() <= 5e-5 * max(1.0, wr.grad.abs().max().item())
        assert (db - br.grad).abs().max().item() <= 5e-5 * max(1.0, br.grad.abs().max().item())


def test_conv_lrelu_bn_block_fused_equals_eager():
    """conv -> LeakyReLU -> BatchNorm2d (DDPM ResBlock head) through the public API: deferred/fused vs eager."""
    import neunet
    import neunet.nn as nn
    from neunet import autograd, b200
    res = []
    for fuse in (False, True):
        prev = autograd.set_fusion(fuse)
        try:
            np.random.seed(4)
            conv, act, bn = nn.Conv2d(8, 32, 3, 1, 1).to("cuda"), nn.LeakyReLU(0.01), nn.BatchNorm2d(32).to("cuda")
            x = neunet.tensor(np.random.randn(6, 8, 16, 16), device="cuda", requires_grad=True)
            with b200.precision("bf16x3"):
                y = bn(act(conv(x)))
                (y * y).mean().backward()
            res.append([y.data.clone(), x.grad.clone(), conv.weight.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone(),
                        bn.running_var.data.clone()])
        finally:
            autograd.set_fusion(prev)
    for a, b in zip(*res):
        assert (a - b).abs().max().item() <= 1e-4 * max(b.abs().max().item(), 1e-3)


@pytest.mark.parametrize("dtype", [torch.int32, torch.int64])
def test_embedding_gather_and_last_write_wins_scatter(dtype):
    """nn.Embedding kernels against NumPy's own semantics of `w[ids]` / `full[ids] = grad` (the reference's backward,
    neunet/autograd.py:909-910): with duplicate ids the LAST occurrence wins -- bit-exact (index work)."""
    b200 = _b200()
    rng = np.random.RandomState(0)
    V, D = 57, 36
    w = rng.randn(V, D).astype(np.float32)
    ids = rng.randint(0, V, (5, 23))
    ids[0, :6] = [3, 3, 3, 9, 9, -1]      # duplicates and a negative (wrapping) index
    g = rng.randn(5, 23, D).astype(np.float32)
    out = b200.embedding_forward(torch.from_numpy(w).cuda(), torch.from_numpy(ids).to(dtype).cuda())
    np.testing.assert_array_equal(out.cpu().numpy(), w[ids])
    full = np.zeros_like(w)
    full[ids] = g
    dw = b200.embedding_backward(torch.from_numpy(ids).to(dtype).cuda(), torch.from_numpy(g).cuda(), V)
    np.testing.assert_array_equal(dw.cpu().numpy(), full)


@pytest.mark.parametrize("prec,tol", [("bf16x3", 1e-4), ("bf16", 6e-3)])
@pytest.mark.parametrize("rows,C,K", [(256, 1500, 64), (130, 1001, 40), (4096, 15000, 512)])
def test_lm_head_cross_entropy_staged_backward(prec, tol, rows, C, K):
    """Row N3: CrossEntropy(Linear(x)) backward with dlogits emitted as bf16 operand planes (never as fp32) against the
    composition of the two stand-alone backward passes AND against torch fp32 autograd (F.cross_entropy(F.linear))."""
    b200 = _b200()
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(rows + C)
    x = torch.randn(rows, K, generator=g, device="cuda")
    w = torch.randn(C, K, generator=g, device="cuda") / K ** 0.5
    b = torch.randn(1, C, generator=g, device="cuda") * 0.1
    tgt = torch.randint(0, C, (rows,), generator=g, device="cuda", dtype=torch.int32)
    tgt[::7] = 0  # ignore_index rows
    with b200.precision(prec):
        logits, _, xst = b200.linear_forward(x, w, b, keep_x_staged=True)
        loss, saved = b200.cross_entropy_forward(logits, tgt, ignore_index=0, reduction="mean")
        up = torch.ones((), device="cuda") * 1.7
        dx, dw, db = b200.cross_entropy_linear_backward(saved, up, x, w, x_staged=xst)
        d = b200.cross_entropy_backward(saved, up)
        dx0, dw0, db0 = b200.linear_backward(x, w, d, x_staged=xst)
    for got, want in ((dx, dx0), (dw, dw0), (db, db0)):
        # same math, but the staged form rounds dlogits to bf16(x3) BEFORE the column sums / contractions
        assert (got - want).abs().max().item() <= tol * max(want.abs().max().item(), 1e-6)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    ref = torch.nn.functional.cross_entropy(torch.nn.functional.linear(xr, wr, br.reshape(-1)), tgt.long(), ignore_index=0) * 1.7
    ref.backward()
    assert abs(float(loss.item()) * 1.7 - float(ref.item())) <= tol * abs(float(ref.item()))
    for got, want in ((dx, xr.grad), (dw, wr.grad), (db, br.grad)):
        assert (got - want).abs().max().item() <= tol * max(want.abs().max().item(), 1e-6)


@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
def test_swish_dropout_one_pass_equals_the_two_ops(prec):
    """nnb_swish_dropout_fused (fc_2(dropout(swish(fc_1 x))) of examples/gpt.ipynb) against the stand-alone Swish followed by
    the stand-alone dropout with the same Philox ticket: same mask, same values, and planes = bf16 split of the result."""
    b200 = _b200()
    g = torch.Generator(device="cuda").manual_seed(3)
    z = torch.randn(6, 50, 72, generator=g, device="cuda") * 3
    ticket = (77, 5, 1234, None)
    with b200.precision(prec):
        y, planes = b200.swish_dropout_apply(z, 1.3, 0.2, ticket, want_planes=True)
        want = b200.dropout_apply(b200.swish_forward(z, 1.3), 0.2, ticket)
        assert (y - want).abs().max().item() <= 2e-6 * want.abs().max().item() and torch.equal(y == 0, want == 0)
        ref = torch.where(want != 0, (z * torch.sigmoid(1.3 * z)) / 0.8, torch.zeros_like(z))
        assert (y - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
        assert abs((y == 0).float().mean().item() - 0.2) < 0.03
        got = _bf16_planes(planes[1], 300, 72, prec == "bf16x3")
        assert (got - y.reshape(300, 72)).abs().max().item() <= (1e-4 if prec == "bf16x3" else 2e-2) * y.abs().max().item()


def test_ffn_block_fused_equals_unfused_and_never_writes_the_swish_output():
    """Linear -> Swish -> Dropout -> Linear through the public API with fusion on and off: same values and gradients;
    with fusion on the Swish output stays pending (it is only produced if somebody reads it)."""
    import neunet
    import neunet.nn as nn
    from neunet import autograd
    b200 = _b200()
    np.random.seed(4)
    fc_1, act, drop, fc_2 = nn.Linear(64, 256).to("cuda"), nn.Swish(), nn.Dropout(0.1), nn.Linear(256, 64).to("cuda")
    xv = np.random.randn(8, 16, 64).astype(np.float32)
    res = {}
    for fuse in (True, False):
        prev = autograd.set_fusion(fuse)
        try:
            b200.manual_seed(9)
            for layer in (fc_1, fc_2):
                layer.weight.grad = None
                layer.bias.grad = None
            x = neunet.tensor(xv, device="cuda", requires_grad=True)
            s = act(fc_1(x))
            out = fc_2(drop(s))
            ov = out.data.clone()
            if fuse:
                assert s.pending  # fc_1 GEMM -> swish_dropout pass -> fc_2 GEMM; no kernel wrote swish(fc_1 x)
            (out * out).sum().backward()
            res[fuse] = [ov, x.grad.clone(), fc_1.weight.grad.clone(), fc_1.bias.grad.clone(), fc_2.weight.grad.clone(), s.data.clone()]
        finally:
            autograd.set_fusion(prev)
    for a, b in zip(res[True], res[False]):
        assert (a - b).abs().max().item() <= 2e-5 * max(1.0, b.abs().max().item())


@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
def test_linear_backward_with_the_dropout_mask_folded_into_staging(prec, act):
    """nnb_linear_backward_dropped == nnb_dropout on the upstream gradient followed by nnb_linear_backward: same planes,
    same GEMMs, so the results are identical."""
    b200 = _b200()
    g = torch.Generator(device="cuda").manual_seed(12)
    M, K, N = 3 * 70, 96, 136
    x = torch.randn(3, 70, K, generator=g, device="cuda")
    w = torch.randn(N, K, generator=g, device="cuda") * 0.1
    z = torch.randn(3, 70, N, generator=g, device="cuda")
    grad = torch.randn(3, 70, N, generator=g, device="cuda")
    ticket = (5, 9, 321, None)
    with b200.precision(prec):
        got = b200.linear_backward(x, w, grad, z=z if act else None, act=act, beta=1.2, grad_drop=(0.3, ticket))
        want = b200.linear_backward(x, w, b200.dropout_apply(grad, 0.3, ticket), z=z if act else None, act=act, beta=1.2)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_optimizer_issued_chunk_by_chunk_equals_one_launch():
    """GradBucket.all_reduce_and_step (one AdamW launch per gradient chunk, nnb_adamw_step_range) against all_reduce() +
    step() (one launch over everything) on one GPU: bit-identical parameters and moments after three steps, with a
    parameter that never receives a gradient in the middle of the list."""
    import neunet
    import neunet.nn as nn
    from neunet.distributed import GradBucket
    from neunet.optim import AdamW
    _b200()

    def run(chunked):
        np.random.seed(21)
        layers = [nn.Linear(48, 96).to("cuda"), nn.Linear(96, 96).to("cuda"), nn.Linear(96, 40).to("cuda")]
        dead = nn.Linear(8, 8).to("cuda")  # never used: grad stays None (optim.py:21-22 skips it)
        params = layers[0].parameters() + dead.parameters() + layers[1].parameters() + layers[2].parameters()
        opt = AdamW(params, lr=1e-2, weight_decay=0.1)
        bucket = GradBucket(params, chunk_bytes=8 << 10)
        rng = np.random.RandomState(2)
        for t in range(4):
            opt.zero_grad()
            h = neunet.tensor(rng.randn(32, 48).astype(np.float32), device="cuda")
            for l in layers:
                h = l(h)
            (h * h).sum().backward()
            if chunked and t >= 1:
                bucket.all_reduce_and_step(opt)
            else:
                bucket.all_reduce()
                opt.step()
            if t == 0:
                bucket.overlap_backward()
                assert len(bucket._chunks) >= 3
        return [p.data.clone() for p in params] + [m.clone() for m in opt.m] + [v.clone() for v in opt.v]

    for a, b in zip(run(True), run(False)):
        assert torch.equal(a, b)
