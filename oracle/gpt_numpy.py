"""CPU restatement (plain NumPy, fp32) of one training step of the reference's GPT example
(examples/gpt.ipynb cells 2-7, 11-12) with hand-written forward/backward built from
oracle/restated.py. TEST INFRASTRUCTURE / CPU BASELINE ONLY (see restated.py's header).

Pinned by tests/test_oracle_golden.py::test_gpt_step against tests/golden/model_gpt.npz, which was
produced by running the same architecture (examples/models.py) on the unmodified reference: loss,
logits, every parameter gradient and the post-Adam parameters.

Parameter order = the reference's Module.parameters() reflection order for that architecture:
embedding; per layer: self_attn(wq,bq,wk,bk,wv,bv,fc,bfc), cross_attn(8 tensors, never used),
ffn(fc_1 w,b, fc_2 w,b), norm1.w, norm2.w; then fc_out w,b. Dropout is the identity (eval / p=0).
"""
from __future__ import annotations

import math

import numpy as np

from . import restated as R

f32 = np.float32
PER_LAYER = 22  # tensors per DecoderLayer in parameters() order


def init_params(vocab, d_model, d_ff, n_layers, seed=0):
    """Same draws, same order as building the model under np.random.seed(seed)."""
    np.random.seed(seed)
    ps = [np.random.randn(vocab, d_model).astype(f32)]
    for _ in range(n_layers):
        for _attn in range(2):
            for _lin in range(4):
                w, b = R.linear_init(d_model, d_model)
                ps += [w, b]
        w, b = R.linear_init(d_model, d_ff); ps += [w, b]
        w, b = R.linear_init(d_ff, d_model); ps += [w, b]
        ps += [np.ones(d_model, f32), np.ones(d_model, f32)]
    w, b = R.linear_init(d_model, vocab)
    ps += [w, b]
    return ps


def positional_encoding(max_len, d_model):
    pe = np.zeros((max_len, d_model), f32)
    pos = np.arange(0, max_len, dtype=f32)[:, None]
    div = np.exp(np.arange(0, d_model, 2).astype(f32) * f32(-math.log(10000.0) / d_model)).astype(f32)
    pe[:, 0::2] = np.sin(pos * div)
    pe[:, 1::2] = np.cos(pos * div)
    return pe


def step_grads(ps, batch, n_heads, pad_idx=0, max_len=64):
    """Forward + backward of one batch. Returns (loss, logits2d, grads) with grads[i] = None for the
    never-used cross_attn tensors."""
    ids, tgt = batch[:, :-1], batch[:, 1:].reshape(-1).astype(np.int32)
    B, T = ids.shape
    d = ps[0].shape[1]
    n_layers = (len(ps) - 3) // PER_LAYER
    dep = d // n_heads
    scale = f32(math.sqrt(d))
    mask = ((ids != pad_idx).astype(int)[:, None, :] & np.logical_not(np.triu(np.ones((T, T)), k=1).astype(int)))[:, None]

    split = lambda t: t.reshape(B, T, n_heads, dep).transpose(0, 2, 1, 3)
    merge = lambda t: t.transpose(0, 2, 1, 3).reshape(B, T, d)

    x = (ps[0][ids] * scale + positional_encoding(max_len, d)[None, :T]).astype(f32)
    saved = []
    for l in range(n_layers):
        o = 1 + l * PER_LAYER
        wq, bq, wk, bk, wv, bv, wf, bf = ps[o:o + 8]
        w1, b1, w2, b2 = ps[o + 16:o + 20]
        g1, g2 = ps[o + 20], ps[o + 21]
        n1, xn1, s1 = R.rmsnorm_forward(x, g1)
        q, k, v = (split(R.linear_forward(n1, w, b)) for w, b in ((wq, bq), (wk, bk), (wv, bv)))
        sc = R.matmul_forward(q, k.transpose(0, 1, 3, 2)) / scale
        sc = np.where(mask == 0, f32(-1e9), sc).astype(f32)
        att = R.softmax_forward(sc, -1)
        ctx = merge(R.matmul_forward(att, v))
        x2 = x + R.linear_forward(ctx, wf, bf)
        n2, xn2, s2 = R.rmsnorm_forward(x2, g2)
        z = R.linear_forward(n2, w1, b1)
        h = R.swish_forward(z)
        x3 = x2 + R.linear_forward(h, w2, b2)
        saved.append((x, n1, xn1, s1, q, k, v, att, ctx, x2, n2, xn2, s2, z, h))
        x = x3.astype(f32)
    wo, bo = ps[-2], ps[-1]
    logits = R.linear_forward(x, wo, bo).reshape(B * T, -1)
    loss, dlog = R.cross_entropy(logits, tgt, ignore_index=pad_idx)

    grads = [None] * len(ps)
    dx, grads[-2], grads[-1] = R.linear_backward(x.reshape(B * T, d), wo, bo, dlog)
    dx = dx.reshape(B, T, d)
    for l in reversed(range(n_layers)):
        o = 1 + l * PER_LAYER
        wq, bq, wk, bk, wv, bv, wf, bf = ps[o:o + 8]
        w1, b1, w2, b2 = ps[o + 16:o + 20]
        g1, g2 = ps[o + 20], ps[o + 21]
        x0, n1, xn1, s1, q, k, v, att, ctx, x2, n2, xn2, s2, z, h = saved[l]
        # feed-forward branch
        dh, grads[o + 18], grads[o + 19] = R.linear_backward(h, w2, b2, dx)
        dz = R.swish_backward(z, dh)
        dn2, grads[o + 16], grads[o + 17] = R.linear_backward(n2, w1, b1, dz)
        dx2n, grads[o + 21], _ = R.rmsnorm_backward(x2, g2, None, xn2, s2, dn2)
        dx2 = dx + dx2n
        # attention branch
        dctx, grads[o + 6], grads[o + 7] = R.linear_backward(ctx, wf, bf, dx2)
        datt, dv = R.matmul_backward(att, v, split(dctx))
        dsc = R.softmax_backward(att, datt, -1)
        dsc = np.where(mask == 0, f32(0), dsc) / scale
        dq, dkt = R.matmul_backward(q, k.transpose(0, 1, 3, 2), dsc.astype(f32))
        dk = dkt.transpose(0, 1, 3, 2)
        dn1 = np.zeros_like(n1)
        for t, w, b, gi in ((dq, wq, bq, 0), (dk, wk, bk, 2), (dv, wv, bv, 4)):
            dd, grads[o + gi], grads[o + gi + 1] = R.linear_backward(n1, w, b, merge(t))
            dn1 = dn1 + dd
        dxn, grads[o + 20], _ = R.rmsnorm_backward(x0, g1, None, xn1, s1, dn1)
        dx = dx2 + dxn
    # Embedding backward through __getitem__: assignment, last write wins (autograd.py:909-910)
    ge = np.zeros_like(ps[0])
    ge[ids] = dx * scale
    grads[0] = ge
    return loss, logits, grads


def train_step(ps, ms, vs, t, batch, n_heads, lr=1.5e-4, betas=(0.9, 0.98), eps=1e-9, pad_idx=0):
    """In-place Adam step (examples/gpt.ipynb cell 11 hyper-parameters). Returns (loss, logits, grads)."""
    loss, logits, grads = step_grads(ps, batch, n_heads, pad_idx)
    for i, g in enumerate(grads):
        if g is None:
            continue  # optim.py:21-22
        ps[i], ms[i], vs[i] = R.adam_step(ps[i], g, ms[i], vs[i], t, lr=lr, betas=betas, eps=eps)
    return loss, logits, grads
