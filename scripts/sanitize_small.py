#!/usr/bin/env python
"""Small-shape pass over every round-2 kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py

Shapes are tiny so the 10-100x slowdown of the tools stays within a minute; values are checked against torch so a silent
out-of-bounds read that corrupts results also fails."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "numpy-nn-model_b200"), ROOT, os.path.join(ROOT, "examples")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import neunet  # noqa: E402
import neunet.nn as nn  # noqa: E402
from neunet import b200  # noqa: E402


def main():
    torch.cuda.set_device(0)
    b200.require_device()
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False  # the torch checks below must be fp32, not TF32
    torch.backends.cuda.matmul.allow_tf32 = False
    with b200.precision("bf16x3"):
        # fused attention (ragged + full tile), dropout, mask kinds
        for (B, H, Tq, Tk, D, p) in ((2, 2, 64, 64, 64, 0.1), (1, 3, 7, 12, 16, 0.0)):
            q = torch.randn(B, Tq, H, D, device="cuda").permute(0, 2, 1, 3)
            k = torch.randn(B, Tk, H, D, device="cuda").permute(0, 2, 1, 3)
            v = torch.randn(B, Tk, H, D, device="cuda").permute(0, 2, 1, 3)
            m = torch.tril(torch.ones(Tq, Tk, device="cuda")).to(torch.int32).expand(B, 1, Tq, Tk).contiguous()
            tk = (3, 1, 5, None) if p > 0 else None
            out, attn, planes = b200.attention_forward(q, k.permute(0, 1, 3, 2), v, (m, 2, 0.0), -1e9, 8.0, p, tk, want_planes=True)
            b200.attention_backward(q, k.permute(0, 1, 3, 2), v, (m, 2, 0.0), -1e9, 8.0, p, tk, torch.randn_like(out))
        # linear (+ swish), rmsnorm fused, dropout fused, CE, embedding, adam through a tiny GPT step (fusion on)
        import models as M
        np.random.seed(0)
        model = M.build_gpt(neunet, nn, vocab=64, d_model=64, n_heads=4, d_ff=128, n_layers=1, pad_idx=0, device="cuda", dropout=0.1)
        model.train()
        from neunet.optim import Adam
        opt = Adam(model.parameters(), lr=1e-3)
        batch = np.random.randint(1, 64, (2, 17))
        for _ in range(2):
            opt.zero_grad()
            loss, _ = M.gpt_train_step(neunet, nn, model, opt, batch)
        assert np.isfinite(float(loss.item()))
        # conv: implicit GEMM forward / dgrad / wgrad, strided dgrad classes, native ConvTranspose2d, fused LeakyReLU + BatchNorm
        x = torch.randn(2, 64, 8, 8, device="cuda")
        for (k_, s_, p_) in ((3, 1, 1), (4, 2, 1)):
            w = torch.randn(64, 64, k_, k_, device="cuda") * 0.05
            o, pl = b200.conv2d_forward(x, w, None, (s_, s_), (p_,) * 4, (1, 1), keep_planes=True)
            ref = torch.nn.functional.conv2d(x, w, None, stride=s_, padding=p_)
            assert (o - ref).abs().max().item() < 1e-3
            b200.conv2d_backward(x, w, torch.randn_like(o), (s_, s_), (p_,) * 4, (1, 1), x_planes=pl)
        w = torch.randn(64, 64, 4, 4, device="cuda") * 0.05
        o, pl = b200.conv_transpose2d_forward(x, w, None, (2, 2), (1, 1, 1, 1), (1, 1), (0, 0))
        b200.conv_transpose2d_backward(x, w, torch.randn_like(o), (2, 2), (1, 1, 1, 1), (1, 1), (0, 0), x_planes=pl)
        x3 = torch.randn(2, 3, 8, 8, device="cuda")  # 3-channel layer: zero-padded channel pitch
        w3 = torch.randn(64, 3, 3, 3, device="cuda")
        o3 = b200.conv2d_forward(x3, w3, None, (1, 1), (1, 1, 1, 1), (1, 1))
        assert (o3 - torch.nn.functional.conv2d(x3, w3, None, padding=1)).abs().max().item() < 1e-3
        b200.conv2d_backward(x3, w3, torch.randn_like(o3), (1, 1), (1, 1, 1, 1), (1, 1))
        rm, rv = torch.zeros(1, 64, device="cuda"), torch.ones(1, 64, device="cuda")
        y, mean, inv = b200.bn_forward(x, torch.ones(64, device="cuda"), torch.zeros(64, device="cuda"), 0.01, 1e-5, 0.1, rm, rv)
        b200.bn_backward(x, torch.randn_like(x), mean, inv, torch.ones(64, device="cuda"), 0.01)
        for shp in ((3, 16, 2, 2), (4, 8, 32, 32), (5, 6, 3, 3)):  # planes smaller / larger than a block, odd plane size
            xb = torch.randn(*shp, device="cuda")
            C = shp[1]
            y, mean, inv = b200.bn_forward(xb, torch.ones(C, device="cuda"), torch.zeros(C, device="cuda"), 0.01, 1e-5, 0.1,
                                           torch.zeros(1, C, device="cuda"), torch.ones(1, C, device="cuda"))
            ref = torch.nn.functional.batch_norm(torch.nn.functional.leaky_relu(xb, 0.01), None, None, training=True, eps=1e-5)
            assert (y - ref).abs().max().item() < 1e-4
            b200.bn_backward(xb, torch.randn_like(xb), mean, inv, torch.ones(C, device="cuda"), 0.01)
    # bf16 mode: the one-product instantiations of the attention kernels; swish + dropout pass; masked dO staging
    with b200.precision("bf16"):
        q = torch.randn(2, 33, 2, 32, device="cuda").permute(0, 2, 1, 3)
        kv = torch.randn(2, 40, 2, 32, device="cuda").permute(0, 2, 1, 3)
        cond = (torch.rand(2, 1, 33, 40, device="cuda") < 0.3).float()
        out, attn, _ = b200.attention_forward(q, kv.permute(0, 1, 3, 2), kv, (cond, 1, 0.0), -1e9, 5.0, 0.2, (1, 2, 3, None))
        b200.attention_backward(q, kv.permute(0, 1, 3, 2), kv, (cond, 1, 0.0), -1e9, 5.0, 0.2, (1, 2, 3, None), torch.randn_like(out))
        z = torch.randn(3, 10, 24, device="cuda")
        y, planes = b200.swish_dropout_apply(z, 1.0, 0.25, (4, 5, 6, None))
        want = b200.dropout_apply(b200.swish_forward(z, 1.0), 0.25, (4, 5, 6, None))
        assert (y - want).abs().max().item() < 1e-5
        xl, wl, gl = torch.randn(30, 40, device="cuda"), torch.randn(24, 40, device="cuda"), torch.randn(30, 24, device="cuda")
        b200.linear_backward(xl, wl, gl, z=z.reshape(30, 24), act=1, beta=1.0, grad_drop=(0.25, (4, 5, 6, None)))
    # optimizer issued as ranges
    from neunet.distributed import GradBucket
    from neunet.optim import AdamW
    lins = [nn.Linear(16, 24).to("cuda"), nn.Linear(24, 8).to("cuda")]
    params = lins[0].parameters() + lins[1].parameters()
    opt2, bucket = AdamW(params, lr=1e-2), GradBucket(params, chunk_bytes=1 << 10)
    for t in range(3):
        opt2.zero_grad()
        h = neunet.tensor(np.random.randn(4, 16).astype(np.float32), device="cuda")
        for l in lins:
            h = l(h)
        (h * h).sum().backward()
        bucket.all_reduce_and_step(opt2)
        if t == 0:
            bucket.overlap_backward()
    torch.cuda.synchronize()
    print("sanitize_small ok")


if __name__ == "__main__":
    main()
