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
    torch.cuda.synchronize()
    print("sanitize_small ok")


if __name__ == "__main__":
    main()
