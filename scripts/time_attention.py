#!/usr/bin/env python
"""Time nnb_attention_forward / backward at the GPT-small shape (B=64, H=8, T=64, D=64: 512 heads per launch) with CUDA
events over rotating operand sets larger than L2, in both precision modes (host-paced: the number includes the
allocations of the Python binding; the in-step figure is the ncu launch list under profiles/)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "numpy-nn-model_b200"))
from neunet import b200  # noqa: E402

b200.require_device()
B, H, Tn, D = 64, 8, 64, 64
SETS = 24  # 24 x (q, k, v, dO) x 8 MB = 768 MB > 126 MB L2
g = torch.Generator(device="cuda").manual_seed(0)
qkv = [torch.randn(B, Tn, 3, H, D, generator=g, device="cuda") for _ in range(SETS)]
dO = [torch.randn(B, Tn, H, D, generator=g, device="cuda").permute(0, 2, 1, 3) for _ in range(SETS)]
mask = (torch.tril(torch.ones(Tn, Tn, device="cuda")).to(torch.int32).expand(B, 1, Tn, Tn).contiguous(), 2, 0.0)
scale = float((H * D) ** 0.5)
ticket = (1, 2, 3, None)


def views(i):
    t = qkv[i]
    q, k, v = (t[:, :, j].permute(0, 2, 1, 3) for j in range(3))
    return q, k.permute(0, 1, 3, 2), v


def timed(fn, reps=5):
    for i in range(SETS):
        fn(i)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(SETS):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / SETS)
    return best


kind = "mma"
for prec in ("bf16", "bf16x3"):
    with b200.precision(prec):
        f = timed(lambda i: b200.attention_forward(*views(i), mask, -1e9, scale, 0.1, ticket, want_planes=True))
        bw = timed(lambda i: b200.attention_backward(*views(i), mask, -1e9, scale, 0.1, ticket, dO[i]))
    print(f"attention[{kind}] {prec}: forward {f:.1f} us, backward {bw:.1f} us per launch (host-paced, includes allocation)")
