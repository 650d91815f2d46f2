#!/usr/bin/env python
"""Conv2d fwd / bwd (dgrad + wgrad + db) TFLOP/s on the BASELINE conv shapes (conv digits
classifier at B=512, DDPM UNet layers at B=64), bf16 operands / fp32 accumulate, through the public
binding (neunet.b200.conv2d_forward / conv2d_backward -> nnb_conv2d_*). Each case is captured into a
CUDA graph (all kernels of the call: gather/im2col + tcgen05 GEMM [+ split-K finish]) and the replay is
timed with CUDA events; inputs rotate over > L2 worth of buffers. Prints a markdown table."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "numpy-nn-model_b200"))
import torch  # noqa: E402
from neunet import b200  # noqa: E402

CASES = [  # name, B, Cin, H, W, Cout, k, stride, pad
    ("mnist conv1", 512, 1, 28, 28, 8, 3, 1, 1),
    ("mnist conv2", 512, 8, 14, 14, 16, 3, 1, 1),
    ("ddpm in 3->128 @32", 64, 3, 32, 32, 128, 3, 1, 1),
    ("ddpm down1 128->256 @32", 64, 128, 32, 32, 256, 3, 1, 1),
    ("ddpm down1 256->256 @32", 64, 256, 32, 32, 256, 3, 1, 1),
    ("ddpm down1 4x4s2 256 @32", 64, 256, 32, 32, 256, 4, 2, 1),
    ("ddpm down2 256->512 @16", 64, 256, 16, 16, 512, 3, 1, 1),
    ("ddpm down2 512->512 @16", 64, 512, 16, 16, 512, 3, 1, 1),
    ("ddpm down3 512->1024 @8", 64, 512, 8, 8, 1024, 3, 1, 1),
    ("ddpm down3 1024->1024 @8", 64, 1024, 8, 8, 1024, 3, 1, 1),
    ("ddpm up1 2048->512 @4", 64, 2048, 4, 4, 512, 3, 1, 1),
    ("ddpm up2 1024->256 @8", 64, 1024, 8, 8, 256, 3, 1, 1),
    ("ddpm up3 512->128 @16", 64, 512, 16, 16, 128, 3, 1, 1),
]


def timed_graph(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    b200._state["capture_epoch"] += 1
    with torch.cuda.graph(g):
        n = fn()
    b200._state["capture_epoch"] += 1
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * n)  # us per call


def main():
    torch.cuda.set_device(0)
    b200.require_device()
    b200.set_precision("bf16")
    print("| layer | B | Cin | HxW | Cout | k/s | GFLOP/pass | fwd us | fwd TFLOP/s | bwd (dX+dW+db) us | bwd TFLOP/s | kernels fwd/bwd |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for name, B, Cin, H, W, Cout, k, s, p in CASES:
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        fl = 2.0 * B * Ho * Wo * Cout * Cin * k * k
        per = 4 * (B * Cin * H * W + 2 * B * Cout * Ho * Wo)
        sets = int(min(16, max(2, -(-(260 << 20) // per))))
        xs = [torch.rand(B, Cin, H, W, device="cuda") * 2 - 1 for _ in range(sets)]
        gs = [torch.rand(B, Cout, Ho, Wo, device="cuda") * 2 - 1 for _ in range(sets)]
        w = (torch.rand(Cout, Cin, k, k, device="cuda") * 2 - 1) / (Cin * k * k) ** 0.5
        bias = torch.zeros(Cout, device="cuda")
        st, pad, dil = (s, s), (p, p, p, p), (1, 1)

        def fwd():
            for x in xs:
                b200.conv2d_forward(x, w, bias, st, pad, dil)
            return sets

        def bwd():
            for x, g in zip(xs, gs):
                b200.conv2d_backward(x, w, g, st, pad, dil, need_dx=True, need_db=True)
            return sets

        b200.reset_launch_count(); b200.conv2d_forward(xs[0], w, bias, st, pad, dil); kf = b200.launch_count()
        b200.reset_launch_count(); b200.conv2d_backward(xs[0], w, gs[0], st, pad, dil); kb = b200.launch_count()
        tf, tb = timed_graph(fwd), timed_graph(bwd)
        print(f"| {name} | {B} | {Cin} | {H}x{W} | {Cout} | {k}/{s} | {fl/1e9:.2f} | {tf:.1f} | {fl/tf*1e-6:.1f} | {tb:.1f} | "
              f"{2*fl/tb*1e-6:.1f} | {kf}/{kb} |")
        del xs, gs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
