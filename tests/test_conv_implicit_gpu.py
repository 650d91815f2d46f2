"""Implicit-GEMM Conv2d (4-D TMA boxes over channels-last planes, no `col` matrix) and the native gather-form
ConvTranspose2d on the B200, through the C-ABI, against torch's fp32 convolutions (TF32 off) as the independent
second oracle of SURVEY.md section 8(c): nn.Conv2d == F.conv2d with the same (out,in,kh,kw) weights;
nn.ConvTranspose2d == F.conv_transpose2d(x, W', b, stride, padding) with W' = flip(W, (2,3)).transpose(0,1)
(reference: neunet/nn/layers/conv2d.py:16-117, 297-355; convtranspose2d.py:165-181, 321), and against the
zero-stuffed formulation of this repository. Error = max|d| / max|ref| (max-norm)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = {"bf16x3": 1e-4, "bf16": 6e-3}


def relerr(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def _setup():
    from neunet import b200
    b200.require_device()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return b200


@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
@pytest.mark.parametrize("B,Cin,Cout,HW,k,s,p", [
    (4, 64, 128, 32, 3, 1, 1),     # one 64-channel k-block per tap, 32-wide rows
    (8, 128, 64, 16, 3, 1, 1),     # box = 16 x 8 pixels
    (16, 256, 256, 8, 3, 1, 1),    # box spans two images
    (64, 128, 128, 4, 3, 1, 1),    # 4 x 4 images: eight images per box
    (4, 128, 128, 32, 4, 2, 1),    # DDPM down-sample: strided boxes forward, zero-stuffed dgrad
    (3, 64, 64, 16, 3, 1, 1),      # ragged last box (B not a multiple of the images per box)
    (2, 40, 48, 8, 3, 1, 1),       # channel counts that are not multiples of 64: ragged 64-channel layout tiles
    (2, 72, 80, 12, 3, 1, 1),      # ... spanning two tiles; 12 x 12 maps (144 pixels: ragged 32-pixel tiles)
])
def test_conv2d_implicit_gemm_vs_torch(prec, B, Cin, Cout, HW, k, s, p):
    b200 = _setup()
    with b200.precision(prec):
        gen = torch.Generator(device="cuda").manual_seed(B * 1000 + Cin)
        x = (torch.rand(B, Cin, HW, HW, device="cuda", generator=gen) * 2 - 1).requires_grad_(True)
        w = ((torch.rand(Cout, Cin, k, k, device="cuda", generator=gen) * 2 - 1) / (Cin * k * k) ** 0.5).requires_grad_(True)
        bias = (torch.rand(Cout, device="cuda", generator=gen) - 0.5).requires_grad_(True)
        ref = torch.nn.functional.conv2d(x, w, bias, stride=s, padding=p)
        g = torch.rand(ref.shape, device="cuda", generator=gen) * 2 - 1
        ref.backward(g)
        out, planes = b200.conv2d_forward(x.detach(), w.detach(), bias.detach(), (s, s), (p, p, p, p), (1, 1), keep_planes=True)
        assert planes is not None
        assert relerr(out, ref.detach()) < TOL[prec]
        for xp in (planes, None):  # planes kept from forward, and converted again
            dx, dw, db = b200.conv2d_backward(x.detach(), w.detach(), g, (s, s), (p, p, p, p), (1, 1), x_planes=xp)
            assert relerr(dx, x.grad) < TOL[prec]
            assert relerr(dw, w.grad) < TOL[prec]
            assert relerr(db, bias.grad) < 1e-4


@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
def test_conv2d_few_output_channels_long_reduction_vs_torch(prec):
    """DDPM's output layer shape (128 -> 3, 3x3): few output channels over a long reduction run on the tensor cores
    (a 3-row A tile), not in the direct kernels."""
    b200 = _setup()
    with b200.precision(prec):
        gen = torch.Generator(device="cuda").manual_seed(77)
        x = (torch.rand(8, 128, 34, 34, device="cuda", generator=gen) * 2 - 1).requires_grad_(True)
        w = ((torch.rand(3, 128, 3, 3, device="cuda", generator=gen) * 2 - 1) / 1152 ** 0.5).requires_grad_(True)
        bias = (torch.rand(3, device="cuda", generator=gen) - 0.5).requires_grad_(True)
        ref = torch.nn.functional.conv2d(x, w, bias)
        g = torch.rand(ref.shape, device="cuda", generator=gen) * 2 - 1
        ref.backward(g)
        out = b200.conv2d_forward(x.detach(), w.detach(), bias.detach(), (1, 1), (0, 0, 0, 0), (1, 1))
        dx, dw, db = b200.conv2d_backward(x.detach(), w.detach(), g, (1, 1), (0, 0, 0, 0), (1, 1))
        assert relerr(out, ref.detach()) < TOL[prec]
        assert relerr(dx, x.grad) < TOL[prec] and relerr(dw, w.grad) < TOL[prec] and relerr(db, bias.grad) < 1e-4


def test_conv2d_implicit_equals_materialised_col(monkeypatch):
    """Same layer through the implicit path and (fresh process env not needed: the C side reads NNB_CONV_IMPLICIT once)
    against the torch reference at both precisions is covered above; here: forward twice gives bit-identical results
    (deterministic tile order, no atomics)."""
    b200 = _setup()
    with b200.precision("bf16"):
        x = torch.randn(4, 128, 16, 16, device="cuda")
        w = torch.randn(128, 128, 3, 3, device="cuda") * 0.03
        a = b200.conv2d_forward(x, w, None, (1, 1), (1, 1, 1, 1), (1, 1))
        b = b200.conv2d_forward(x, w, None, (1, 1), (1, 1, 1, 1), (1, 1))
        assert torch.equal(a, b)


@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
@pytest.mark.parametrize("B,C,Cout,HW,k,s,p", [
    (4, 128, 128, 16, 4, 2, 1),    # DDPM up-sample 16 -> 32
    (8, 64, 128, 8, 4, 2, 1),
    (16, 256, 64, 4, 4, 2, 1),     # 4 x 4 -> 8 x 8
    (2, 64, 64, 8, 3, 1, 1),       # stride 1: a single class
])
def test_conv_transpose2d_native_vs_torch(prec, B, C, Cout, HW, k, s, p):
    b200 = _setup()
    with b200.precision(prec):
        gen = torch.Generator(device="cuda").manual_seed(C + Cout + HW)
        x = (torch.rand(B, C, HW, HW, device="cuda", generator=gen) * 2 - 1).requires_grad_(True)
        w = ((torch.rand(Cout, C, k, k, device="cuda", generator=gen) * 2 - 1) / (C * k * k) ** 0.5).requires_grad_(True)
        bias = (torch.rand(Cout, device="cuda", generator=gen) - 0.5).requires_grad_(True)
        assert b200.conv_transpose2d_supported(x.shape, w.shape, (s, s), (p, p, p, p), (1, 1), (0, 0))
        # reference semantics in torch terms: un-flipped (out,in,kh,kw) correlation over the stuffed input
        wt = torch.flip(w, (2, 3)).transpose(0, 1)
        ref = torch.nn.functional.conv_transpose2d(x, wt, bias, stride=s, padding=p)
        g = torch.rand(ref.shape, device="cuda", generator=gen) * 2 - 1
        ref.backward(g)
        out, planes = b200.conv_transpose2d_forward(x.detach(), w.detach(), bias.detach(), (s, s), (p, p, p, p), (1, 1), (0, 0))
        assert out.shape == ref.shape
        assert relerr(out, ref.detach()) < TOL[prec]
        dx, dw, db = b200.conv_transpose2d_backward(x.detach(), w.detach(), g, (s, s), (p, p, p, p), (1, 1), (0, 0), x_planes=planes)
        assert relerr(dx, x.grad) < TOL[prec]
        assert relerr(dw, w.grad) < TOL[prec]
        assert relerr(db, bias.grad) < 1e-4


def test_conv_transpose_layer_native_equals_zero_stuffed_path():
    """nn.ConvTranspose2d through the public API: the native gather form (64-multiple channels) against this repository's
    zero-stuffed formulation over nn.Conv2d kernels (forced by an unsupported channel count is not comparable, so the same
    layer is run with the native path disabled via output_padding-free fallback: weights copied to a CPU twin)."""
    import neunet
    import neunet.nn as nn
    from neunet import b200
    np.random.seed(8)
    layer = nn.ConvTranspose2d(64, 64, 4, 2, 1)
    x = np.random.randn(2, 64, 8, 8).astype(np.float32)
    xc = neunet.tensor(x, requires_grad=True)
    yc = layer(xc)
    (yc * yc).mean().backward()
    cpu = [np.array(yc.data), np.array(xc.grad), np.array(layer.weight.grad), np.array(layer.bias.grad)]
    dev = layer.to("cuda")
    xd = neunet.tensor(x, device="cuda", requires_grad=True)
    with b200.precision("bf16x3"):
        yd = dev(xd)
        assert yd.op == "convtranspose2d" and len(yd.args) == 8  # native node
        (yd * yd).mean().backward()
    for got, want in zip((yd.data, xd.grad, dev.weight.grad, dev.bias.grad), cpu):
        assert relerr(got, torch.from_numpy(want).cuda().reshape(got.shape)) < 1e-4
