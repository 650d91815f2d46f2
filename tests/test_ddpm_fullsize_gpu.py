"""Full-size DDPM UNet (examples/ddpm.ipynb cell 8: 3x32x32, down (128,256,512,1024), up (1024,512,256,128), 61.7 M
parameters) for one training step on the B200 against torch fp32 (TF32 off) as the independent second oracle of
SURVEY.md section 8(c) -- the NumPy reference needs ~2 minutes per sample at this size. The torch mirror below re-states
the model with F.conv2d / F.conv_transpose2d (weights flipped + transposed, the mapping of SURVEY 8c) /
F.batch_norm / F.leaky_relu on the SAME parameter arrays; compared: prediction, loss, every parameter gradient.
Also: the README conv classifier at its full size and batch 512 (BASELINE.json configs[2]) one step vs torch fp32."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _t(p):
    return p.data.detach().clone().requires_grad_(True)


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def _torch_unet(model, x_t, tf):
    """Mirror of examples/models.py: build_ddpm_unet.forward with torch ops; returns (pred, [torch params in
    model.parameters() order])."""
    import torch.nn.functional as F
    params = {}

    def P(par):
        if par is None:
            return None
        if id(par) not in params:
            params[id(par)] = _t(par)
        return params[id(par)]

    def conv(layer, x):
        return F.conv2d(x, P(layer.weight), P(layer.bias), stride=layer.stride, padding=layer.padding[:2] if len(layer.padding) == 2 else layer.padding)

    def convT(layer, x):
        w = torch.flip(P(layer.weight), (2, 3)).transpose(0, 1)
        return F.conv_transpose2d(x, w, P(layer.bias), stride=layer.stride, padding=layer.padding)

    def bn(layer, x):
        return F.batch_norm(x, None, None, P(layer.weight).reshape(-1), P(layer.bias).reshape(-1), training=True, eps=layer.eps)

    def lin(layer, x):
        return F.linear(x, P(layer.weight), P(layer.bias).reshape(-1))

    def block(b, x, t):
        h = bn(b.bnorm1, F.leaky_relu(conv(b.conv1, x), 0.01))
        h = h + F.leaky_relu(lin(b.time_embedding, t), 0.01)[:, :, None, None]
        h = bn(b.bnorm2, F.leaky_relu(conv(b.conv2, h), 0.01))
        return convT(b.transform, h) if type(b.transform).__name__ == "ConvTranspose2d" else conv(b.transform, h)

    te = model.time_embedding
    mods = [te[i] for i in range(len(te))] if hasattr(te, "__len__") else list(te.modules)
    t = tf + mods[0].pe.data[: tf.shape[0], :]
    t = F.leaky_relu(lin(mods[1], t), 0.01).reshape(tf.shape[0], -1)
    x = conv(model.input_conv, x_t)
    skips = []
    for d in model.down_layers:
        x = block(d, x, t)
        skips.append(x)
    for u in model.up_layers:
        x = block(u, torch.cat((x, skips.pop()), 1), t)
    out = convT(model.output_conv, x)
    return out, [params.get(id(p)) for p in model.parameters()]


def test_ddpm_unet_full_size_step_vs_torch_fp32():
    import models as M
    import neunet
    import neunet.nn as nn
    from neunet import b200
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    b200.require_device()
    np.random.seed(0)
    B = 8
    model = M.build_ddpm_unet(neunet, nn, device="cuda")
    rng = np.random.RandomState(3)
    x0 = rng.uniform(-1, 1, (B, 3, 32, 32)).astype(np.float32)
    noise = rng.randn(B, 3, 32, 32).astype(np.float32)
    a = rng.uniform(0.2, 0.95, (B, 1, 1, 1)).astype(np.float32)
    b = np.sqrt(1 - a * a).astype(np.float32)
    tf = rng.uniform(0.05, 0.95, (B, 1, 1)).astype(np.float32)
    x_t = a * x0 + b * noise
    with b200.precision("bf16x3"):
        pred = model.forward(neunet.tensor(x_t, device="cuda"), neunet.tensor(tf, device="cuda"))
        loss = nn.MSELoss()(pred, neunet.tensor(noise, device="cuda"))
        loss.backward()
    ref, tparams = _torch_unet(model, torch.from_numpy(x_t).cuda(), torch.from_numpy(tf).cuda())
    rloss = ((ref - torch.from_numpy(noise).cuda()) ** 2).mean()
    rloss.backward()
    assert abs(float(loss.item()) - float(rloss.item())) <= 1e-4 * abs(float(rloss.item()))
    # 17 conv layers + 12 BatchNorms deep, bf16x3 contractions (~5e-6 each): max-norm relative error
    assert _rel(pred.data, ref.detach()) < 2e-4
    pairs = [(i, p.grad.reshape(tp.grad.shape), tp.grad) for i, (p, tp) in enumerate(zip(model.parameters(), tparams))
             if tp is not None and tp.grad is not None]
    assert len(pairs) >= 50
    gmax = max(w.abs().max().item() for _, _, w in pairs)
    report = []
    for i, got, want in pairs:
        # every tensor is measured against max(its own scale, 2 % of the largest gradient in the model): biases that feed a
        # BatchNorm have gradients that cancel almost exactly and are mostly round-off in BOTH implementations
        floor_l2 = max(want.norm().item(), 0.02 * gmax * want.numel() ** 0.5)
        cos = (got * want).sum().item() / max(got.norm().item() * want.norm().item(), 1e-30)
        report.append(((got - want).norm().item() / floor_l2, cos, i, tuple(want.shape)))
    # Measured on B200 (round 2): worst tensor 2.1e-2, identical to 4 digits with the implicit-GEMM paths on or off and with
    # fusion on or off -- i.e. it is the bf16x3 operand rounding (5e-6 per contraction, tests/test_conv_implicit_gpu.py holds
    # every layer to 1e-4 against the same torch oracle) amplified by 12 BatchNorm backward passes at batch 8 with
    # random-init weights, not a property of any one kernel. Bars = measured x 2; the small reference-generated golden
    # (tests/test_models.py::test_ddpm_unet_gpu) holds the same code to 1e-3.
    assert max(r[0] for r in report) < 5e-2, sorted(report, reverse=True)[:6]
    big = [r for r in report if r[0] > 0 and r[3] and np.prod(r[3]) >= 1024]
    assert min(r[1] for r in big) > 0.999, sorted(big, key=lambda r: r[1])[:6]


def test_conv_classifier_full_size_batch512_step_vs_torch_fp32():
    """BASELINE.json configs[2] at its real size: README model, 28x28, channels (8, 16), batch 512, one MSE step."""
    import torch.nn.functional as F
    import models as M
    import neunet
    import neunet.nn as nn
    from neunet import b200
    b200.require_device()
    np.random.seed(1)
    net = M.build_conv_classifier(neunet, nn, device="cuda", side=28, channels=(8, 16))
    rng = np.random.RandomState(5)
    x = rng.uniform(-1, 1, (512, 1, 28, 28)).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rng.randint(0, 10, 512)]
    with b200.precision("bf16x3"):
        out = net.forward(neunet.tensor(x, device="cuda"))
        loss = nn.MSELoss()(out, neunet.tensor(y, device="cuda"))
        loss.backward()
    ps = [_t(p) for p in net.parameters()]
    w1, b1, w2, b2, bw, bb, fw, fb = ps
    h = F.max_pool2d(F.leaky_relu(F.conv2d(torch.from_numpy(x).cuda(), w1, b1, padding=1), 0.01), 2, 2)
    h = F.max_pool2d(F.leaky_relu(F.conv2d(h, w2, b2, padding=1), 0.01), 2, 2)
    h = F.batch_norm(h, None, None, bw.reshape(-1), bb.reshape(-1), training=True, eps=1e-5)
    ref = torch.sigmoid(F.linear(h.reshape(512, -1), fw, fb.reshape(-1)))
    rloss = ((ref - torch.from_numpy(y).cuda()) ** 2).mean()
    rloss.backward()
    assert _rel(out.data, ref.detach()) < 1e-4
    assert abs(float(loss.item()) - float(rloss.item())) <= 1e-4 * abs(float(rloss.item()))
    for p, tp in zip(net.parameters(), ps):
        if not p.requires_grad:
            continue
        assert _rel(p.grad.reshape(tp.grad.shape), tp.grad) < 5e-4
