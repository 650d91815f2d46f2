"""Generate tests/golden/*.npz by running the UNMODIFIED reference (``/root/reference``).

Run HERE (the build container) only -- ``/root/reference`` does not exist on the GPU box:

    python oracle/make_golden.py

The reference does ``import cupy`` unconditionally (neunet/__init__.py:5, autograd.py:3, ...), and
CuPy is not installed, so a stub module that re-exports NumPy is placed in ``sys.modules`` first; the
CPU path (device="cpu") never calls into it. ``neunet.nn.experimental`` is never imported.

Every fixture stores the inputs (fp32) and the reference's outputs/gradients. Shapes are kept small
so the whole directory stays around a megabyte; the first Linear case is the reference's own test
shape (tests/test_linear_cuda.py:17-20) scaled down 4x per dim, plus one full-size copy of it.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REF = os.environ.get("NEUNET_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    stub = types.ModuleType("cupy")
    for name in dir(np):
        if not name.startswith("__"):
            setattr(stub, name, getattr(np, name))
    stub.ndarray = np.ndarray
    sys.modules["cupy"] = stub
    sys.path.insert(0, REF)
    import neunet  # noqa: E402
    import neunet.nn as nn  # noqa: E402
    from neunet import optim  # noqa: E402

    assert os.path.realpath(neunet.__file__).startswith(os.path.realpath(REF)), neunet.__file__
    return neunet, nn, optim


def main():
    neunet, nn, optim = import_reference()
    from neunet.autograd import Tensor

    os.makedirs(OUT, exist_ok=True)
    f32 = np.float32

    def save(name, **arrs):
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
        print(f"{name}: " + ", ".join(f"{k}{tuple(np.shape(v))}" for k, v in arrs.items()))

    # ---- nn.Linear: 2-D (reference test shape /4 and full), 3-D batched, no bias ---------------
    for tag, (lead, k, n, bias) in {
        "linear_ref_small": ((32,), 64, 128, True),
        "linear_ref_shape": ((128,), 256, 512, True),   # tests/test_linear_cuda.py:17-20
        "linear_3d": ((3, 5), 24, 40, True),
        "linear_nobias": ((7,), 6, 6, False),          # tests/test_reparam_slicing_cpu.py shape family
        "linear_n10": ((64,), 128, 10, True),           # MLP head 128 -> 10
    }.items():
        np.random.seed(42)
        layer = nn.Linear(k, n, bias=bias)
        x = np.random.uniform(-1, 1, lead + (k,)).astype(f32)
        g = np.random.uniform(-1, 1, lead + (n,)).astype(f32)
        xt = Tensor(x.copy(), requires_grad=True)
        out = layer(xt)
        out.backward(g.copy())
        arrs = dict(x=x, w=layer.weight.data.copy(), g=g, out=out.data, dx=xt.grad, dw=layer.weight.grad)
        if bias:
            arrs.update(b=layer.bias.data.copy(), db=layer.bias.grad)
        save(tag, **arrs)

    # ---- Tensor.matmul: 2-D, batched 4-D with transposed views (GPT attention), broadcast, vec ---
    np.random.seed(7)
    cases = {
        "matmul_2d": ((9, 13), (13, 11)),
        "matmul_4d": ((2, 3, 8, 16), (2, 3, 16, 8)),
        "matmul_bcast": ((4, 6, 10), (10, 5)),
        "matmul_vecmat": ((12,), (12, 7)),
        "matmul_matvec": ((5, 12), (12,)),
        "matmul_vecvec": ((12,), (12,)),
    }
    for tag, (sa, sb) in cases.items():
        a = np.random.uniform(-1, 1, sa).astype(f32)
        b = np.random.uniform(-1, 1, sb).astype(f32)
        ta, tb = Tensor(a.copy(), requires_grad=True), Tensor(b.copy(), requires_grad=True)
        out = ta.matmul(tb)
        g = np.random.uniform(-1, 1, out.shape).astype(f32)
        out.backward(g.copy())
        save(tag, a=a, b=b, g=g, out=out.data, da=ta.grad, db=tb.grad)
    # q.k^T with a transposed view, scaled like examples/gpt.ipynb cell 2
    q = np.random.uniform(-1, 1, (2, 6, 4, 8)).astype(f32)   # (B, T, h, d) storage
    kk = np.random.uniform(-1, 1, (2, 6, 4, 8)).astype(f32)
    tq, tk = Tensor(q.copy(), requires_grad=True), Tensor(kk.copy(), requires_grad=True)
    s = tq.transpose(0, 2, 1, 3).matmul(tk.transpose(0, 2, 1, 3).transpose(0, 1, 3, 2))
    g = np.random.uniform(-1, 1, s.shape).astype(f32)
    s.backward(g.copy())
    save("matmul_attn_views", q=q, k=kk, g=g, out=s.data, dq=tq.grad, dk=tk.grad)

    # ---- README autograd known answer (README.md:144-170) ---------------------------------------
    x = neunet.tensor([[7.0, 6.0, 5.0], [4.0, 5.0, 6.0]], requires_grad=True)
    y = neunet.tensor([[1.1, 2.2], [3.3, 4.4], [5.5, 6.6]], requires_grad=True)
    z = neunet.tensor([[2.3, 3.4], [4.5, 5.6]], requires_grad=True)
    out = neunet.tanh(1 / neunet.log(neunet.concatenate([(x @ y) @ z, neunet.exp(x) / neunet.sqrt(x)], axis=1)))
    out.backward(np.ones_like(out.data))
    save("readme_autograd", out=out.data, dx=x.grad, dy=y.grad, dz=z.grad)

    # ---- nn.Conv2d -------------------------------------------------------------------------------
    conv_cases = {
        # tag: (B, Cin, H, W, Cout, k, stride, padding, dilation, bias)
        "conv_3x3_p1": (2, 3, 9, 9, 4, 3, 1, 1, 1, True),          # cfg 3 / DDPM 3x3 family
        "conv_mnist1": (2, 1, 28, 28, 8, 3, 1, 1, 1, True),        # conv classifier layer 1
        "conv_4x4_s2_p1": (2, 4, 8, 8, 6, 4, 2, 1, 1, True),       # DDPM down-sample
        "conv_s2_odd": (2, 3, 10, 11, 5, 3, 2, 0, 1, False),       # stride does not tile the input
        "conv_dil2": (1, 2, 12, 12, 3, 3, 1, 2, 2, True),
        "conv_rect": (2, 3, 9, 12, 4, (2, 3), (1, 2), (1, 0), (1, 1), True),
        # NB: the string paddings ("same", "real same", "valid") cannot be exercised: Conv2d.__init__
        # wraps any non-tuple padding as (p, p) (conv2d.py:162), so build() never sees the bare string
        # and fails with TypeError at conv2d.py:246 -- a reference bug, recorded in DESIGN.md.
        "conv_asym_pad": (2, 2, 7, 8, 3, 3, 1, (1, 2), 1, True),
        "conv_wide": (2, 16, 6, 6, 24, 3, 1, 1, 1, True),
    }
    for tag, (bsz, cin, h, w, cout, k, s, p, d, bias) in conv_cases.items():
        np.random.seed(11)
        layer = nn.Conv2d(cin, cout, k, s, p, d, bias=bias)
        if bias:
            layer.bias.data = np.random.uniform(-0.5, 0.5, cout).astype(f32)  # init is zeros; make it bite
        x = np.random.uniform(-1, 1, (bsz, cin, h, w)).astype(f32)
        xt = Tensor(x.copy(), requires_grad=True)
        w0 = layer.weight.data.copy()
        out = layer(xt)
        g = np.random.uniform(-1, 1, out.shape).astype(f32)
        out.backward(g.copy())
        arrs = dict(x=x, w=w0, g=g, out=out.data, dx=xt.grad, dw=layer.weight.grad,
                    stride=np.array(layer.stride), pad4=np.array(layer.padding), dil=np.array(layer.dilation))
        if bias:
            arrs.update(b=layer.bias.data.copy(), db=layer.bias.grad)
        assert np.array_equal(layer.weight.data, w0), "dilation round-trip must restore the weight"
        save(tag, **arrs)

    # ---- Swish / Softmax / RMSNorm --------------------------------------------------------------
    np.random.seed(5)
    for beta in (1.0, 1.5):
        x = np.random.randn(6, 33).astype(f32) * 2
        g = np.random.randn(6, 33).astype(f32)
        xt = Tensor(x.copy(), requires_grad=True)
        out = nn.Swish(beta)(xt)
        out.backward(g.copy())
        save(f"swish_b{beta}", x=x, g=g, out=out.data, dx=xt.grad, beta=np.array(beta))
    for tag, shape, axis in (("softmax_last", (4, 3, 17), -1), ("softmax_axis1", (5, 19, 6), 1),
                             ("softmax_ref", (32, 128), -1)):   # tests/test_softmax_cuda.py:18
        x = (np.random.randn(*shape) * 3).astype(f32)
        g = np.random.randn(*shape).astype(f32)
        xt = Tensor(x.copy(), requires_grad=True)
        out = nn.Softmax(axis=axis)(xt)
        out.backward(g.copy())
        save(tag, x=x, g=g, out=out.data, dx=xt.grad, axis=np.array(axis))
    for tag, shape, bias in (("rmsnorm_2d", (32, 128), False), ("rmsnorm_3d_bias", (3, 7, 48), True)):
        layer = nn.RMSNorm(shape[-1], bias=bias)
        layer.weight.data = np.random.uniform(0.5, 1.5, shape[-1]).astype(f32)
        if bias:
            layer.bias.data = np.random.uniform(-0.5, 0.5, shape[-1]).astype(f32)
        x = np.random.randn(*shape).astype(f32)
        g = np.random.randn(*shape).astype(f32)
        xt = Tensor(x.copy(), requires_grad=True)
        out = layer(xt)
        out.backward(g.copy())
        arrs = dict(x=x, w=layer.weight.data.copy(), g=g, out=out.data, dx=xt.grad, dw=layer.weight.grad,
                    eps=np.array(layer.eps))
        if bias:
            arrs.update(b=layer.bias.data.copy(), db=layer.bias.grad)
        save(tag, **arrs)

    # ---- Adam / AdamW, three steps, two tensors (one with grad None) ----------------------------
    for tag, cls, kw in (("adam", optim.Adam, dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8)),
                         ("adam_l2", optim.Adam, dict(lr=1e-2, betas=(0.9, 0.98), eps=1e-9, weight_decay=0.1)),
                         ("adamw", optim.AdamW, dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)),
                         ("adamw_nowd", optim.AdamW, dict(lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0))):
        np.random.seed(42)
        p0 = np.random.randn(17, 9).astype(f32)
        p1 = np.random.randn(33).astype(f32)
        grads = np.random.randn(3, 17, 9).astype(f32)
        params = [Tensor(p0.copy(), requires_grad=True), Tensor(p1.copy(), requires_grad=True)]
        opt = cls(params, **kw)
        traj = []
        for t in range(3):
            params[0].grad = grads[t].copy()
            params[1].grad = None  # skipped, like unused cross_attn params in the GPT example
            opt.step()
            traj.append(params[0].data.copy())
        save("opt_" + tag, p0=p0, p1=p1, grads=grads, traj=np.stack(traj), p1_after=params[1].data,
             m=opt.m[0], v=opt.v[0], lr=np.array(kw["lr"]), betas=np.array(kw["betas"]), eps=np.array(kw["eps"]),
             wd=np.array(kw.get("weight_decay", 0.0 if cls is optim.Adam else 0.01)))

    # ---- quick-start MLP step (README.md:44-71, Swish variant = BASELINE config 2 at small batch) -
    np.random.seed(0)
    l1, l2 = nn.Linear(20, 16), nn.Linear(16, 10)
    act = nn.Swish()
    x = np.random.randn(12, 20).astype(f32)
    labels = np.random.randint(0, 10, 12).astype(np.int32)
    w10, b10, w20, b20 = (l1.weight.data.copy(), l1.bias.data.copy(), l2.weight.data.copy(), l2.bias.data.copy())
    params = l1.parameters() + l2.parameters()
    opt = optim.AdamW(params, lr=1e-3)
    loss_fn = nn.CrossEntropyLoss()
    losses = []
    for _ in range(2):
        opt.zero_grad()
        out = l2(act(l1(neunet.tensor(x))))
        loss = loss_fn(out, neunet.tensor(labels, dtype=np.int32))
        loss.backward()
        opt.step()
        losses.append(loss.data.copy())
    save("mlp_step", x=x, labels=labels, w1=w10, b1=b10, w2=w20, b2=b20, losses=np.array(losses),
         w1_after=l1.weight.data, b1_after=l1.bias.data, w2_after=l2.weight.data, b2_after=l2.bias.data,
         dw1=l1.weight.grad, dw2=l2.weight.grad)


def model_goldens():
    """Whole-model fixtures: the SAME model definitions (examples/models.py) run on the reference."""
    neunet, nn, optim = import_reference()
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(OUT)), "examples"))
    import models as M
    f32 = np.float32

    def save(name, **arrs):
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
        print(f"{name}: " + ", ".join(f"{k}{tuple(np.shape(v))}" for k, v in arrs.items()))

    def flat_params(model):
        return {f"p{i}": p.data.copy() for i, p in enumerate(model.parameters())}

    # ---- GPT (examples/gpt.ipynb architecture), eval mode (dropout off), 2 Adam steps -------------
    np.random.seed(0)
    model = M.build_gpt(neunet, nn)
    model.eval()
    init = flat_params(model)
    opt = optim.Adam(model.parameters(), lr=1.5e-4, betas=(0.9, 0.98), eps=1e-9)
    rng = np.random.RandomState(1)
    batch = rng.randint(3, 50, (3, 11))
    batch[0, 8:] = 0  # padding in one row (ignore_index / pad mask path)
    losses, first_logits, grads = [], None, None
    for step in range(2):
        opt.zero_grad()
        loss, logits = M.gpt_train_step(neunet, nn, model, opt, batch)
        losses.append(float(loss.data))
        if step == 0:
            first_logits = logits.data.copy()
            grads = {f"g{i}": (p.grad.copy() if p.grad is not None else np.zeros(0, f32))
                     for i, p in enumerate(model.parameters())}
    after = {f"a{i}": p.data.copy() for i, p in enumerate(model.parameters())}
    save("model_gpt", batch=batch, losses=np.array(losses), logits=first_logits, **init, **grads, **after)

    # ---- conv digits classifier, train mode (BatchNorm batch statistics), 2 Adam steps ------------
    np.random.seed(0)
    net = M.build_conv_classifier(neunet, nn)
    init = flat_params(net)
    opt = optim.Adam(net.parameters(), lr=1e-3)
    x = rng.uniform(-1, 1, (6, 1, 12, 12)).astype(f32)
    y = np.eye(10, dtype=f32)[rng.randint(0, 10, 6)]
    losses = []
    for step in range(2):
        opt.zero_grad()
        out = net(neunet.tensor(x))
        loss = nn.MSELoss()(out, neunet.tensor(y))
        loss.backward()
        opt.step()
        losses.append(float(loss.data))
        if step == 0:
            first_out = out.data.copy()
            grads = {f"g{i}": p.grad.copy() for i, p in enumerate(net.parameters())}
    after = {f"a{i}": p.data.copy() for i, p in enumerate(net.parameters())}
    save("model_conv_classifier", x=x, y=y, losses=np.array(losses), out=first_out,
         running_mean=net.bn.running_mean.data.copy(), **init, **grads, **after)

    # ---- two-level DDPM-style UNet, train mode, 1 Adam step -------------------------------------
    np.random.seed(0)
    unet = M.build_unet(neunet, nn)
    init = flat_params(unet)
    opt = optim.Adam(unet.parameters(), lr=2e-4)
    x = rng.uniform(-1, 1, (2, 3, 8, 8)).astype(f32)
    temb = rng.randn(2, 8).astype(f32)
    noise = rng.randn(2, 3, 8, 8).astype(f32)
    opt.zero_grad()
    out = unet(neunet.tensor(x), neunet.tensor(temb))
    loss = nn.MSELoss()(out, neunet.tensor(noise))
    loss.backward()
    opt.step()
    grads = {f"g{i}": p.grad.copy() for i, p in enumerate(unet.parameters())}
    after = {f"a{i}": p.data.copy() for i, p in enumerate(unet.parameters())}
    save("model_unet", x=x, temb=temb, noise=noise, loss=np.array(float(loss.data)), out=out.data.copy(),
         **init, **grads, **after)

    # ---- examples/ddpm.ipynb's SimpleUNet (examples/models.py: build_ddpm_unet) at small widths: 8x8 images, down
    # (8, 16, 32), up (32, 16, 8) -- same layer graph as the notebook (3 levels), one Algorithm-1 step with Adam
    np.random.seed(0)
    du = M.build_ddpm_unet(neunet, nn, image_size=8, down_channels=(8, 16, 32), up_channels=(32, 16, 8))
    init = flat_params(du)
    opt = optim.Adam(du.parameters(), lr=2e-4)
    x0 = rng.uniform(-1, 1, (3, 3, 8, 8)).astype(f32)
    noise = rng.randn(3, 3, 8, 8).astype(f32)
    t_frac = np.array([0.1, 0.45, 0.9], dtype=f32)
    a = np.array([0.95, 0.6, 0.2], dtype=f32).reshape(3, 1, 1, 1)
    b = np.sqrt(1 - a * a).astype(f32)
    opt.zero_grad()
    loss, pred = M.ddpm_train_step(neunet, nn, du, opt, neunet.tensor(x0), neunet.tensor(noise), t_frac, neunet.tensor(a),
                                   neunet.tensor(b))
    grads = {f"g{i}": p.grad.copy() for i, p in enumerate(du.parameters())}
    after = {f"a{i}": p.data.copy() for i, p in enumerate(du.parameters())}
    save("model_ddpm_unet", x0=x0, noise=noise, t_frac=t_frac, a=a, b=b, loss=np.array(float(loss.data)), out=pred.data.copy(),
         **init, **grads, **after)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "models":
        model_goldens()
    else:
        main()
        model_goldens()
