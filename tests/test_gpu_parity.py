"""GPU parity tests (run with -m gpu on the B200 box). Every case goes neunet API -> ctypes C-ABI ->
sm_100a kernels and is compared with the oracle (oracle/restated.py) on the same seeded inputs and
with the reference-generated golden vectors.

Tolerances (stated per the north-star): contractions in BF16X3 mode <= 1e-4 relative to fp32 NumPy
(measured ~5e-6); plain BF16 mode <= 6e-3 (measured ~2e-3, reported not hidden); element-wise,
normalisation and optimizer kernels <= 1e-5; index/shape results bit-exact."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import neunet  # noqa: E402
import neunet.nn as nn  # noqa: E402
from conftest import load_golden  # noqa: E402
from neunet import Tensor, b200  # noqa: E402
from neunet.optim import Adam, AdamW  # noqa: E402
from oracle import restated as R  # noqa: E402


@pytest.fixture(autouse=True)
def _device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    b200.require_device()
    b200.set_precision("bf16x3")
    yield
    torch.cuda.synchronize()


def host(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def relerr(got, ref):
    got, ref = host(got).astype(np.float64), host(ref).astype(np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


X3, BF = 1e-4, 6e-3


def dev(a, rg=False, dtype=np.float32):
    return Tensor(a, requires_grad=rg, dtype=dtype, device="cuda")


# ---- nn.Linear ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["linear_ref_small", "linear_ref_shape", "linear_3d", "linear_nobias", "linear_n10"])
@pytest.mark.parametrize("prec,tol", [("bf16x3", X3), ("bf16", BF)])
def test_linear_golden(name, prec, tol):
    g = load_golden(name)
    b200.set_precision(prec)
    layer = nn.Linear(g["w"].shape[1], g["w"].shape[0], bias="b" in g, device="cuda")
    layer.weight.data = layer.xp.array(g["w"])
    if "b" in g:
        layer.bias.data = layer.xp.array(g["b"])
    x = dev(g["x"], True)
    out = layer(x)
    out.backward(g["g"])
    assert relerr(out.data, g["out"]) < tol
    assert relerr(x.grad, g["dx"]) < tol
    assert relerr(layer.weight.grad, g["dw"]) < tol
    if "b" in g:
        assert tuple(layer.bias.grad.shape) == g["db"].shape
        assert relerr(layer.bias.grad, g["db"]) < 1e-5


@pytest.mark.parametrize("M,K,N", [(4096, 784, 128), (4096, 128, 10), (1, 5, 3), (129, 65, 257), (2048, 512, 2048)])
def test_linear_seeded_vs_oracle(M, K, N):
    rng = np.random.RandomState(M + K + N)
    x = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    w = rng.uniform(-1, 1, (N, K)).astype(np.float32) / np.sqrt(K)
    b = rng.uniform(-1, 1, (1, N)).astype(np.float32)
    g = rng.uniform(-1, 1, (M, N)).astype(np.float32)
    layer = nn.Linear(K, N, device="cuda")
    layer.weight.data, layer.bias.data = layer.xp.array(w), layer.xp.array(b)
    xt = dev(x, True)
    out = layer(xt)
    out.backward(g)
    dx, dw, db = R.linear_backward(x, w, b, g)
    assert relerr(out.data, R.linear_forward(x, w, b)) < X3
    assert relerr(xt.grad, dx) < X3 and relerr(layer.weight.grad, dw) < X3 and relerr(layer.bias.grad, db) < 1e-5


def test_linear_swish_fused_vs_oracle():
    rng = np.random.RandomState(5)
    M, K, N, beta = 300, 96, 130, 1.5
    x, g = rng.randn(M, K).astype(np.float32), rng.randn(M, N).astype(np.float32)
    layer = nn.LinearSwish(K, N, beta=beta, device="cuda")
    w, b = host(layer.weight.data), host(layer.bias.data)
    xt = dev(x, True)
    out = layer(xt)
    out.backward(g)
    z = R.linear_forward(x, w, b)
    dz = R.swish_backward(z, g, beta)
    dx, dw, db = R.linear_backward(x, w, b, dz)
    assert relerr(out.data, R.swish_forward(z, beta)) < X3
    assert relerr(xt.grad, dx) < X3 and relerr(layer.weight.grad, dw) < X3 and relerr(layer.bias.grad, db) < 1e-4


def test_linear_skips_dgrad_for_non_grad_input_and_reuses_staged_weight():
    layer = nn.Linear(64, 32, device="cuda")
    x = dev(np.ones((8, 64), np.float32), False)
    out = layer(x)
    out.data  # noqa: B018 -- results are produced on first read (deferred evaluation): this launches the GEMM
    staged = layer.weight._b200_staged
    out2 = layer(x)
    out2.data  # noqa: B018
    assert layer.weight._b200_staged is staged  # cached until the weights change
    out.backward(np.ones((8, 32), np.float32))
    assert x.grad is None and layer.weight.grad is not None
    opt = AdamW(layer.parameters(), lr=1e-2)
    opt.step()
    layer(x).data  # noqa: B018
    assert layer.weight._b200_staged is not staged  # optimizer step invalidated the bf16 planes
    assert relerr(out.data, out2.data) == 0.0        # deterministic


def test_linear_full_size_properties():
    """BASELINE batch (4096 x 784 -> 128): linearity in X and agreement of a checksum of checksums
    with a float64 host contraction of the column/row sums (size-independent checks)."""
    rng = np.random.RandomState(0)
    M, K, N = 4096, 784, 128
    x1, x2 = rng.randn(M, K).astype(np.float32), rng.randn(M, K).astype(np.float32)
    layer = nn.Linear(K, N, bias=False, device="cuda")
    w = host(layer.weight.data).astype(np.float64)
    o1, o2, o12 = layer(dev(x1)).data, layer(dev(x2)).data, layer(dev(x1 + x2)).data
    assert relerr(o12, host(o1) + host(o2)) < 1e-4
    want = x1.astype(np.float64).sum(0) @ w.T.sum(1)
    got = float(host(o1).astype(np.float64).sum())
    assert abs(got - want) / (abs(want) + 1e-9) < 1e-3 or abs(got - want) < 1e-1


# ---- Tensor.matmul -------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["matmul_2d", "matmul_4d", "matmul_bcast", "matmul_vecmat", "matmul_matvec", "matmul_vecvec"])
def test_matmul_golden(name):
    g = load_golden(name)
    a, b = dev(g["a"], True), dev(g["b"], True)
    out = a @ b
    assert tuple(out.shape) == g["out"].shape
    out.backward(g["g"])
    assert relerr(out.data, g["out"]) < X3
    assert relerr(a.grad, g["da"]) < X3 and relerr(b.grad, g["db"]) < X3


def test_matmul_attention_transposed_views():
    g = load_golden("matmul_attn_views")
    q, k = dev(g["q"], True), dev(g["k"], True)
    s = q.transpose(0, 2, 1, 3).matmul(k.transpose(0, 2, 1, 3).transpose(0, 1, 3, 2))
    s.backward(g["g"])
    assert relerr(s.data, g["out"]) < X3
    assert relerr(q.grad, g["dq"]) < X3 and relerr(k.grad, g["dk"]) < X3


def test_matmul_gpt_attention_shapes():
    """(B,8,T,64) x (B,8,64,T) and (B,8,T,T) x (B,8,T,64) at B=4, T=64 (examples/gpt.ipynb cell 2)."""
    rng = np.random.RandomState(1)
    B, H, T, D = 4, 8, 64, 64
    q, k, v = (rng.randn(B, T, H, D).astype(np.float32) for _ in range(3))
    tq, tk, tv = dev(q, True), dev(k, True), dev(v, True)
    qh, kh, vh = (t.transpose(0, 2, 1, 3) for t in (tq, tk, tv))
    s = qh.matmul(kh.transpose(0, 1, 3, 2))
    o = s.matmul(vh)
    g = rng.randn(B, H, T, D).astype(np.float32)
    o.backward(g)
    qn, kn, vn = (a.transpose(0, 2, 1, 3) for a in (q, k, v))
    sn = R.matmul_forward(qn, kn.transpose(0, 1, 3, 2))
    on = R.matmul_forward(sn, vn)
    ds, dv = R.matmul_backward(sn, vn, g)
    dq, dkt = R.matmul_backward(qn, kn.transpose(0, 1, 3, 2), ds)
    assert relerr(o.data, on) < X3
    assert relerr(tv.grad, dv.transpose(0, 2, 1, 3)) < X3
    assert relerr(tq.grad, dq.transpose(0, 2, 1, 3)) < X3
    assert relerr(tk.grad, dkt.transpose(0, 1, 3, 2).transpose(0, 2, 1, 3)) < X3


def test_readme_autograd_example_on_device():
    g = load_golden("readme_autograd")
    x = neunet.tensor([[7.0, 6.0, 5.0], [4.0, 5.0, 6.0]], requires_grad=True, device="cuda")
    y = neunet.tensor([[1.1, 2.2], [3.3, 4.4], [5.5, 6.6]], requires_grad=True, device="cuda")
    z = neunet.tensor([[2.3, 3.4], [4.5, 5.6]], requires_grad=True, device="cuda")
    out = neunet.tanh(1 / neunet.log(neunet.concatenate([(x @ y) @ z, neunet.exp(x) / neunet.sqrt(x)], axis=1)))
    out.backward(np.ones((2, 5), np.float32))
    assert relerr(out.data, g["out"]) < 1e-5
    assert relerr(x.grad, g["dx"]) < 1e-4 and relerr(y.grad, g["dy"]) < 1e-4 and relerr(z.grad, g["dz"]) < 1e-4


# ---- nn.Conv2d --------------------------------------------------------------------------------------
CONV = ["conv_3x3_p1", "conv_mnist1", "conv_4x4_s2_p1", "conv_s2_odd", "conv_dil2", "conv_rect", "conv_asym_pad", "conv_wide"]


@pytest.mark.parametrize("name", CONV)
def test_conv2d_golden(name):
    g = load_golden(name)
    cout, cin, kh, kw = g["w"].shape
    layer = nn.Conv2d(cin, cout, (kh, kw), tuple(int(v) for v in g["stride"]), tuple(int(v) for v in g["pad4"]),
                      tuple(int(v) for v in g["dil"]), bias="b" in g, device="cuda")
    layer.weight.data = layer.xp.array(g["w"])
    if "b" in g:
        layer.bias.data = layer.xp.array(g["b"])
    x = dev(g["x"], True)
    out = layer(x)
    assert tuple(out.shape) == g["out"].shape
    out.backward(g["g"])
    assert relerr(out.data, g["out"]) < X3
    assert relerr(x.grad, g["dx"]) < X3 and relerr(layer.weight.grad, g["dw"]) < X3
    if "b" in g:
        assert relerr(layer.bias.grad, g["db"]) < 1e-5


@pytest.mark.parametrize("B,Cin,H,W,Cout,k,s,p,d", [
    (4, 64, 16, 16, 128, 3, 1, 1, 1),      # GEMM path, DDPM-style 3x3
    (3, 32, 16, 16, 64, 4, 2, 1, 1),       # DDPM down-sample 4x4 s2 p1
    (2, 48, 9, 11, 40, 3, 2, 1, 2),        # ragged: odd sizes, stride 2, dilation 2
    (2, 3, 32, 32, 128, 3, 1, 1, 1),       # DDPM input conv (K = 27)
    (5, 16, 8, 8, 32, 1, 1, 0, 1),         # 1x1
    (8, 8, 14, 14, 16, 3, 1, 1, 1),        # conv classifier layer 2 (direct kernels)
])
def test_conv2d_seeded_vs_oracle(B, Cin, H, W, Cout, k, s, p, d):
    rng = np.random.RandomState(B * 7 + Cin)
    x = rng.uniform(-1, 1, (B, Cin, H, W)).astype(np.float32)
    layer = nn.Conv2d(Cin, Cout, k, s, p, d, device="cuda")
    layer.bias.data = layer.xp.array(rng.uniform(-0.5, 0.5, Cout).astype(np.float32))
    w, b = host(layer.weight.data), host(layer.bias.data)
    pad4 = (p, p, p, p)
    ref = R.conv2d_forward(x, w, b, (s, s), pad4, (d, d))
    g = rng.uniform(-1, 1, ref.shape).astype(np.float32)
    xt = dev(x, True)
    out = layer(xt)
    out.backward(g)
    dx, dw, db = R.conv2d_backward(x, w, b, g, (s, s), pad4, (d, d))
    assert relerr(out.data, ref) < X3
    assert relerr(xt.grad, dx) < X3 and relerr(layer.weight.grad, dw) < X3 and relerr(layer.bias.grad, db) < 1e-5


def test_conv_transpose_on_device_vs_cpu():
    np.random.seed(3)
    cpu = nn.ConvTranspose2d(32, 48, 4, 2, 1)
    gpu = nn.ConvTranspose2d(32, 48, 4, 2, 1, device="cuda")
    gpu.weight.data, gpu.bias.data = gpu.xp.array(cpu.weight.data), gpu.xp.array(cpu.bias.data)
    x = np.random.randn(2, 32, 6, 6).astype(np.float32)
    xc, xg = Tensor(x, requires_grad=True), dev(x, True)
    oc, og = cpu(xc), gpu(xg)
    g = np.random.randn(*oc.shape).astype(np.float32)
    oc.backward(g), og.backward(g)
    assert relerr(og.data, oc.data) < X3
    assert relerr(xg.grad, xc.grad) < X3 and relerr(gpu.weight.grad, cpu.weight.grad) < X3


def test_conv_index_helpers_bit_exact_on_device():
    from neunet.nn.layers import conv2d as C
    a = np.arange(2 * 3 * 4 * 5, dtype=np.float32).reshape(2, 3, 4, 5)
    t = b200  # noqa: F841
    da = neunet.tensor(a, device="cuda").data
    p = (1, 2, 0, 3)
    assert np.array_equal(host(C.set_padding(da, p)), R.set_padding(a, p))
    assert np.array_equal(host(C.remove_padding(C.set_padding(da, p), p)), a)
    assert np.array_equal(host(C.set_stride(da, (2, 3))), R.set_stride(a, (2, 3)))
    assert np.array_equal(host(C.remove_stride(C.set_stride(da, (2, 3)), (2, 3))), a)


# ---- Swish / Softmax / RMSNorm ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["swish_b1.0", "swish_b1.5"])
def test_swish_golden(name):
    g = load_golden(name)
    x = dev(g["x"], True)
    out = nn.Swish(float(g["beta"]))(x)
    out.backward(g["g"])
    assert relerr(out.data, g["out"]) < 1e-5 and relerr(x.grad, g["dx"]) < 1e-5


@pytest.mark.parametrize("name", ["softmax_last", "softmax_axis1", "softmax_ref"])
def test_softmax_golden(name):
    g = load_golden(name)
    x = dev(g["x"], True)
    out = nn.Softmax(axis=int(g["axis"]))(x)
    out.backward(g["g"])
    assert relerr(out.data, g["out"]) < 1e-5 and relerr(x.grad, g["dx"]) < 1e-5


def test_softmax_masked_attention_rows():
    """-1e9 masked scores (gpt cell 2 l.32): masked entries come out exactly 0, rows sum to 1."""
    rng = np.random.RandomState(0)
    s = rng.randn(4, 8, 64, 64).astype(np.float32)
    mask = np.tril(np.ones((64, 64), bool))
    s = np.where(mask, s, np.float32(-1e9))
    out = nn.Softmax(axis=-1)(dev(s))
    o = host(out.data)
    assert np.all(o[..., ~mask] == 0) and np.allclose(o.sum(-1), 1, atol=1e-5)
    assert relerr(out.data, R.softmax_forward(s, -1)) < 1e-5


@pytest.mark.parametrize("name", ["rmsnorm_2d", "rmsnorm_3d_bias"])
def test_rmsnorm_golden(name):
    g = load_golden(name)
    layer = nn.RMSNorm(g["w"].shape[0], eps=float(g["eps"]), bias="b" in g, device="cuda")
    layer.weight.data = layer.xp.array(g["w"])
    if "b" in g:
        layer.bias.data = layer.xp.array(g["b"])
    x = dev(g["x"], True)
    out = layer(x)
    out.backward(g["g"])
    assert relerr(out.data, g["out"]) < 1e-5 and relerr(x.grad, g["dx"]) < 1e-5
    assert relerr(layer.weight.grad, g["dw"]) < 1e-5
    if "b" in g:
        assert relerr(layer.bias.grad, g["db"]) < 1e-5


def test_rmsnorm_gpt_shape():
    rng = np.random.RandomState(2)
    x, g = rng.randn(4, 64, 512).astype(np.float32), rng.randn(4, 64, 512).astype(np.float32)
    layer = nn.RMSNorm(512, device="cuda")
    w = rng.uniform(0.5, 1.5, 512).astype(np.float32)
    layer.weight.data = layer.xp.array(w)
    xt = dev(x, True)
    out = layer(xt)
    out.backward(g)
    y, xn, std = R.rmsnorm_forward(x, w, None, 1e-6)
    dx, dw, _ = R.rmsnorm_backward(x, w, None, xn, std, g)
    assert relerr(out.data, y) < 1e-5 and relerr(xt.grad, dx) < 1e-5
    assert relerr(layer.weight.grad, R.reverse_broadcast(np.sum(g * xn, axis=0), (512,))) < 1e-5


# ---- Adam / AdamW --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,cls", [("opt_adam", Adam), ("opt_adam_l2", Adam), ("opt_adamw", AdamW), ("opt_adamw_nowd", AdamW)])
def test_optimizer_golden(name, cls):
    g = load_golden(name)
    p0, p1 = dev(g["p0"], True), dev(g["p1"], True)
    opt = cls([p0, p1], lr=float(g["lr"]), betas=tuple(float(b) for b in g["betas"]), eps=float(g["eps"]),
              weight_decay=float(g["wd"]))
    for t in range(3):
        p0.grad = p0.xp.array(g["grads"][t])
        p1.grad = None
        opt.step()
        np.testing.assert_allclose(host(p0.data), g["traj"][t], rtol=2e-6, atol=2e-7)
    assert np.array_equal(host(p1.data), g["p1"])  # grad None => untouched
    np.testing.assert_allclose(host(opt.m[0]), g["m"], rtol=2e-6, atol=1e-8)
    np.testing.assert_allclose(host(opt.v[0]), g["v"], rtol=2e-6, atol=1e-8)


def test_adamw_many_ragged_tensors_vs_oracle():
    """Reference test shapes (tests/test_fusedadamw_cuda.py:53-54) plus odd sizes / unaligned tails."""
    rng = np.random.RandomState(123)
    shapes = [(64, 128), (32, 64), (7,), (1, 1), (8191,), (8193,), (3, 5, 7), (16384 + 3,)]
    ps = [rng.randn(*s).astype(np.float32) for s in shapes]
    gs = [rng.randn(*s).astype(np.float32) for s in shapes]
    ts = [dev(p, True) for p in ps]
    opt = AdamW(ts, lr=1e-3, weight_decay=1e-2)
    ms, vs = [np.zeros_like(p) for p in ps], [np.zeros_like(p) for p in ps]
    for step in range(1, 4):
        for t, g in zip(ts, gs):
            t.grad = t.xp.array(g * step)
        opt.step()
        for i in range(len(ps)):
            ps[i], ms[i], vs[i] = R.adamw_step(ps[i], gs[i] * step, ms[i], vs[i], step, lr=1e-3, weight_decay=1e-2)
    for t, p in zip(ts, ps):
        np.testing.assert_allclose(host(t.data), p, rtol=3e-6, atol=3e-7)


# ---- end-to-end steps ----------------------------------------------------------------------------------------------
def test_mlp_training_step_golden_on_device():
    g = load_golden("mlp_step")
    np.random.seed(0)
    l1, l2 = nn.Linear(20, 16).to("cuda"), nn.Linear(16, 10).to("cuda")
    act, lf = nn.Swish(), nn.CrossEntropyLoss()
    opt = AdamW(l1.parameters() + l2.parameters(), lr=1e-3)
    for t in range(2):
        opt.zero_grad()
        loss = lf(l2(act(l1(neunet.tensor(g["x"], device="cuda")))), neunet.tensor(g["labels"], dtype=np.int32, device="cuda"))
        loss.backward()
        opt.step()
        np.testing.assert_allclose(loss.item(), g["losses"][t], rtol=1e-5)
    assert relerr(l1.weight.data, g["w1_after"]) < 1e-4 and relerr(l2.weight.data, g["w2_after"]) < 1e-4
    assert relerr(l2.weight.grad, g["dw2"]) < 1e-4 and relerr(l1.weight.grad, g["dw1"]) < 1e-4


def test_smoke_entry_point():
    import __graft_entry__ as ge
    ge.smoke()


# ---- fused CrossEntropyLoss ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,C,ignore", [(12, 10, -100), (4096, 10, -100), (256, 15000, 0), (37, 2049, 3)])
def test_fused_cross_entropy_vs_oracle(rows, C, ignore):
    rng = np.random.RandomState(rows + C)
    logits = (rng.randn(rows, C) * 3).astype(np.float32)
    labels = rng.randint(0, C, rows).astype(np.int32)
    if ignore >= 0:
        labels[::5] = ignore
    ref_loss, ref_d = R.cross_entropy(logits, labels, ignore_index=ignore)
    x = dev(logits, True)
    loss = nn.CrossEntropyLoss(ignore_index=ignore)(x, dev(labels, dtype=np.int32))
    assert loss.op == "cross_entropy"
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(ref_loss), rtol=2e-5)
    assert relerr(x.grad, ref_d) < 2e-5


def test_fused_cross_entropy_matches_composed_path_and_reductions():
    rng = np.random.RandomState(4)
    logits = rng.randn(64, 33).astype(np.float32)
    labels = rng.randint(0, 33, 64).astype(np.int32)
    for red in ("mean", "sum"):
        a, b = dev(logits, True), dev(logits, True)
        fused = nn.CrossEntropyLoss(reduction=red)(a, dev(labels, dtype=np.int32))
        comp = nn.NLLLoss(reduction=red)(nn.LogSoftmax(axis=1)(b), dev(labels, dtype=np.int32))
        fused.backward(), comp.backward()
        np.testing.assert_allclose(fused.item(), comp.item(), rtol=2e-5)
        assert relerr(a.grad, b.grad) < 2e-5


def test_graphed_step_tracks_eager_training():
    """A CUDA-graph replay of the public-API training step must produce the same parameter
    trajectory as the eager loop (bias corrections advance on the device, weight bf16 planes are
    re-staged inside the graph, an interleaved eager step must not corrupt later replays)."""
    def build():
        np.random.seed(7)
        l1, l2 = nn.LinearSwish(48, 32).to("cuda"), nn.Linear(32, 10).to("cuda")
        opt = AdamW(l1.parameters() + l2.parameters(), lr=1e-2)
        return l1, l2, opt

    rng = np.random.RandomState(0)
    xs = [rng.randn(64, 48).astype(np.float32) for _ in range(6)]
    ys = [rng.randint(0, 10, 64).astype(np.int32) for _ in range(6)]
    lf = nn.CrossEntropyLoss()

    def make_step(l1, l2, opt):
        def step(xb, yb):
            opt.zero_grad()
            loss = lf(l2(l1(xb)), yb)
            loss.backward()
            opt.step()
            return loss
        return step

    # eager reference trajectory: 2 warm-up-equivalent steps on batch 0, then batches 1..5
    l1, l2, opt = build()
    step = make_step(l1, l2, opt)
    x, y = dev(xs[0]), dev(ys[0], dtype=np.int32)
    seq = [0, 0] + [1, 2, 3, 4, 5]
    eager_losses = []
    for i in seq:
        x.data.copy_(torch.from_numpy(xs[i])), y.data.copy_(torch.from_numpy(ys[i]))
        eager_losses.append(step(x, y).item())
    eager_w = host(l1.weight.data)

    l1, l2, opt = build()
    step = make_step(l1, l2, opt)
    x, y = dev(xs[0]), dev(ys[0], dtype=np.int32)
    g = b200.GraphedStep(step, [x, y], optimizer=opt, warmup=2)   # 2 eager steps on batch 0
    graph_losses = []
    for i in [1, 2, 3]:
        graph_losses.append(g(xs[i], ys[i]).item())
    x.data.copy_(torch.from_numpy(xs[4])), y.data.copy_(torch.from_numpy(ys[4]))
    graph_losses.append(step(x, y).item())                       # interleaved eager step
    graph_losses.append(g(xs[5], ys[5]).item())
    np.testing.assert_allclose(graph_losses, eager_losses[2:], rtol=1e-4)
    assert relerr(l1.weight.data, eager_w) < 1e-4


# ---- Dropout (device RNG; neunet/nn/layers/dropout.py:17-46) ---------------------------------------
def test_dropout_device_mask_statistics_and_backward_consistency():
    """The reference draws mask ~ Bernoulli(1-p)/(1-p) from xp.random (a different generator on every
    back-end), so parity is on the distribution and on the contract: y = x*mask, dx = grad*mask with
    the SAME mask, kept elements scaled by exactly 1/(1-p), eval mode is the identity."""
    b200.manual_seed(123)
    p = 0.1
    x = np.random.RandomState(0).randn(4096, 512).astype(np.float32) + 3.0  # no zeros in x
    layer = nn.Dropout(p)
    xt = dev(x, True)
    y = layer(xt)
    g = np.random.RandomState(1).randn(*x.shape).astype(np.float32) + 5.0
    y.backward(neunet.tensor(g, device="cuda").data)
    yh, dxh = host(y.data), host(xt.grad)
    keep = yh != 0
    assert abs(keep.mean() - (1 - p)) < 3e-3                       # 2M draws: sigma ~ 2e-4
    np.testing.assert_array_equal(keep, dxh != 0)                  # same mask both ways
    scale = np.float32(1.0 / (1.0 - p))
    np.testing.assert_array_equal(yh[keep], (x * scale)[keep])     # bit-exact scaling of kept elements
    np.testing.assert_array_equal(dxh[keep], (g * scale)[keep])
    # rows and columns are not correlated (a counter bug would show as stripes)
    assert abs(keep.mean(axis=0) - (1 - p)).max() < 0.03 and abs(keep.mean(axis=1) - (1 - p)).max() < 0.08
    # a second call draws a different mask; re-seeding reproduces the first
    y2 = host(layer(xt).data)
    assert ((y2 != 0) != keep).mean() > 0.1
    b200.manual_seed(123)
    np.testing.assert_array_equal(host(layer(dev(x)).data), yh)
    layer.eval()
    np.testing.assert_array_equal(host(layer(dev(x)).data), x)


def test_dropout_ragged_size_and_graph_replays_draw_fresh_masks():
    x = np.ones(1003, dtype=np.float32)   # not a multiple of 4: scalar tail
    layer = nn.Dropout(0.5)
    xt = dev(x)
    y = host(layer(xt).data)
    assert set(np.unique(y)) <= {0.0, 2.0} and 0.4 < (y != 0).mean() < 0.6

    def step(t):
        return layer(t)
    g = b200.GraphedStep(step, [xt], warmup=1)
    a = host(g.replay().data).copy()
    b = host(g.replay().data).copy()
    assert 0.3 < (a != b).mean() < 0.7      # independent masks per replay
    assert 0.4 < (a != 0).mean() < 0.6


# ---- fused weight staging: the Adam(W) kernel emits the bf16 planes the next forward GEMM reads --------
@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
@pytest.mark.parametrize("k_in", [72, 70, 513])   # ld == cols, padded pitch + scalar tail path, odd
def test_adamw_emitted_weight_planes_equal_fresh_staging(prec, k_in):
    b200.set_precision(prec)
    np.random.seed(3)
    lin = nn.Linear(k_in, 40).to("cuda")
    opt = AdamW(lin.parameters(), lr=1e-2)
    x = dev(np.random.RandomState(0).randn(96, k_in).astype(np.float32))
    for _ in range(3):
        opt.zero_grad()
        lin(x).mean().backward()
        opt.step()
    st = lin.weight._b200_staged
    assert st.persistent                                  # planes came from the optimizer kernel
    out_emitted = lin(x).data.clone()
    assert lin.weight._b200_staged is st                  # ... and the forward used them as they are
    b200.weights_changed()                                # invalidate: the layer converts the weights itself
    out_fresh = lin(x).data
    assert not lin.weight._b200_staged.persistent
    assert torch.equal(out_emitted, out_fresh)            # bit-identical planes -> bit-identical GEMM
    # and the optimizer takes the cache over again at its next step
    opt.zero_grad()
    lin(x).mean().backward()
    opt.step()
    assert lin.weight._b200_staged.persistent


def test_matmul_backward_reuses_forward_planes_bit_exactly():
    """Tensor.matmul keeps the bf16 planes of both operands for backward; gradients must equal the
    path that converts them again (staged=None)."""
    rng = np.random.RandomState(5)
    q = dev(rng.randn(2, 4, 64, 32).astype(np.float32), True)
    k = dev(rng.randn(2, 4, 64, 32).astype(np.float32), True)
    g = torch.from_numpy(rng.randn(2, 4, 64, 64).astype(np.float32)).cuda()
    out = neunet.matmul(q, k.transpose(0, 1, 3, 2))
    out.backward(g)
    da, db = b200.matmul_backward(q.data, k.data.permute(0, 1, 3, 2), g)
    assert torch.equal(q.grad, da)
    # k's gradient flows back through the transpose: compare in the transposed frame
    assert torch.equal(k.grad, db.permute(0, 1, 3, 2))


@pytest.mark.parametrize("batch,M,K,N", [((20,), 70, 33, 100), ((4, 8), 64, 64, 64), ((3, 6), 17, 128, 5)])
def test_matmul_small_batched_products_fp32_path(batch, M, K, N):
    """Opt-in fp32 CUDA-core kernel for >= 16 batch elements with M, K, N <= 128, straight from the strided
    views (nnb_matmul_set_small_path): results equal NumPy fp32 to round-off, forward and backward, for plain,
    transposed and broadcast operands."""
    prev = b200.lib().nnb_matmul_set_small_path(1)   # opt-in path (off by default)
    try:
        _small_products_case(batch, M, K, N)
    finally:
        b200.lib().nnb_matmul_set_small_path(prev)
    assert b200.lib().nnb_matmul_uses_tensor_cores(int(np.prod(batch)), 1, M, K, N) == (0 if prev else 1)


def _small_products_case(batch, M, K, N):
    assert b200.lib().nnb_matmul_uses_tensor_cores(int(np.prod(batch)), 1, M, K, N) == 0
    rng = np.random.RandomState(11)
    a = rng.randn(*batch, M, K).astype(np.float32)
    bt = rng.randn(*batch, N, K).astype(np.float32)            # used as a transposed view, like k in attention
    g = rng.randn(*batch, M, N).astype(np.float32)
    A, BT = dev(a, True), dev(bt, True)
    nd = len(batch)
    perm = tuple(range(nd)) + (nd + 1, nd)
    out = neunet.matmul(A, BT.transpose(*perm))
    ref = np.matmul(a, np.swapaxes(bt, -1, -2))
    assert relerr(out.data, ref) < 2e-6
    out.backward(torch.from_numpy(g).cuda())
    assert relerr(A.grad, np.matmul(g, bt)) < 2e-6
    assert relerr(BT.grad, np.swapaxes(np.matmul(np.swapaxes(a, -1, -2), g), -1, -2)) < 2e-6
    # broadcast right operand: one [K, N] matrix shared by every batch element
    w = rng.randn(K, N).astype(np.float32)
    Wt = dev(w, True)
    A2 = dev(a, True)
    o2 = neunet.matmul(A2, Wt)
    assert relerr(o2.data, np.matmul(a, w)) < 2e-6
    o2.backward(torch.from_numpy(g).cuda())
    assert relerr(A2.grad, np.matmul(g, w.T)) < 2e-6
    assert relerr(Wt.grad, np.einsum("bmk,bmn->kn", a.reshape(-1, M, K), g.reshape(-1, M, N))) < 1e-5


# ---- Conv2d channels-last fast path (Cin, Cout multiples of 8) at ragged geometry and at DDPM size ----------
@pytest.mark.parametrize("B,Cin,H,W,Cout,k,s,pad4,d", [
    (2, 16, 13, 10, 32, (2, 3), (1, 2), (1, 0, 2, 1), (1, 1)),   # rectangular kernel/stride, asymmetric padding
    (3, 24, 12, 9, 40, (3, 3), (2, 1), (0, 2, 1, 1), (2, 1)),    # mixed stride + dilation, ragged Ho*Wo
    (2, 64, 8, 8, 64, (4, 4), (2, 2), (1, 1, 1, 1), (1, 1)),     # DDPM down-sample, Ho*Wo = 16 (scalar NCHW epilogue)
])
def test_conv2d_fast_path_ragged_geometry_vs_oracle(B, Cin, H, W, Cout, k, s, pad4, d):
    rng = np.random.RandomState(Cin + Cout)
    x = rng.uniform(-1, 1, (B, Cin, H, W)).astype(np.float32)
    w = (rng.uniform(-1, 1, (Cout, Cin) + k) / np.sqrt(Cin * k[0] * k[1])).astype(np.float32)
    b = rng.uniform(-0.5, 0.5, Cout).astype(np.float32)
    ref = R.conv2d_forward(x, w, b, s, pad4, d)
    g = rng.uniform(-1, 1, ref.shape).astype(np.float32)
    dx, dw, db = R.conv2d_backward(x, w, b, g, s, pad4, d)
    xd, wd, gd = [torch.from_numpy(a).cuda() for a in (x, w, g)]
    out = b200.conv2d_forward(xd, wd, torch.from_numpy(b).cuda(), s, pad4, d)
    gdx, gdw, gdb = b200.conv2d_backward(xd, wd, gd, s, pad4, d)
    assert relerr(out, ref) < X3
    assert relerr(gdx, dx) < X3 and relerr(gdw, dw) < X3 and relerr(gdb, db) < 1e-5


@pytest.mark.parametrize("prec,tol", [("bf16x3", X3), ("bf16", BF)])
def test_conv2d_ddpm_size_vs_torch_fp32(prec, tol):
    """A full-size DDPM layer (256 -> 256 3x3 @16x16 and the 4x4 s2 p1 down-sample, B = 16) against torch's
    own fp32 convolution (TF32 off) as an independent second oracle (torch reproduces the
    reference's conv semantics to round-off); the NumPy oracle needs minutes at this size."""
    b200.set_precision(prec)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator(device="cuda").manual_seed(5)
    for k, s, p in ((3, 1, 1), (4, 2, 1)):
        x = (torch.rand(16, 256, 16, 16, device="cuda", generator=gen) * 2 - 1).requires_grad_(True)
        w = ((torch.rand(256, 256, k, k, device="cuda", generator=gen) * 2 - 1) / (256 * k * k) ** 0.5).requires_grad_(True)
        bias = (torch.rand(256, device="cuda", generator=gen) - 0.5).requires_grad_(True)
        ref = torch.nn.functional.conv2d(x, w, bias, stride=s, padding=p)
        g = torch.rand(ref.shape, device="cuda", generator=gen) * 2 - 1
        ref.backward(g)
        out = b200.conv2d_forward(x.detach(), w.detach(), bias.detach(), (s, s), (p, p, p, p), (1, 1))
        dx, dw, db = b200.conv2d_backward(x.detach(), w.detach(), g, (s, s), (p, p, p, p), (1, 1))
        assert relerr(out, ref.detach()) < tol
        assert relerr(dx, x.grad) < tol and relerr(dw, w.grad) < tol and relerr(db, bias.grad) < 1e-4
