"""CPU tests of the host-side mirror (`neunet` package, device="cpu") against the reference-generated
golden vectors, plus API-contract checks the named examples rely on."""
import pickle

import numpy as np
import pytest

import neunet
import neunet.nn as nn
from conftest import load_golden
from neunet import Tensor
from neunet.optim import Adam, AdamW

TOL = dict(rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["linear_ref_small", "linear_3d", "linear_nobias", "linear_n10"])
def test_linear_layer(name):
    g = load_golden(name)
    layer = nn.Linear(g["w"].shape[1], g["w"].shape[0], bias="b" in g)
    layer.weight.data = g["w"].copy()
    if "b" in g:
        layer.bias.data = g["b"].copy()
    x = Tensor(g["x"], requires_grad=True)
    out = layer(x)
    assert out.op == "linear" and out.requires_grad
    np.testing.assert_allclose(out.data, g["out"], **TOL)
    out.backward(g["g"])
    np.testing.assert_allclose(x.grad, g["dx"], **TOL)
    np.testing.assert_allclose(layer.weight.grad, g["dw"], **TOL)
    if "b" in g:
        assert layer.bias.grad.shape == g["db"].shape
        np.testing.assert_allclose(layer.bias.grad, g["db"], **TOL)


def test_linear_init_matches_reference_rng():
    """Same np.random draws in the same order (linear.py:34-43) => identical weights under one seed."""
    g = load_golden("linear_ref_small")
    np.random.seed(42)
    layer = nn.Linear(64, 128)
    assert np.array_equal(layer.weight.data, g["w"]) and np.array_equal(layer.bias.data, g["b"])


@pytest.mark.parametrize("name", ["matmul_2d", "matmul_4d", "matmul_bcast", "matmul_vecmat", "matmul_matvec", "matmul_vecvec"])
def test_matmul(name):
    g = load_golden(name)
    a, b = Tensor(g["a"], requires_grad=True), Tensor(g["b"], requires_grad=True)
    out = a @ b
    np.testing.assert_allclose(out.data, g["out"], **TOL)
    out.backward(g["g"])
    np.testing.assert_allclose(a.grad, g["da"], **TOL)
    np.testing.assert_allclose(b.grad, g["db"], **TOL)


def test_readme_autograd_example():
    g = load_golden("readme_autograd")
    x = neunet.tensor([[7.0, 6.0, 5.0], [4.0, 5.0, 6.0]], requires_grad=True)
    y = neunet.tensor([[1.1, 2.2], [3.3, 4.4], [5.5, 6.6]], requires_grad=True)
    z = neunet.tensor([[2.3, 3.4], [4.5, 5.6]], requires_grad=True)
    out = neunet.tanh(1 / neunet.log(neunet.concatenate([(x @ y) @ z, neunet.exp(x) / neunet.sqrt(x)], axis=1)))
    out.backward(np.ones_like(out.data))
    np.testing.assert_allclose(out.data, g["out"], rtol=1e-6)
    np.testing.assert_allclose(x.grad, g["dx"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(y.grad, g["dy"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(z.grad, g["dz"], rtol=1e-5, atol=1e-8)


CONV = ["conv_3x3_p1", "conv_mnist1", "conv_4x4_s2_p1", "conv_s2_odd", "conv_dil2", "conv_rect", "conv_asym_pad", "conv_wide"]


@pytest.mark.parametrize("name", CONV)
def test_conv2d_layer(name):
    g = load_golden(name)
    cout, cin, kh, kw = g["w"].shape
    p = tuple(int(v) for v in g["pad4"])
    layer = nn.Conv2d(cin, cout, (kh, kw), tuple(int(v) for v in g["stride"]), p, tuple(int(v) for v in g["dil"]), bias="b" in g)
    layer.weight.data = g["w"].copy()
    if "b" in g:
        layer.bias.data = g["b"].copy()
    x = Tensor(g["x"], requires_grad=True)
    out = layer(x)
    np.testing.assert_allclose(out.data, g["out"], **TOL)
    out.backward(g["g"])
    np.testing.assert_allclose(x.grad, g["dx"], **TOL)
    np.testing.assert_allclose(layer.weight.grad, g["dw"], rtol=1e-5, atol=2e-5)
    assert np.array_equal(layer.weight.data, g["w"])  # no in-place dilation left behind
    if "b" in g:
        np.testing.assert_allclose(layer.bias.grad, g["db"], rtol=1e-5, atol=2e-5)


def test_conv_transpose_matches_torch():
    """ConvTranspose2d == F.conv_transpose2d with W' = flip(W).transpose(1,0)."""
    torch = pytest.importorskip("torch")
    np.random.seed(3)
    for (cin, cout, k, s, p) in [(4, 5, 4, 2, 1), (3, 2, 3, 1, 1)]:
        layer = nn.ConvTranspose2d(cin, cout, k, s, p)
        x = np.random.randn(2, cin, 5, 6).astype(np.float32)
        xt = Tensor(x, requires_grad=True)
        out = layer(xt)
        g = np.random.randn(*out.shape).astype(np.float32)
        out.backward(g)
        w = torch.tensor(layer.weight.data, requires_grad=True)
        wt = torch.flip(w, (2, 3)).transpose(0, 1)
        xx = torch.tensor(x, requires_grad=True)
        ref = torch.nn.functional.conv_transpose2d(xx, wt, torch.tensor(layer.bias.data), stride=s, padding=p)
        ref.backward(torch.tensor(g))
        np.testing.assert_allclose(out.data, ref.detach().numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(xt.grad, xx.grad.numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(layer.weight.grad, w.grad.numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", ["swish_b1.0", "swish_b1.5"])
def test_swish(name):
    g = load_golden(name)
    x = Tensor(g["x"], requires_grad=True)
    out = nn.Swish(float(g["beta"]))(x)
    out.backward(g["g"])
    np.testing.assert_allclose(out.data, g["out"], **TOL)
    np.testing.assert_allclose(x.grad, g["dx"], **TOL)


@pytest.mark.parametrize("name", ["softmax_last", "softmax_axis1", "softmax_ref"])
def test_softmax(name):
    g = load_golden(name)
    x = Tensor(g["x"], requires_grad=True)
    out = nn.Softmax(axis=int(g["axis"]))(x)
    out.backward(g["g"])
    np.testing.assert_allclose(out.data, g["out"], **TOL)
    np.testing.assert_allclose(x.grad, g["dx"], **TOL)


@pytest.mark.parametrize("name", ["rmsnorm_2d", "rmsnorm_3d_bias"])
def test_rmsnorm(name):
    g = load_golden(name)
    layer = nn.RMSNorm(g["w"].shape[0], eps=float(g["eps"]), bias="b" in g)
    layer.weight.data = g["w"].copy()
    if "b" in g:
        layer.bias.data = g["b"].copy()
    x = Tensor(g["x"], requires_grad=True)
    out = layer(x)
    out.backward(g["g"])
    np.testing.assert_allclose(out.data, g["out"], **TOL)
    np.testing.assert_allclose(x.grad, g["dx"], **TOL)
    np.testing.assert_allclose(layer.weight.grad, g["dw"], rtol=1e-5, atol=2e-5)
    if "b" in g:
        np.testing.assert_allclose(layer.bias.grad, g["db"], rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("name,cls", [("opt_adam", Adam), ("opt_adam_l2", Adam), ("opt_adamw", AdamW), ("opt_adamw_nowd", AdamW)])
def test_optimizers(name, cls):
    g = load_golden(name)
    p0, p1 = Tensor(g["p0"], requires_grad=True), Tensor(g["p1"], requires_grad=True)
    opt = cls([p0, p1], lr=float(g["lr"]), betas=tuple(float(b) for b in g["betas"]), eps=float(g["eps"]),
              weight_decay=float(g["wd"]))
    for t in range(3):
        p0.grad = g["grads"][t].copy()
        p1.grad = None
        opt.step()
        np.testing.assert_allclose(p0.data, g["traj"][t], rtol=1e-6, atol=1e-7)
    assert np.array_equal(p1.data, g["p1"])


def test_mlp_training_step_bit_parity():
    g = load_golden("mlp_step")
    np.random.seed(0)
    l1, l2 = nn.Linear(20, 16), nn.Linear(16, 10)
    act, lf = nn.Swish(), nn.CrossEntropyLoss()
    opt = AdamW(l1.parameters() + l2.parameters(), lr=1e-3)
    for t in range(2):
        opt.zero_grad()
        loss = lf(l2(act(l1(neunet.tensor(g["x"])))), neunet.tensor(g["labels"], dtype=np.int32))
        loss.backward()
        opt.step()
        np.testing.assert_allclose(loss.data, g["losses"][t], rtol=1e-6)
    np.testing.assert_allclose(l1.weight.data, g["w1_after"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(l2.weight.grad, g["dw2"], rtol=1e-5, atol=1e-7)


def test_fused_linear_swish_equals_unfused():
    np.random.seed(1)
    fused = nn.LinearSwish(12, 7, beta=1.5)
    lin, act = nn.Linear(12, 7), nn.Swish(1.5)
    lin.weight.data, lin.bias.data = fused.weight.data.copy(), fused.bias.data.copy()
    x = np.random.randn(5, 12).astype(np.float32)
    g = np.random.randn(5, 7).astype(np.float32)
    xa, xb = Tensor(x, requires_grad=True), Tensor(x, requires_grad=True)
    oa, ob = fused(xa), act(lin(xb))
    oa.backward(g), ob.backward(g)
    np.testing.assert_allclose(oa.data, ob.data, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(xa.grad, xb.grad, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(fused.weight.grad, lin.weight.grad, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(fused.bias.grad, lin.bias.grad, rtol=1e-5, atol=1e-6)


# ---- API contract ------------------------------------------------------------------------------
def test_tensor_defaults_and_contract():
    assert neunet.tensor([1, 2]).requires_grad is False and Tensor([1, 2]).requires_grad is True
    t = Tensor(np.arange(6, dtype=np.float64).reshape(2, 3))
    assert t.dtype == np.float32 and t.device == "cpu" and t.shape == (2, 3)
    with pytest.raises(ValueError):
        Tensor([1.0], device="tpu")
    with pytest.raises(ValueError):
        Tensor([1.0]).numpy()  # requires_grad tensors refuse .numpy()
    assert neunet.tensor([3.5]).item() == 3.5
    z = neunet.zeros(2, 3)
    z[:, 0::2] = neunet.ones(2, 2)
    assert z.data.sum() == 4
    with pytest.raises(RuntimeError):
        Tensor([1.0])[0] = 2.0  # setitem on requires_grad tensor
    assert (neunet.tensor([1.0, 0.0]) == 0).data.tolist() == [0.0, 1.0]
    w = neunet.where(neunet.tensor([1, 0]) == 0, -1e9, neunet.tensor([5.0, 6.0]))
    assert w.data.tolist() == [5.0, -1e9]


def test_broadcast_and_fancy_index_grads():
    a = Tensor(np.ones((2, 3)), requires_grad=True)
    b = Tensor(np.ones((1, 3)), requires_grad=True)
    (a * b + 2.0).sum().backward()
    assert a.grad.shape == (2, 3) and b.grad.shape == (1, 3) and b.grad[0, 0] == 2
    coef = Tensor(np.arange(10.0), requires_grad=True)
    picked = coef[np.array([1, 3, 3]), None, None]
    assert picked.shape == (3, 1, 1)
    picked.sum().backward()
    assert coef.grad[3] == 1.0  # assignment semantics: duplicates do not accumulate (autograd.py:909-910)


def test_module_reflection_to_and_state_dict(tmp_path):
    class Net(nn.Module):  # no super().__init__(), like the reference examples
        def __init__(self):
            self.body = nn.Sequential(nn.Linear(4, 3), nn.ReLU(), nn.Linear(3, 2))
            self.blocks = nn.ModuleList([nn.RMSNorm(2), nn.Dropout(0.1)])
            self.scale = 2.0

        def forward(self, x):
            x = self.body(x)
            for blk in self.blocks:
                x = blk(x)
            return x * self.scale

    np.random.seed(0)
    net = Net()
    assert len(net.parameters()) == 5
    net = net.to("cpu")
    net.eval()
    assert net.blocks[1].training is False
    out = net(neunet.tensor(np.ones((2, 4), np.float32)))
    assert out.shape == (2, 2)
    sd = net.state_dict()
    assert list(sd)[:2] == ["body.0.weight", "body.0.bias"] and isinstance(sd["body.0.weight"], np.ndarray)
    neunet.save(sd, tmp_path / "m.pkl")
    net2 = Net()
    net2.load_state_dict(neunet.load(tmp_path / "m.pkl"))
    assert np.array_equal(net2.body.modules[0].weight.data, net.body.modules[0].weight.data)
    pickle.dumps(sd)


def test_backward_deep_tape_no_recursion_limit():
    x = Tensor(np.ones(3), requires_grad=True)
    y = x
    for _ in range(3000):
        y = y + 1.0
    y.sum().backward()
    assert np.array_equal(x.grad, np.ones(3))


def test_cuda_without_device_fails_loudly():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    with pytest.raises(RuntimeError):
        neunet.tensor([1.0], device="cuda")


def test_reference_names_outside_the_scope_say_so():
    import neunet.nn as nn
    import neunet.optim as optim
    with pytest.raises(NotImplementedError, match="outside the hot path"):
        nn.LSTM
    with pytest.raises(NotImplementedError, match="outside the hot path"):
        optim.NAdam
    with pytest.raises(AttributeError):
        nn.NoSuchLayer
    assert hasattr(nn, "Linear") and not hasattr(nn, "NoSuchLayer") and not hasattr(nn, "GRU")
