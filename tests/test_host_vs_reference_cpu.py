"""Differential test of the host mirror (device="cpu") against the UNMODIFIED reference package.

Both packages are called ``neunet``, so the reference runs in a subprocess (with the ``cupy`` -> NumPy stub
of oracle/make_golden.py, because it imports CuPy unconditionally) and this process runs the same case
code against this repository's package; outputs and gradients of every case must agree to fp32 round-off.
The cases walk the API surface the named examples use: broadcasting arithmetic with
scalars and reflected operators, reductions, reshape/transpose/indexing (slices, None, Ellipsis, integer
arrays with duplicates), where/comparisons, concatenate, the element-wise functions, activations, losses and
the non-hot-path layers (BatchNorm2d, MaxPool2d, Embedding, RMSNorm, ConvTranspose2d, Dropout in eval).
Skipped when the reference tree is not mounted."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "neunet")), reason="reference tree not mounted")

CASES = r'''
import numpy as np

def run_cases(neunet, nn, device="cpu"):
    def T(data, **kw):
        return neunet.tensor(data, device=device, **kw)

    def host(a):
        return a if isinstance(a, np.ndarray) else a.detach().cpu().numpy()
    rng = np.random.RandomState(7)
    out = {}

    def rec(name, result, *leaves, grad=None):
        g = rng.standard_normal(result.shape).astype(np.float32) if grad is None else grad
        result.backward(g)
        out[name + ".out"] = np.asarray(host(result.data), dtype=np.float32)
        for i, l in enumerate(leaves):
            out[f"{name}.g{i}"] = np.asarray(host(l.grad), dtype=np.float32)

    def leaf(*shape, lo=None):
        a = rng.standard_normal(shape).astype(np.float32)
        if lo is not None:
            a = np.abs(a) + lo
        return T(a, requires_grad=True)

    # 1. broadcasting arithmetic, scalars, reflected operators, unary minus, power
    a, b, c = leaf(3, 4), leaf(1, 4), leaf(3, 1, lo=0.5)
    rec("arith", ((a + b * 2 - 1.5) / (c + 3)) ** 2 + (2 - a) * (1 / (c + 3)) - (-b), a, b, c)
    # 2. reductions
    a = leaf(2, 3, 4)
    rec("mean", a.mean(), a)
    a = leaf(2, 3, 4)
    rec("sum_axis", a.sum(axis=1), a)
    a = leaf(2, 3, 4)
    rec("mean_axis_keep", a.mean(axis=-1, keepdims=True) * a, a)
    # 3. shape ops and indexing
    a = leaf(2, 6, 4)
    rec("reshape_transpose", a.reshape(2, 2, 3, 4).transpose(0, 2, 1, 3).reshape(2, 3, -1), a)
    a = leaf(5, 4)
    rec("slices", a[1:4, ::2] * 3 + a[None, 0, ...][0][:2], a)
    coef = leaf(10)
    x = leaf(4, 2, 3, 3)
    idx = np.array([1, 7, 7, 3])
    rec("fancy_coef", coef[idx, None, None, None] * x, coef, x)
    a = leaf(6, 3)
    rec("fancy_rows_dup", a[np.array([0, 2, 2, 5, 0])] * 2.0, a)          # duplicates: reference assigns, last wins
    # 4. where / comparisons
    a, b = leaf(3, 5), leaf(3, 5)
    mask = T((rng.rand(3, 5) > 0.5).astype(np.int32), dtype=np.int32)
    rec("where_scalar", neunet.where(mask == 0, -1e9, a) * 1e-9 + neunet.where(a > b, a, b), a, b)
    # 5. concatenate
    a, b = leaf(2, 3, 2, 2), leaf(2, 5, 2, 2)
    rec("concat", neunet.concatenate(a, b, axis=1) * 1.5, a, b)
    # 6. element-wise functions
    a = leaf(4, 3, lo=0.2)
    rec("funcs", neunet.exp(a * 0.3) + neunet.sin(a) * neunet.cos(a) + neunet.sqrt(a) + a.log() + a.abs() + a.tanh(), a)
    # 7. activations and losses
    for name, act in [("leaky", nn.LeakyReLU()), ("relu", nn.ReLU()), ("sigmoid", nn.Sigmoid()), ("swish", nn.Swish()),
                      ("softmax", nn.Softmax(axis=-1)), ("tanh", nn.Tanh())]:
        a = leaf(3, 7)
        rec("act_" + name, act(a), a)
    a = leaf(4, 6)
    tgt = T(rng.standard_normal((4, 6)).astype(np.float32))
    rec("mse", nn.MSELoss()(a, tgt), a)
    a = leaf(8, 11)
    labels = T(np.array([1, 0, 3, 10, 0, 5, 5, 2]), dtype=np.int32)
    rec("ce_ignore", nn.CrossEntropyLoss(ignore_index=0)(a, labels), a)
    # 8. non-hot-path layers
    np.random.seed(3)
    bn = nn.BatchNorm2d(3).to(device)
    a = leaf(4, 3, 5, 5)
    rec("batchnorm", bn(a), a, bn.weight, bn.bias)
    out["batchnorm.running_mean"] = np.asarray(host(bn.running_mean.data), dtype=np.float32)
    out["batchnorm.running_var"] = np.asarray(host(bn.running_var.data), dtype=np.float32)
    bn.eval()
    a = leaf(2, 3, 5, 5)
    rec("batchnorm_eval", bn(a), a)
    a = leaf(2, 3, 6, 6)
    rec("maxpool", nn.MaxPool2d(2, 2)(a), a)
    emb = nn.Embedding(9, 4).to(device)
    ids = T(np.array([[1, 2, 2], [8, 1, 0]]), dtype=np.int32)
    rec("embedding", emb(ids) * 2.0, emb.weight)
    rms = nn.RMSNorm(6).to(device)
    a = leaf(2, 5, 6)
    rec("rmsnorm", rms(a), a, rms.weight)
    ct = nn.ConvTranspose2d(3, 4, 4, 2, 1).to(device)
    a = leaf(2, 3, 4, 4)
    rec("convtranspose", ct(a), a, ct.weight, ct.bias)
    cv = nn.Conv2d(3, 5, 3, 2, 1).to(device)
    a = leaf(2, 3, 7, 7)
    rec("conv_s2", cv(a), a, cv.weight, cv.bias)
    drop = nn.Dropout(0.3)
    drop.eval()
    a = leaf(3, 3)
    rec("dropout_eval", drop(a) * 2, a)
    lin = nn.Linear(6, 5).to(device)
    a = leaf(2, 3, 6)
    rec("linear3d", lin(a), a, lin.weight, lin.bias)
    return out
'''

REF_RUNNER = r'''
import sys, types
import numpy as np
stub = types.ModuleType("cupy")
for name in dir(np):
    if not name.startswith("__"):
        setattr(stub, name, getattr(np, name))
stub.ndarray = np.ndarray
sys.modules["cupy"] = stub
sys.path.insert(0, %(ref)r)
import neunet, neunet.nn as nn
import os
assert os.path.realpath(neunet.__file__).startswith(os.path.realpath(%(ref)r))
exec(open(%(cases)r).read())
res = run_cases(neunet, nn)
np.savez(%(out)r, **res)
'''


def test_host_mirror_matches_reference_on_the_example_api_surface():
    with tempfile.TemporaryDirectory() as d:
        cases = os.path.join(d, "cases.py")
        open(cases, "w").write(CASES)
        out = os.path.join(d, "ref.npz")
        runner = os.path.join(d, "runner.py")
        open(runner, "w").write(REF_RUNNER % dict(ref=REF, cases=cases, out=out))
        env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
        r = subprocess.run([sys.executable, runner], capture_output=True, text=True, timeout=300, env=env, cwd=d)
        assert r.returncode == 0, r.stderr[-3000:]
        ref = dict(np.load(out))
    import neunet
    import neunet.nn as nn
    ns = {}
    exec(CASES, ns)
    ours = ns["run_cases"](neunet, nn)
    assert sorted(ours) == sorted(ref)
    bad = []
    for k in sorted(ref):
        a, b = np.asarray(ours[k], np.float64), np.asarray(ref[k], np.float64)
        if a.shape != b.shape:
            bad.append((k, "shape", a.shape, b.shape))
            continue
        err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)
        if not err < 2e-5:
            bad.append((k, float(err)))
    assert not bad, bad
