"""Host-side logic of the deferred-evaluation / fusion layer (neunet/autograd.py) WITHOUT a GPU: the native
entry points are replaced by torch-CPU formulas (tests/mock_b200.py), so these tests pin the pattern matching,
tape wiring and gradient routing of the fused path against (a) the strictly eager path and (b) the NumPy
device="cpu" path of the same model. The kernels themselves are validated on the B200 (tests/test_fused_gpu.py)."""
import numpy as np
import pytest

import mock_b200

torch = pytest.importorskip("torch")


def _gpt(device, dropout):
    import models as M
    import neunet
    import neunet.nn as nn
    np.random.seed(3)
    model = M.build_gpt(neunet, nn, vocab=40, d_model=32, n_heads=4, d_ff=64, n_layers=2, pad_idx=0, device=device,
                        dropout=dropout)
    return neunet, nn, model


def _run(device, dropout, fuse, train=True):
    import models as M
    from neunet import autograd, b200
    neunet, nn, model = _gpt(device, dropout)
    if train:
        model.train()
    else:
        model.eval()
    rng = np.random.RandomState(0)
    batch = rng.randint(1, 40, (3, 9))
    batch[0, -2:] = 0  # some padding: exercises the key-padding part of the mask and ignore_index
    prev = autograd.set_fusion(fuse)
    try:
        if device == "cuda":
            b200.manual_seed(11)
            b200._mock_reset_rng()
            mock_b200.calls.clear()
        loss_fn = nn.CrossEntropyLoss(ignore_index=0)
        out, attn = model.forward(batch[:, :-1])
        logits = out.reshape(out.shape[0] * out.shape[1], out.shape[2])
        loss = loss_fn(logits, neunet.tensor(batch[:, 1:].flatten(), device=device, dtype=neunet.int32))
        loss.backward()
        grads = [None if p.grad is None else mock_b200.to_np(p.grad).copy() for p in model.parameters()]
        return float(mock_b200.to_np(loss.data)), mock_b200.to_np(out.data).copy(), mock_b200.to_np(attn.data).copy(), grads, \
            list(mock_b200.calls)
    finally:
        autograd.set_fusion(prev)


def _close(a, b, tol):
    assert (a is None) == (b is None)
    if a is not None:
        assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-3), np.abs(a - b).max()  # floor: wk.bias gradients are exactly 0 in theory


def test_fused_equals_eager_equals_numpy_without_dropout():
    ref = _run("cpu", 0.0, False, train=False)
    with mock_b200.mocked():
        eager = _run("cuda", 0.0, False, train=False)
        fused = _run("cuda", 0.0, True, train=False)
    for got in (eager, fused):
        assert abs(got[0] - ref[0]) < 1e-5
        _close(got[1], ref[1], 1e-5)
        _close(got[2], ref[2], 1e-5)
        for g, r in zip(got[3], ref[3]):
            _close(g, r, 2e-5)
    calls = fused[4]
    assert "attention_forward" in calls and "attention_backward" in calls
    assert "matmul" not in calls and "softmax_forward" not in calls  # the whole chain went into one kernel
    assert "swish_forward" not in calls                               # Swish became the fc_1 GEMM epilogue


def test_fused_equals_eager_with_dropout_masks():
    with mock_b200.mocked():
        eager = _run("cuda", 0.1, False)
        fused = _run("cuda", 0.1, True)
    assert abs(eager[0] - fused[0]) < 1e-5
    _close(fused[1], eager[1], 1e-5)
    _close(fused[2], eager[2], 1e-5)
    for g, r in zip(fused[3], eager[3]):
        _close(g, r, 2e-5)
    calls = fused[4]
    from neunet.nn.layers import linear as _lin
    assert _lin.fusion_stats["group_calls"] >= 2 and _lin.fusion_stats["backward_zero_copy"] >= 2  # q/k/v: ONE GEMM each way
    assert calls.count("linear_forward") + calls.count("linear_forward_staged") == 9   # 2 x (qkv, fc, fc_1, fc_2) + fc_out
    assert "rmsnorm_forward_fused" in calls      # x + dropout(a) -> RMSNorm in one kernel
    assert "rmsnorm_backward_acc" in calls       # residual gradient accumulated by the norm's backward
    assert "linear_forward_staged" in calls      # Linear consumed ready-made operand planes
    assert "softmax_forward" not in calls and "matmul" not in calls
    # fc_2(dropout(swish(fc_1 x))): one pass over the pre-activation, the Swish output is never materialised
    assert calls.count("swish_dropout") == 2 and "swish_forward" not in calls
    # the backward of every nn.Dropout that sits on an nn.Linear result is applied by that Linear's staging pass:
    # 2 layers x (attention fc, fc_1 via swish, fc_2); only the embedding dropout keeps a stand-alone mask pass
    assert calls.count("linear_backward_dropped") == 6 and calls.count("dropout") == 2  # embedding dropout: forward + backward


def test_swish_output_absorbed_by_dropout_still_materialises_when_read():
    with mock_b200.mocked():
        import neunet
        import neunet.nn as nn
        from neunet import autograd, b200
        np.random.seed(0)
        lin, act, drop = nn.Linear(8, 16).to("cuda"), nn.Swish(), nn.Dropout(0.25)
        x = neunet.tensor(np.random.randn(4, 8), device="cuda", requires_grad=True)
        out = {}
        for fuse in (True, False):
            prev = autograd.set_fusion(fuse)
            try:
                b200.manual_seed(5)
                b200._mock_reset_rng()
                mock_b200.calls.clear()
                x.grad = None
                lin.weight.grad = None
                s = act(lin(x))
                d = drop(s)
                dv = mock_b200.to_np(d.data).copy()
                if fuse:
                    assert "swish_dropout" in mock_b200.calls and "swish_forward" not in mock_b200.calls and s.pending
                sv = mock_b200.to_np(s.data).copy()  # read AFTER the dropout absorbed it
                (d * 2.0).sum().backward()
                out[fuse] = (dv, sv, mock_b200.to_np(x.grad).copy(), mock_b200.to_np(lin.weight.grad).copy())
            finally:
                autograd.set_fusion(prev)
        for a, b in zip(out[True], out[False]):
            np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-6)


def test_pending_results_materialise_when_read():
    """Anything that reads a pending result gets the same values as the eager path."""
    with mock_b200.mocked():
        import neunet
        import neunet.nn as nn
        from neunet import autograd
        np.random.seed(0)
        lin = nn.Linear(8, 8).to("cuda")
        x = neunet.tensor(np.random.randn(2, 4, 8), device="cuda", requires_grad=True)
        prev = autograd.set_fusion(True)
        try:
            y = lin(x)
            assert isinstance(y, autograd._Deferred) and y.pending and y.shape == (2, 4, 8)
            v = y.reshape(2, 4, 2, 4).transpose(0, 2, 1, 3)
            assert v.pending and v.shape == (2, 2, 4, 4)
            s = neunet.matmul(v, v.transpose(0, 1, 3, 2)) / 2.0
            assert s.pending
            got = mock_b200.to_np(s.data)
            assert not y.pending
        finally:
            autograd.set_fusion(prev)
        autograd.set_fusion(False)
        try:
            y2 = lin(x)
            v2 = y2.reshape(2, 4, 2, 4).transpose(0, 2, 1, 3)
            want = mock_b200.to_np((neunet.matmul(v2, v2.transpose(0, 1, 3, 2)) / 2.0).data)
        finally:
            autograd.set_fusion(True)
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)


def test_api_surface_and_ddpm_unet_on_the_mocked_device_match_the_reference_goldens():
    """The whole example API surface (72 reference-generated arrays) and the DDPM UNet step through the device="cuda"
    host logic (deferred LeakyReLU -> BatchNorm2d, eval-mode BatchNorm backward quirk, ...), kernels mocked."""
    from conftest import load_golden
    from test_host_vs_reference_cpu import CASES
    import test_models
    with mock_b200.mocked():
        import neunet
        import neunet.nn as nn
        ref = load_golden("api_surface")
        ns = {}
        exec(CASES, ns)
        ours = ns["run_cases"](neunet, nn, device="cuda")
        assert sorted(ours) == sorted(ref)
        for k in sorted(ref):
            a, b = np.asarray(mock_b200.to_np(ours[k]), np.float64), np.asarray(ref[k], np.float64)
            assert a.shape == b.shape, k
            assert np.abs(a - b).max() <= 1e-4 * max(np.abs(b).max(), 1e-12), k
        mock_b200.calls.clear()

        import models as M
        g = load_golden("model_ddpm_unet")
        np.random.seed(0)
        du = M.build_ddpm_unet(neunet, nn, device="cuda", image_size=8, down_channels=(8, 16, 32), up_channels=(32, 16, 8))
        test_models._load_params(du, g)
        t = lambda a: neunet.tensor(a, device="cuda")  # noqa: E731
        x_t = t(g["a"]) * t(g["x0"]) + t(g["b"]) * t(g["noise"])
        pred = du.forward(neunet.tensor(x_t, requires_grad=False, device="cuda"), g["t_frac"])
        loss = nn.MSELoss()(pred, t(g["noise"]))
        loss.backward()
        assert abs(float(mock_b200.to_np(loss.data)) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
        assert test_models._rel(pred.data, g["out"]) < 5e-5
        test_models._check_grads_and_params(du, g, 5e-4, after=False)
        assert "bn_forward_lrelu" in mock_b200.calls  # LeakyReLU was absorbed by the BatchNorm kernels


@pytest.mark.parametrize("order", [0, 1])
def test_parked_dropout_gradient_with_a_second_consumer_of_the_linear_result(order):
    """The backward of an nn.Dropout on an nn.Linear result is left to that Linear's dO staging (autograd._MaskedGrad). If the
    Linear result has ANOTHER consumer, its contribution must be added to the MASKED gradient, whichever arrives first."""
    with mock_b200.mocked():
        import neunet
        import neunet.nn as nn
        from neunet import autograd, b200
        np.random.seed(1)
        lin, drop = nn.Linear(8, 16).to("cuda"), nn.Dropout(0.5)
        xv = np.random.randn(6, 8).astype(np.float32)
        res = {}
        for fuse in (True, False):
            prev = autograd.set_fusion(fuse)
            try:
                b200.manual_seed(3)
                b200._mock_reset_rng()
                mock_b200.calls.clear()
                lin.weight.grad = lin.bias.grad = None
                x = neunet.tensor(xv, device="cuda", requires_grad=True)
                h = lin(x)
                a, b = drop(h) * 3.0, h * h
                loss = (a.sum() + b.sum()) if order == 0 else (b.sum() + a.sum())
                loss.backward()
                res[fuse] = [mock_b200.to_np(t).copy() for t in (x.grad, lin.weight.grad, lin.bias.grad, h.grad)]
            finally:
                autograd.set_fusion(prev)
        for got, want in zip(res[True], res[False]):
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
