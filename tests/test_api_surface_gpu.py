"""The example API surface on device="cuda" against the reference's own results.

tests/golden/api_surface.npz was produced by the UNMODIFIED reference (oracle/make_api_surface_golden.py) for
the case code in tests/test_host_vs_reference_cpu.py; here the same cases run on the B200 back-end
(array facade + sm_100a kernels for Linear / Conv2d / ConvTranspose2d / RMSNorm / Softmax / Swish /
CrossEntropy, BF16X3 contractions) and every output and gradient must match: 1e-4 for anything that passed
through a tensor-core contraction, 2e-5 otherwise."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import load_golden  # noqa: E402
from test_host_vs_reference_cpu import CASES  # noqa: E402


def test_example_api_surface_on_device_matches_reference():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import neunet
    import neunet.nn as nn
    from neunet import b200
    b200.require_device()
    b200.set_precision("bf16x3")
    ref = load_golden("api_surface")
    ns = {}
    exec(CASES, ns)
    ours = ns["run_cases"](neunet, nn, device="cuda")
    torch.cuda.synchronize()
    assert sorted(ours) == sorted(ref)
    bad = []
    for k in sorted(ref):
        a, b = np.asarray(ours[k], np.float64), np.asarray(ref[k], np.float64)
        if a.shape != b.shape:
            bad.append((k, "shape", a.shape, b.shape))
            continue
        tol = 1e-4 if k.split(".")[0] in ("linear3d", "conv_s2", "convtranspose") else 2e-5
        err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)
        if not err < tol:
            bad.append((k, float(err)))
    assert not bad, bad
