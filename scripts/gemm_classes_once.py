#!/usr/bin/env python
"""Launch every nn.Linear GEMM class of the GPT-small step (5 layer shapes x fwd/dgrad/wgrad, M = B*T; q/k/v as one GEMM)
a few times through nnb_probe_linear_gemm, for an ncu pass that collects DRAM bytes per launch:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:gemm_tcgen05 --csv --log-file gpurun_out/gemm_classes_dram.csv python scripts/gemm_classes_once.py
The classes run in a fixed order (printed), 3 kernel executions each (1 warm-up + 2 graph replays of 1 launch)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from neunet import b200  # noqa: E402

torch.cuda.set_device(0)
b200.require_device()
b200.set_precision("bf16")
c = bench.GPT
M = c["batch"] * c["seq"]
d, ff, V, L = c["d_model"], c["d_ff"], c["vocab"], c["layers"]
for K, N, n, name in [(d, 3 * d, L, "wq|wk|wv (one GEMM)"), (d, d, L, "attn.fc"), (d, ff, L, "ffn.fc_1"), (ff, d, L, "ffn.fc_2"),
                      (d, V, 1, "fc_out")]:  # = bench.GptWorkload.linear_shapes() with fusion on
    for form, fname in enumerate(("fwd", "dgrad", "wgrad")):
        us, nl = b200.probe_linear_gemm(M, K, N, form=form, with_bias=(form == 0), rounds=1, sets=1)
        print(f"class {name} {fname} M={M} K={K} N={N} per_step={n} kernels_per_gemm={nl} us={us:.2f}", flush=True)
