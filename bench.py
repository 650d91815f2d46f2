#!/usr/bin/env python
"""bench.py -- training samples/s of the neunet dense hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload mlp] [--impl ours|reference]

Workload "mlp" = BASELINE.json configs[1]: 2-layer MLP 784 -> 128 -> 10, batch 4096 per GPU, bf16
tensor-core contractions with fp32 accumulation, Linear fwd/bwd + fused Swish + multi-tensor AdamW,
CrossEntropy loss, synthetic data, random-init weights (layer init = reference's U(+-1/sqrt(in))).

One "step" = zero_grad, forward, loss, backward, optimizer.step on one batch.
  value : whole-job samples/s with the batch already resident in HBM; the step is replayed as a CUDA
          graph captured from the public neunet API (no Python between kernels); each step is timed
          with CUDA events on the launching stream, L2 is flushed between timed steps, max over ranks.
  e2e   : the same metric through the public API with HOST batches: pinned host -> device copy of the
          step's inputs and device -> host read of the loss inside the timed region, every step.
  roofline / cpu_baseline: see DESIGN.md ("Measurement").
--impl reference times the reference's CPU implementation of the path (the oracle port, NumPy on the
host cores) on the same config and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "numpy-nn-model_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

MLP = dict(name="mlp_784_128_10", d_in=784, d_hid=128, d_out=10, batch=4096, lr=1e-3)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the same step on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_mlp_run(cfg, steps, warmup, batch=None, seed=0):
    from oracle import restated as R
    B = batch or cfg["batch"]
    np.random.seed(seed)
    w1, b1 = R.linear_init(cfg["d_in"], cfg["d_hid"])
    w2, b2 = R.linear_init(cfg["d_hid"], cfg["d_out"])
    x = np.random.randn(B, cfg["d_in"]).astype(np.float32)
    y = np.random.randint(0, cfg["d_out"], B).astype(np.int32)
    ps = [w1, b1, w2, b2]
    ms, vs = [np.zeros_like(p) for p in ps], [np.zeros_like(p) for p in ps]

    def step(t):
        z1 = R.linear_forward(x, ps[0], ps[1])
        h = R.swish_forward(z1)
        out = R.linear_forward(h, ps[2], ps[3])
        loss, dout = R.cross_entropy(out, y)
        dh, dw2, db2 = R.linear_backward(h, ps[2], ps[3], dout)
        _, dw1, db1 = R.linear_backward(x, ps[0], ps[1], R.swish_backward(z1, dh))
        for i, g in enumerate((dw1, db1, dw2, db2)):
            ps[i], ms[i], vs[i] = R.adamw_step(ps[i], g, ms[i], vs[i], t, lr=cfg["lr"])
        return loss

    for t in range(1, warmup + 1):
        step(t)
    t0 = time.perf_counter()
    for t in range(warmup + 1, warmup + steps + 1):
        step(t)
    dt = time.perf_counter() - t0
    return dict(samples_per_s=B * steps / dt, ms_per_step=dt / steps * 1e3, batch=B, steps=steps)


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([i.get("num_threads", 1) for i in threadpool_info()] or [1])
    except Exception:
        n = 1
    return max(n, 1), len(os.sched_getaffinity(0))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = MLP
    steps = max(args.steps, 1)
    r = cpu_mlp_run(cfg, steps=steps, warmup=max(args.warmup, 1))
    blas, cores = host_threads()
    line = {
        "impl": "reference", "metric": "training samples/sec", "value": r["samples_per_s"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{cfg['name']} batch {cfg['batch']} (BASELINE.json configs[1]), reference CPU path",
                   "global_batch": cfg["batch"]},
        "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": blas, "kind": "port",
                         "sample": f"{steps} steps of batch {cfg['batch']} (oracle/restated.py, NumPy/OpenBLAS, "
                                   f"{cores} cores visible)"},
        "e2e": {"value": r["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.samples, self._stop, self._t = gpu_index, [], threading.Event(), None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit())
        mx = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import neunet
    import neunet.nn as nn
    from neunet import b200
    from neunet.optim import AdamW

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a B200; there is no CPU fallback. Use --impl reference for the CPU arm.")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b200.require_device()
    b200.set_precision("bf16")
    cfg = MLP
    B = cfg["batch"]
    pk = peaks()

    # ---- model (identical init on every rank), per-rank synthetic data ---------------------------
    np.random.seed(0)
    l1 = nn.LinearSwish(cfg["d_in"], cfg["d_hid"]).to("cuda")
    l2 = nn.Linear(cfg["d_hid"], cfg["d_out"]).to("cuda")
    params = l1.parameters() + l2.parameters()
    opt = AdamW(params, lr=cfg["lr"])
    loss_fn = nn.CrossEntropyLoss()
    rng = np.random.RandomState(1000 + rank)
    n_host = 8  # distinct host batches cycled through (pinned)
    hx = [torch.from_numpy(rng.randn(B, cfg["d_in"]).astype(np.float32)).pin_memory() for _ in range(n_host)]
    hy = [torch.from_numpy(rng.randint(0, cfg["d_out"], B).astype(np.int32)).pin_memory() for _ in range(n_host)]
    x = neunet.tensor(hx[0].numpy(), device="cuda")
    y = neunet.tensor(hy[0].numpy(), dtype=np.int32, device="cuda")

    from neunet.distributed import GradBucket
    bucket = GradBucket(params) if world > 1 else None
    if bucket is not None:
        bucket.broadcast_parameters()
        opt.grad_scale = 1.0 / world

    def train_step(xb, yb):
        opt.zero_grad()
        loss = loss_fn(l2(l1(xb)), yb)
        loss.backward()
        if bucket is not None:
            bucket.all_reduce()  # the one collective: sum of gradients over NVLink (NCCL)
        opt.step()
        return loss

    # ---- warm-up (eager) then capture the whole step as a CUDA graph ------------------------------
    for _ in range(max(args.warmup, 3)):
        train_step(x, y)
    torch.cuda.synchronize()
    graphed, graph_err = None, None
    if not args.no_graph:
        try:
            graphed = b200.GraphedStep(train_step, [x, y], optimizer=opt, warmup=2)
        except Exception as e:  # report, never hide
            graph_err = f"{type(e).__name__}: {e}"
            graphed = None
            torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # 2x L2: evicts the working set

    def one_step():
        return graphed.replay() if graphed is not None else train_step(x, y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        one_step()
    barrier()

    # ---- timed: K steps, each bracketed by CUDA events, L2 flushed in between ---------------------
    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    b200.reset_launch_count()
    launches_before = b200.launch_count()
    with ClockSampler(local) as clk:
        barrier()
        wall0 = time.perf_counter()
        for i in range(K):
            flush.zero_()
            ev[i][0].record()
            one_step()
            ev[i][1].record()
        barrier()
        wall = time.perf_counter() - wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_s = sum(step_ms) / 1e3
    # launches per step: count one eager step (graph replays do not pass through the counter)
    b200.reset_launch_count()
    train_step(x, y)
    torch.cuda.synchronize()
    launches_per_step = b200.launch_count()
    t = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s = float(t.item())
    value = B * world * K / dev_s

    # ---- e2e: host batches through the public API, H2D + D2H inside the timed region --------------
    def e2e_step(i):
        hb, yb = hx[i % n_host], hy[i % n_host]
        if graphed is not None:
            graphed.load(hb, yb)
            loss = graphed.replay()
        else:
            x.data.copy_(hb, non_blocking=True)
            y.data.copy_(yb, non_blocking=True)
            loss = train_step(x, y)
        return loss.item()  # device -> host read of the step's result

    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        last_loss = e2e_step(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * world * K / float(t.item())
    h2d = hx[0].numel() * 4 + hy[0].numel() * 4
    d2h = 4

    # ---- roofline of the dominant kernel: layer-1 forward GEMM (tcgen05), timed alone ---------------
    roof = roofline_probe(cfg, pk)

    line = None
    if rank == 0:
        cpu = cpu_mlp_run(cfg, steps=150, warmup=3)
        blas, cores = host_threads()
        line = {
            "metric": "training samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_s / K * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{cfg['name']} batch {B}/GPU, Linear fwd/bwd + fused Swish + fused AdamW "
                                   "(BASELINE.json configs[1])",
                       "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2": "flushed (256 MiB write) between timed steps",
                       "step_execution": "cuda-graph replay of the public-API step" if graphed is not None else "eager",
                       "graph_error": graph_err, "precision": b200.get_precision(),
                       "host_wall_ms_per_step": wall / K * 1e3},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_step * K,
            "roofline": roof,
            "cpu_baseline": {"value": cpu["samples_per_s"], "unit": "samples/s", "cores": blas, "kind": "port",
                             "sample": f"{cpu['steps']} steps of batch {cpu['batch']} (oracle/restated.py, "
                                       f"NumPy/OpenBLAS, {cores} cores visible)"},
            "final_loss": last_loss,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def roofline_probe(cfg, pk):
    """Times the dominant kernel of the step alone: the layer-1 forward GEMM
    (4096 x 784 x 128, bf16 operands already staged, fp32 Z and Swish(Z) written by the epilogue).
    Algorithmic bytes = X bf16 + W bf16 + Z fp32 + O fp32; bound = HBM."""
    import torch

    from neunet import b200
    B, K, N = cfg["batch"], cfg["d_in"], cfg["d_hid"]
    x = torch.randn(B, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    bias = torch.zeros(1, N, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    iters = 20
    times = []
    for it in range(iters + 3):
        flush.zero_()
        t = b200.time_linear_forward_gemm(x, w, bias, act=b200.ACT_SWISH)
        if it >= 3:
            times.append(t)
    ms = float(np.median(times))
    alg_bytes = B * K * 2 + N * K * 2 + 2 * B * N * 4
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "gemm_tcgen05_kernel (Linear-1 forward + bias + Swish epilogue)",
            "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
            "traffic": None, "us_per_launch": ms * 1e3, "algorithmic_bytes": alg_bytes, "peak_source": pk["source"],
            "flops_per_launch": 2 * B * K * N}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mlp", choices=["mlp"])
    ap.add_argument("--no-graph", action="store_true", help="time the eager public-API step instead of a CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
