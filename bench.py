#!/usr/bin/env python
"""bench.py -- training samples/s of the neunet dense hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload gpt|mlp|conv|ddpm] [--impl ours|reference]

Workloads = the configurations BASELINE.json names (synthetic data of the named shapes, random-init weights):
  gpt  (default) configs[3]: examples/gpt.ipynb's model verbatim (V=15000, d=512, 8 heads, d_ff=2048, 8 layers,
       dropout 0.1, Adam), T=64, 64 sequences per GPU -- the config the headline metric ("training samples/sec at
       1/2/4/8 B200") is quoted on. Data parallel over N GPUs: batch sharded, gradients all-reduced with NCCL in
       chunks overlapped with backward.
  mlp  configs[1]: MLP 784 -> 128 -> 10, batch 4096 per GPU, Linear fwd/bwd + fused Swish + multi-tensor AdamW, CE.
  conv configs[2]: the README / notebook digits classifier (Conv 1->8, LeakyReLU, MaxPool, Conv 8->16, LeakyReLU,
       MaxPool, BatchNorm2d, Linear 784->10, Sigmoid, MSE, Adam), 28x28, batch 512 per GPU.
  ddpm configs[4]: examples/ddpm.ipynb's SimpleUNet verbatim (3x32x32, down (128,256,512,1024), up (1024,512,256,128),
       61.7 M parameters, Adam 2e-4, MSE), batch 64 per GPU, batch-sharded data parallel.
At N = 1 the default run also measures the other three workloads (short runs) and embeds them in `config.also`.

One "step" = zero_grad, forward, loss, backward, (all-reduce,) optimizer.step on one batch.
  value : whole-job samples/s with the batch already resident in HBM; the step is replayed as a CUDA graph captured
          from the public neunet API (no Python between kernels), timed with CUDA events on the launching stream,
          max over ranks. gpt/ddpm: K steps back to back (the per-step working set is far larger than L2);
          mlp/conv: L2 flushed between the per-step event pairs.
  e2e   : the same metric through the public API with HOST batches: pinned host -> device copy of the step's inputs
          and device -> host read of the loss inside the timed region, every step.
  roofline / cpu_baseline / precision modes: see DESIGN.md ("Measurement").
--impl reference times the reference's CPU implementation of the path (the unmodified reference tree when one is
present next to the repo, else the oracle port, NumPy on the host cores) and prints the same JSON line with
"impl": "reference".
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
for p in (os.path.join(ROOT, "numpy-nn-model_b200"), ROOT, os.path.join(ROOT, "examples")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

_T0 = time.perf_counter()


def hb(phase):
    """Per-phase heartbeat: a hang then leaves its last phase in the log tail (stderr; every rank also appends to
    $BENCH_HB_DIR/hb_r<rank>.log when that directory is set)."""
    rank = int(os.environ.get("RANK", "0"))
    msg = f"[bench r{rank} +{time.perf_counter() - _T0:6.1f}s] {phase}"
    if rank == 0:
        print(msg, file=sys.stderr, flush=True)
    d = os.environ.get("BENCH_HB_DIR")
    if d:
        try:
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, f"hb_r{rank}.log"), "a") as f:
                f.write(msg + "\n")
        except OSError:
            pass


def arm_watchdog(seconds):
    """A hung collective must end the run with stacks on stderr, not sit until the driver's limit: SIGTERM dumps every
    thread's stack; after `seconds` the process dumps them itself and exits non-zero."""
    import faulthandler
    import signal
    try:
        faulthandler.register(signal.SIGTERM, all_threads=True, chain=True)
    except (AttributeError, ValueError):
        pass
    if seconds > 0:
        faulthandler.dump_traceback_later(seconds, exit=True)


# N > 1: persistent GEMM grids leave this many SMs' worth of room for NCCL (0 = use every SM); see DESIGN.md section 6
DEFAULT_DP_GEMM_SMS = 0
MLP = dict(name="mlp_784_128_10", d_in=784, d_hid=128, d_out=10, batch=4096, lr=1e-3)
# examples/gpt.ipynb cells 8, 11: V=15000, d=512, 8 heads, d_ff=2048, 8 layers, Adam(1.5e-4, (0.9, 0.98), 1e-9);
# synthetic tokens in [3, V) (ids 0/1/2 are pad/sos/eos), T=64, 64 sequences per GPU
GPT = dict(name="gpt_small", vocab=15000, d_model=512, heads=8, d_ff=2048, layers=8, seq=64, batch=64,
           lr=1.5e-4, betas=(0.9, 0.98), eps=1e-9, dropout=0.1)
# README.md:227-258 / examples/convolutional_digits_classifier.ipynb cell 2 at batch 512 (BASELINE.json configs[2])
CONV = dict(name="conv_digits_classifier", side=28, channels=(8, 16), batch=512, lr=1e-3)
# examples/ddpm.ipynb cells 5-8 (BASELINE.json configs[4]); 64 images per GPU
DDPM = dict(name="ddpm_unet_32", image=(3, 32, 32), down=(128, 256, 512, 1024), up=(1024, 512, 256, 128), batch=64, lr=2e-4,
            timesteps=300)


def traffic_table():
    """DRAM bytes per launch measured with ncu (profiles/*.json; the newest round wins); None if absent."""
    for name in ("r2_gemm_classes_dram.json", "r1_gemm_classes_dram.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)
            d["file"] = "profiles/" + name
            return d
        except Exception:
            continue
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU path of the same step on the host cores
# --------------------------------------------------------------------------------------------------
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms must use the host cores they can get
    (the reference's NumPy path is multi-threaded through OpenBLAS), so lift the BLAS pool limit at run time."""
    n = len(os.sched_getaffinity(0))
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return n


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([i.get("num_threads", 1) for i in threadpool_info()] or [1])
    except Exception:
        n = 1
    return max(n, 1), len(os.sched_getaffinity(0))


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
    return dict(samples_per_s=B * steps / dt, ms_per_step=dt / steps * 1e3, batch=B, steps=steps, warmup=warmup,
                kind="port", what="oracle/restated.py")


def cpu_gpt_run(cfg, steps, warmup, batch=8, seed=0):
    """Oracle port of the GPT step (oracle/gpt_numpy.py) on `batch` sequences (cost is linear in the batch)."""
    from oracle import gpt_numpy as G
    ps = G.init_params(cfg["vocab"], cfg["d_model"], cfg["d_ff"], cfg["layers"], seed=seed)
    ms, vs = [np.zeros_like(p) for p in ps], [np.zeros_like(p) for p in ps]
    rng = np.random.RandomState(seed)
    data = rng.randint(3, cfg["vocab"], (batch, cfg["seq"] + 1))
    for t in range(1, warmup + 1):
        G.train_step(ps, ms, vs, t, data, cfg["heads"], lr=cfg["lr"], betas=cfg["betas"], eps=cfg["eps"])
    t0 = time.perf_counter()
    for t in range(warmup + 1, warmup + steps + 1):
        G.train_step(ps, ms, vs, t, data, cfg["heads"], lr=cfg["lr"], betas=cfg["betas"], eps=cfg["eps"])
    dt = time.perf_counter() - t0
    return dict(samples_per_s=batch * steps / dt, ms_per_step=dt / steps * 1e3, batch=batch, steps=steps, warmup=warmup,
                kind="port", what="oracle/gpt_numpy.py, dropout off (slightly LESS work than the GPU arm)")


def find_reference_tree():
    """The UNMODIFIED reference package, when a tree travels with the repo (it is not pip-installable offline:
    its build backend is poetry)."""
    for base in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(base, "neunet", "autograd.py")):
            return base
    return None


def cpu_gpt_run_reference_tree(tree, cfg, steps, warmup, batch, seed=0):
    """examples/gpt.ipynb's model on the unmodified reference, device="cpu", under a stub `cupy` module (the package
    imports cupy unconditionally; the CPU path never calls into it). Runs in a child process: both packages are
    called `neunet`."""
    code = r'''
import sys, time, types, json
import numpy as np
stub = types.ModuleType("cupy")
for k in dir(np):
    if not k.startswith("__"):
        setattr(stub, k, getattr(np, k))
stub.ndarray = np.ndarray
sys.modules["cupy"] = stub
sys.path.insert(0, TREE); sys.path.insert(0, EXAMPLES)
import neunet, neunet.nn as nn
from neunet.optim import Adam
import models as M
cfg = CFG
np.random.seed(SEED)
model = M.build_gpt(neunet, nn, vocab=cfg["vocab"], d_model=cfg["d_model"], n_heads=cfg["heads"], d_ff=cfg["d_ff"],
                    n_layers=cfg["layers"], pad_idx=0, device="cpu", dropout=cfg["dropout"])
model.train()
opt = Adam(model.parameters(), lr=cfg["lr"], betas=tuple(cfg["betas"]), eps=cfg["eps"])
rng = np.random.RandomState(SEED)
data = rng.randint(3, cfg["vocab"], (BATCH, cfg["seq"] + 1))
def step():
    opt.zero_grad()
    M.gpt_train_step(neunet, nn, model, opt, data, pad_idx=0)
for _ in range(WARMUP): step()
t0 = time.perf_counter()
for _ in range(STEPS): step()
print(json.dumps({"dt": time.perf_counter() - t0}))
'''
    code = (code.replace("TREE", repr(tree)).replace("EXAMPLES", repr(os.path.join(ROOT, "examples")))
            .replace("CFG", repr(dict(cfg))).replace("SEED", str(seed)).replace("BATCH", str(batch))
            .replace("WARMUP", str(warmup)).replace("STEPS", str(steps)))
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = env["OPENBLAS_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1500, env=env)
    if out.returncode != 0:
        raise RuntimeError(out.stderr[-400:])
    dt = json.loads(out.stdout.strip().splitlines()[-1])["dt"]
    return dict(samples_per_s=batch * steps / dt, ms_per_step=dt / steps * 1e3, batch=batch, steps=steps, warmup=warmup,
                kind="reference", what=f"unmodified reference tree at {tree} (device='cpu', cupy stub), dropout on")


def unet_conv_layers(cfg):
    """(Cin, Cout, H, k, stride, pad, transposed, count) of every conv contraction of the DDPM UNet (ddpm cell 5-7)."""
    layers = [(cfg["image"][0], cfg["down"][0], cfg["image"][1], 3, 1, 1, False)]
    h = cfg["image"][1]
    for i in range(len(cfg["down"]) - 1):
        ci, co = cfg["down"][i], cfg["down"][i + 1]
        layers += [(ci, co, h, 3, 1, 1, False), (co, co, h, 3, 1, 1, False), (co, co, h, 4, 2, 1, False)]
        h //= 2
    for i in range(len(cfg["up"]) - 1):
        ci, co = cfg["up"][i], cfg["up"][i + 1]
        layers += [(2 * ci, co, h, 3, 1, 1, False), (co, co, h, 3, 1, 1, False), (co, co, h, 4, 2, 1, True)]
        h *= 2
    layers.append((cfg["up"][-1], cfg["image"][0], h, 3, 1, 1, True))
    return layers


def unet_conv_flops_per_sample(cfg):
    """Algorithmic conv FLOPs of one training step per sample (fwd + dgrad + wgrad; ConvTranspose2d counted on REAL
    taps, i.e. 1/stride^2 of the zero-stuffed form the reference executes) -- SURVEY.md 8d: ~25.8 GFLOP."""
    total = 0.0
    for ci, co, h, k, s, p, tr in unet_conv_layers(cfg):
        ho = h * s if tr else (h + 2 * p - k) // s + 1
        pos = h * h if tr else ho * ho   # real multiply positions: input pixels for the transposed layers
        total += 2.0 * pos * ci * co * k * k
    return 3.0 * total


def cpu_ddpm_run(cfg, batch=1):
    """Oracle conv contractions (oracle/restated.py, the reference's formulation incl. zero-stuffing for the transposed
    layers) of every conv layer of the UNet, forward + backward, on `batch` images: the conv path is > 95 % of the
    reference's step on CPU, so this is an UPPER bound of the reference's samples/s."""
    from oracle import restated as R
    rng = np.random.RandomState(0)
    t_total = 0.0
    for ci, co, h, k, s, p, tr in unet_conv_layers(cfg):
        w = rng.uniform(-1, 1, (co, ci, k, k)).astype(np.float32) * 0.05
        b = np.zeros(co, np.float32)
        x = rng.uniform(-1, 1, (batch, ci, h, h)).astype(np.float32)
        if tr:  # the reference's ConvTranspose2d: stride-1 correlation over the stuffed, (k-1)-padded, cropped input
            xs = R.set_padding(R.set_stride(x, (s, s)), (k - 1 - p,) * 4)
            args = (xs, w, b, (1, 1), (0, 0, 0, 0), (1, 1))
        else:
            args = (x, w, b, (s, s), (p, p, p, p), (1, 1))
        t0 = time.perf_counter()
        o = R.conv2d_forward(*args)
        g = np.ones_like(o)
        R.conv2d_backward(args[0], w, b, g, *args[3:])
        t_total += time.perf_counter() - t0
    return dict(samples_per_s=batch / t_total, ms_per_step=t_total * 1e3, batch=batch, steps=1, warmup=0, kind="port",
                what="oracle/restated.py conv2d forward+backward of all 17 conv layers (conv contractions only: upper bound)")


def cpu_conv_run(cfg, batch=64):
    """Oracle conv + linear contractions of the digits classifier (conv1, conv2, fc; forward + backward)."""
    from oracle import restated as R
    rng = np.random.RandomState(0)
    c1, c2 = cfg["channels"]
    s = cfg["side"]
    x = rng.uniform(-1, 1, (batch, 1, s, s)).astype(np.float32)
    w1, w2 = rng.randn(c1, 1, 3, 3).astype(np.float32), rng.randn(c2, c1, 3, 3).astype(np.float32)
    wf, bf = R.linear_init(c2 * (s // 4) ** 2, 10)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        o1 = R.conv2d_forward(x, w1, np.zeros(c1, np.float32), (1, 1), (1, 1, 1, 1), (1, 1))
        h1 = o1[:, :, ::2, ::2]
        o2 = R.conv2d_forward(h1, w2, np.zeros(c2, np.float32), (1, 1), (1, 1, 1, 1), (1, 1))
        f = o2[:, :, ::2, ::2].reshape(batch, -1)
        z = R.linear_forward(f, wf, bf)
        R.linear_backward(f, wf, bf, np.ones_like(z))
        R.conv2d_backward(h1, w2, np.zeros(c2, np.float32), np.ones_like(o2), (1, 1), (1, 1, 1, 1), (1, 1))
        R.conv2d_backward(x, w1, np.zeros(c1, np.float32), np.ones_like(o1), (1, 1), (1, 1, 1, 1), (1, 1))
    dt = (time.perf_counter() - t0) / reps
    return dict(samples_per_s=batch / dt, ms_per_step=dt * 1e3, batch=batch, steps=reps, warmup=0, kind="port",
                what="oracle/restated.py conv/linear contractions of the classifier (contractions only: upper bound)")


def workload_label(name):
    if name == "gpt":
        c = GPT
        return c, (f"{c['name']} V={c['vocab']} d={c['d_model']} h={c['heads']} ff={c['d_ff']} L={c['layers']} T={c['seq']}, "
                   f"batch {c['batch']}/GPU (BASELINE.json configs[3], examples/gpt.ipynb dims)")
    if name == "mlp":
        c = MLP
        return c, f"{c['name']} batch {c['batch']}/GPU, Linear fwd/bwd + fused Swish + fused AdamW (BASELINE.json configs[1])"
    if name == "conv":
        c = CONV
        return c, (f"{c['name']} 1x{c['side']}x{c['side']}, Conv 1->{c['channels'][0]}->{c['channels'][1]} + Linear, batch "
                   f"{c['batch']}/GPU (BASELINE.json configs[2]; the notebook/README model is 2 convs + 1 linear)")
    c = DDPM
    return c, (f"{c['name']} {c['image']} down {c['down']} up {c['up']}, batch {c['batch']}/GPU "
               "(BASELINE.json configs[4], examples/ddpm.ipynb cells 5-8)")


def cpu_run(name, cfg, steps, full=False):
    """(result dict, sample description) of the CPU arm for a workload. `full`: the --impl reference arm (larger
    sample, the unmodified reference tree when present); otherwise the bounded cpu_baseline leg of the GPU run."""
    if name == "gpt":
        tree = find_reference_tree() if full else None
        if tree is not None:
            try:
                r = cpu_gpt_run_reference_tree(tree, cfg, steps=min(steps, 3), warmup=1, batch=cfg["batch"])
                return r, f"{r['steps']} steps of {r['batch']} sequences x T={cfg['seq']} ({r['what']})"
            except Exception as e:  # fall through to the port, and say so
                hb(f"unmodified reference failed ({type(e).__name__}: {str(e)[:120]}); using the oracle port")
        r = cpu_gpt_run(cfg, steps=min(steps, 5) if full else 2, warmup=1, batch=cfg["batch"] if full else 8)
        return r, f"{r['steps']} steps of {r['batch']} sequences x T={cfg['seq']} ({r['what']}; cost is linear in the batch)"
    if name == "mlp":
        r = cpu_mlp_run(cfg, steps=max(steps, 20) if full else 100, warmup=3)
        return r, f"{r['steps']} steps of batch {r['batch']} ({r['what']})"
    if name == "conv":
        r = cpu_conv_run(cfg, batch=cfg["batch"] if full else 64)
        return r, f"{r['steps']} passes over {r['batch']} images ({r['what']})"
    r = cpu_ddpm_run(cfg, batch=2 if full else 1)
    return r, f"{r['batch']} image(s) ({r['what']})"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg, label = workload_label(args.workload)
    use_all_host_threads()
    hb("reference arm (CPU)")
    r, sample = cpu_run(args.workload, cfg, max(args.steps, 1), full=True)
    blas, cores = host_threads()
    line = {
        "impl": "reference", "metric": "training samples/sec", "value": r["samples_per_s"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": label + " -- reference CPU path", "global_batch": r["batch"],
                   "requested": {"steps": args.steps, "warmup": args.warmup}},
        "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": blas, "kind": r["kind"],
                         "sample": sample + f", NumPy/OpenBLAS, {cores} cores visible"},
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


def time_graphed(fn, iters, flush=None):
    """Average microseconds of `fn()` (one or more of our launches), GPU-paced: captured once into a CUDA graph and
    replayed `iters` times between CUDA events on the launching stream; `flush` (a > L2 buffer) is rewritten before
    every replay when the operands would otherwise stay L2-resident."""
    import torch
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    torch.cuda.synchronize()
    if flush is None:
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / iters
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot * 1e3 / iters


class MlpWorkload:
    """BASELINE configs[1]: LinearSwish(784,128) -> Linear(128,10) -> CrossEntropy, AdamW."""
    name, flush_l2, dp = "mlp", True, False

    def __init__(self, neunet, nn, optim, rank):
        import torch
        c = self.cfg = MLP
        self.B = c["batch"]
        np.random.seed(0)  # identical init on every rank
        self.l1 = nn.LinearSwish(c["d_in"], c["d_hid"]).to("cuda")
        self.l2 = nn.Linear(c["d_hid"], c["d_out"]).to("cuda")
        self.params = self.l1.parameters() + self.l2.parameters()
        self.opt = optim.AdamW(self.params, lr=c["lr"])
        self.loss_fn = nn.CrossEntropyLoss()
        rng = np.random.RandomState(1000 + rank)
        self.host = [(torch.from_numpy(rng.randn(self.B, c["d_in"]).astype(np.float32)).pin_memory(),
                      torch.from_numpy(rng.randint(0, c["d_out"], self.B).astype(np.int32)).pin_memory()) for _ in range(8)]
        self.inputs = [neunet.tensor(self.host[0][0].numpy(), device="cuda"),
                       neunet.tensor(self.host[0][1].numpy(), dtype=np.int32, device="cuda")]

    def forward_loss(self, x, y):
        return self.loss_fn(self.l2(self.l1(x)), y)

    def roofline(self, pk, b200):
        """Dominant kernel = layer-1 forward GEMM (4096 x 784 x 128 + bias + Swish, Z and O written)."""
        c = self.cfg
        B, K, N = c["batch"], c["d_in"], c["d_hid"]
        us, nl = b200.probe_linear_gemm(B, K, N, form=0, with_bias=True, swish=True, rounds=5)
        alg = B * K * 2 + N * K * 2 + 2 * B * N * 4
        ach = alg / (us * 1e-6) / 1e9
        tr = (traffic_table() or {}).get("mlp_layer1_fwd")
        return {"bound": "hbm", "kernel": "gemm_tcgen05_kernel: Linear-1 forward 4096x784x128 + bias + Swish epilogue",
                "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": (tr["dram_read"] + tr["dram_write"]) if tr else None,
                "us_per_launch": us, "algorithmic_bytes": alg, "flops_per_launch": 2 * B * K * N,
                "peak_source": pk["source"] + " hbm_gbs",
                "how": "graph-paced launches over rotating operand sets > L2, CUDA events (nnb_probe_linear_gemm)"}


class GptWorkload:
    """BASELINE configs[3]: examples/gpt.ipynb model (examples/models.py), Adam, CrossEntropy(ignore pad)."""
    name, flush_l2, dp = "gpt", False, True

    def __init__(self, neunet, nn, optim, rank):
        import models as M
        import torch
        c = self.cfg = GPT
        self.B, self.T = c["batch"], c["seq"]
        np.random.seed(0)
        self.model = M.build_gpt(neunet, nn, vocab=c["vocab"], d_model=c["d_model"], n_heads=c["heads"], d_ff=c["d_ff"],
                                 n_layers=c["layers"], pad_idx=0, device="cuda", dropout=c["dropout"])
        self.model.train()
        self.params = self.model.parameters()
        self.opt = optim.Adam(self.params, lr=c["lr"], betas=c["betas"], eps=c["eps"])
        self.loss_fn = nn.CrossEntropyLoss(ignore_index=0)
        rng = np.random.RandomState(1000 + rank)
        self.host = []
        for _ in range(8):
            tok = rng.randint(3, c["vocab"], (self.B, self.T + 1))
            self.host.append((torch.from_numpy(tok[:, :-1].astype(np.int32)).pin_memory(),
                              torch.from_numpy(tok[:, 1:].reshape(-1).astype(np.int32)).pin_memory()))
        # no padding in synthetic data -> the mask is the causal mask (built once, like GPT.get_sub_mask)
        causal = np.logical_not(np.triu(np.ones((self.T, self.T)), k=1).astype(int)).astype(np.int32)
        self.mask = neunet.tensor(np.broadcast_to(causal, (self.B, self.T, self.T)).copy(), dtype=np.int32, device="cuda")
        self.inputs = [neunet.tensor(self.host[0][0].numpy(), dtype=np.int32, device="cuda"),
                       neunet.tensor(self.host[0][1].numpy(), dtype=np.int32, device="cuda")]

    def forward_loss(self, ids, tgt):
        out, _ = self.model.decoder(ids, self.mask)
        return self.loss_fn(out.reshape(out.shape[0] * out.shape[1], out.shape[2]), tgt)

    def linear_shapes(self):
        """(K, N, launches per step) of every nn.Linear on the path (gpt cell 2-7)."""
        c = self.cfg
        d, ff, V, L = c["d_model"], c["d_ff"], c["vocab"], c["layers"]
        from neunet import autograd
        if autograd.fusion_enabled():  # the q/k/v projections of a block run as ONE GEMM over a shared weight buffer
            return [(d, 3 * d, L, "wq|wk|wv (one GEMM)"), (d, d, L, "attn.fc"), (d, ff, L, "ffn.fc_1"), (ff, d, L, "ffn.fc_2"),
                    (d, V, 1, "fc_out")]
        return [(d, d, 4 * L, "wq/wk/wv/fc"), (d, ff, L, "ffn.fc_1"), (ff, d, L, "ffn.fc_2"), (d, V, 1, "fc_out")]

    def roofline(self, pk, b200, with_ladder=True):
        """Dominant kernel = gemm_tcgen05_kernel. One step launches it in 15 Linear shape x form classes
        (5 layer shapes x fwd/dgrad/wgrad); every class is timed GPU-paced on rotating operand sets > L2 and
        `achieved` = (sum of algorithmic 2MKN over all Linear GEMM launches of a step) / (sum of their times)."""
        c = self.cfg
        M_ = c["batch"] * c["seq"]
        out = {}
        for prec in ("bf16", "bf16x3"):
            tot_fl, tot_us, rows = 0.0, 0.0, []
            with b200.precision(prec):
                for K, N, n, name in self.linear_shapes():
                    for form, fname in enumerate(("fwd", "dgrad", "wgrad")):
                        us, nl = b200.probe_linear_gemm(M_, K, N, form=form, with_bias=(form == 0), rounds=3)
                        fl = 2.0 * M_ * K * N
                        tot_fl += fl * n
                        tot_us += us * n
                        rows.append({"layer": name, "form": fname, "M": M_, "K": K, "N": N, "per_step": n, "us": round(us, 2),
                                     "tflops": round(fl / us * 1e-6, 1), "kernels": nl})
            out[prec] = (tot_fl, tot_us, rows)
        tot_fl, tot_us, rows = out["bf16"]
        ach = tot_fl / tot_us * 1e-6
        ngemm = sum(r["per_step"] for r in rows)
        top = max(rows, key=lambda r: r["us"] * r["per_step"])
        tr = traffic_table() or {}
        roof = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel: all nn.Linear fwd/dgrad/wgrad launches of one step "
                                             f"({ngemm} GEMMs, M={M_})",
                "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops_sustained"],
                "traffic": tr.get("traffic_per_launch_weighted"),
                "traffic_note": "DRAM read+write bytes per launch, weighted over the same GEMM classes (%s; algorithmic operand+output "
                                "bytes per launch: %.0f)" % (tr.get("file"), tr.get("algorithmic_bytes_per_launch_weighted", float("nan"))),
                "us_per_launch": tot_us / ngemm,
                "flops_per_step": tot_fl, "gemm_us_per_step": tot_us,
                "largest_class": f"{top['layer']} {top['form']} {top['M']}x{top['K']}x{top['N']}: {top['us']} us, {top['tflops']} TFLOP/s",
                "classes": rows,
                "bf16x3": {"achieved": out["bf16x3"][0] / out["bf16x3"][1] * 1e-6,
                           "frac": out["bf16x3"][0] / out["bf16x3"][1] * 1e-6 / pk["bf16_tflops_sustained"],
                           "gemm_us_per_step": out["bf16x3"][1],
                           "note": "algorithmic 2MKN credit; the hi/lo split does 3x the MMA work"},
                "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                "how": "graph-paced launches over rotating operand sets > L2, CUDA events (nnb_probe_linear_gemm)"}
        if with_ladder:
            roof["ladder"] = linear_ladder(pk, b200)
        return roof


def linear_ladder(pk, b200):
    """The Linear TFLOP/s ladder of SURVEY.md 8(d) (north-star: >= 70 % of the bf16 tensor peak at batch >= 4096),
    all three forms, both precision modes, GPU-paced like the classes above."""
    rows = []
    for (M_, K, N) in ((4096, 1024, 4096), (8192, 4096, 4096), (8192, 8192, 8192), (16384, 512, 2048)):
        for prec in ("bf16", "bf16x3"):
            if prec == "bf16x3" and M_ * K * N > 4096 * 4096 * 8192:
                continue
            with b200.precision(prec):
                for form, fname in enumerate(("fwd", "dgrad", "wgrad")):
                    us, nl = b200.probe_linear_gemm(M_, K, N, form=form, with_bias=(form == 0), rounds=2, sets=2 if M_ * N >= 8192 * 8192 else None)
                    tf = 2.0 * M_ * K * N / us * 1e-6
                    rows.append({"M": M_, "K": K, "N": N, "form": fname, "precision": prec, "us": round(us, 1), "tflops": round(tf, 1),
                                 "of_sustained": round(tf / pk["bf16_tflops_sustained"], 3), "of_burst": round(tf / pk["bf16_tflops"], 3)})
    return rows


class ConvWorkload:
    """BASELINE configs[2]: the README digits classifier at batch 512 (tiny-channel convs: direct fp32 kernels; HBM /
    latency bound)."""
    name, flush_l2, dp = "conv", True, False

    def __init__(self, neunet, nn, optim, rank):
        import models as M
        import torch
        c = self.cfg = CONV
        self.B = c["batch"]
        np.random.seed(0)
        self.model = M.build_conv_classifier(neunet, nn, device="cuda", side=c["side"], channels=c["channels"])
        self.params = self.model.parameters()
        self.opt = optim.Adam(self.params, lr=c["lr"])
        self.loss_fn = nn.MSELoss()
        rng = np.random.RandomState(1000 + rank)
        self.host = []
        for _ in range(8):
            x = rng.uniform(-1, 1, (self.B, 1, c["side"], c["side"])).astype(np.float32)
            y = np.eye(10, dtype=np.float32)[rng.randint(0, 10, self.B)]
            self.host.append((torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()))
        self.inputs = [neunet.tensor(self.host[0][0].numpy(), device="cuda"), neunet.tensor(self.host[0][1].numpy(), device="cuda")]

    def forward_loss(self, x, y):
        return self.loss_fn(self.model.forward(x), y)

    def roofline(self, pk, b200):
        """Dominant kernels = conv2 (8 -> 16 @14x14) forward and backward, direct fp32 kernels: HBM roofline on the
        algorithmic bytes (x, w, out in fp32 for forward; x, dO read + dx, dw written for backward)."""
        import torch
        c = self.cfg
        B, c1, c2, h = c["batch"], c["channels"][0], c["channels"][1], c["side"] // 2
        x = torch.randn(B, c1, h, h, device="cuda")
        w = torch.randn(c2, c1, 3, 3, device="cuda")
        bias = torch.zeros(c2, device="cuda")
        g = torch.randn(B, c2, h, h, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        us_f = time_graphed(lambda: b200.conv2d_forward(x, w, bias, (1, 1), (1, 1, 1, 1), (1, 1)), 10, flush)
        us_b = time_graphed(lambda: b200.conv2d_backward(x, w, g, (1, 1), (1, 1, 1, 1), (1, 1)), 10, flush)
        bytes_f = 4.0 * (x.numel() + w.numel() + g.numel())
        bytes_b = 4.0 * (2 * x.numel() + 2 * g.numel() + 2 * w.numel())
        ach = (bytes_f + bytes_b) / ((us_f + us_b) * 1e-6) / 1e9
        return {"bound": "hbm", "kernel": "direct_fwd_kernel / direct_dgrad_kernel / direct_wgrad_kernel: conv2 8->16 3x3 @14x14, batch 512",
                "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None,
                "us_forward": us_f, "us_backward": us_b, "algorithmic_bytes": bytes_f + bytes_b,
                "flops": 3 * 2.0 * B * h * h * c1 * c2 * 9, "peak_source": pk["source"] + " hbm_gbs",
                "how": "one call captured in a CUDA graph, L2 flushed (256 MiB write) before every replay, CUDA events"}


class DdpmWorkload:
    """BASELINE configs[4]: examples/ddpm.ipynb SimpleUNet, Algorithm-1 training step (x_t mix, eps-prediction, MSE)."""
    name, flush_l2, dp = "ddpm", False, True

    def __init__(self, neunet, nn, optim, rank):
        import models as M
        import torch
        c = self.cfg = DDPM
        self.B = c["batch"]
        np.random.seed(0)
        self.model = M.build_ddpm_unet(neunet, nn, device="cuda", image_channels=c["image"][0], image_size=c["image"][1],
                                       down_channels=c["down"], up_channels=c["up"])
        self.params = self.model.parameters()
        self.opt = optim.Adam(self.params, lr=c["lr"])
        self.loss_fn = nn.MSELoss()
        betas = np.linspace(1e-4, 0.02, c["timesteps"], dtype=np.float32)
        abar = np.cumprod(1 - betas).astype(np.float32)
        rng = np.random.RandomState(1000 + rank)
        self.host = []
        for _ in range(4):
            x0 = rng.uniform(-1, 1, (self.B,) + c["image"]).astype(np.float32)
            noise = rng.normal(size=x0.shape).astype(np.float32)
            t = rng.randint(1, c["timesteps"], self.B)
            a = np.sqrt(abar[t]).reshape(-1, 1, 1, 1).astype(np.float32)
            b = np.sqrt(1 - abar[t]).reshape(-1, 1, 1, 1).astype(np.float32)
            tf = (t / c["timesteps"]).astype(np.float32).reshape(-1, 1, 1)
            self.host.append(tuple(torch.from_numpy(v).pin_memory() for v in (x0, noise, a, b, tf)))
        self.inputs = [neunet.tensor(h.numpy(), device="cuda") for h in self.host[0]]

    def forward_loss(self, x0, noise, a, b, tf):
        x_t = a * x0 + b * noise          # ddpm cell 4: sqrt(a_bar_t) x0 + sqrt(1 - a_bar_t) eps, per-sample coefficients
        pred = self.model.forward(x_t, tf)
        return self.loss_fn(pred, noise)

    def roofline(self, pk, b200):
        """Dominant kernel = gemm_tcgen05_kernel in its implicit-GEMM conv form. Every distinct conv layer is timed
        as a whole op (layout pass + weight staging + GEMM; forward, and dgrad + wgrad) GPU-paced; `achieved` = the
        algorithmic conv FLOPs of a step (ConvTranspose on real taps) / the sum of those times."""
        import torch
        c = self.cfg
        B = c["batch"]
        tot_us, rows = 0.0, []
        for ci, co, h, k, s, p, tr in unet_conv_layers(c):
            x = torch.randn(B, ci, h, h, device="cuda")
            w = torch.randn(co, ci, k, k, device="cuda") * 0.02
            bias = torch.zeros(co, device="cuda")
            st, pad4, dil = (s, s), (p, p, p, p), (1, 1)
            native_t = tr and b200.conv_transpose2d_supported(x.shape, w.shape, st, pad4, dil, (0, 0))
            if native_t:
                o, planes = b200.conv_transpose2d_forward(x, w, bias, st, pad4, dil, (0, 0))
                g = torch.randn_like(o)
                us_f = time_graphed(lambda: b200.conv_transpose2d_forward(x, w, bias, st, pad4, dil, (0, 0)), 5)
                us_b = time_graphed(lambda: b200.conv_transpose2d_backward(x, w, g, st, pad4, dil, (0, 0), x_planes=planes), 5)
                real = h * h
            elif tr:  # 128 -> 3 output layer: the zero-stuffed formulation over nnb_conv2d (stride 1: only a padded copy)
                xs = torch.nn.functional.pad(x, (k - 1 - p,) * 4)
                o = b200.conv2d_forward(xs, w, bias, (1, 1), (0, 0, 0, 0), dil)
                g = torch.randn_like(o)
                us_f = time_graphed(lambda: b200.conv2d_forward(xs, w, bias, (1, 1), (0, 0, 0, 0), dil), 5)
                us_b = time_graphed(lambda: b200.conv2d_backward(xs, w, g, (1, 1), (0, 0, 0, 0), dil), 5)
                real = h * h
            else:
                o, planes = b200.conv2d_forward(x, w, bias, st, pad4, dil, keep_planes=True)
                g = torch.randn_like(o)
                us_f = time_graphed(lambda: b200.conv2d_forward(x, w, bias, st, pad4, dil), 5)
                us_b = time_graphed(lambda: b200.conv2d_backward(x, w, g, st, pad4, dil, x_planes=planes), 5)
                real = o.shape[2] * o.shape[3]
            fl = 2.0 * B * real * ci * co * k * k
            tot_us += us_f + us_b
            rows.append({"layer": f"{'convT' if tr else 'conv'} {ci}->{co} {k}x{k} s{s} @{h}", "us_fwd": round(us_f, 1),
                         "us_bwd": round(us_b, 1), "tflops_fwd": round(fl / us_f * 1e-6, 1), "tflops_bwd": round(2 * fl / us_b * 1e-6, 1)})
            del x, w, g, o
        fl_step = unet_conv_flops_per_sample(c) * B
        ach = fl_step / tot_us * 1e-6
        return {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (implicit-GEMM Conv2d / ConvTranspose2d forward, dgrad, wgrad incl. their layout and weight-staging passes)",
                "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"],
                "traffic": None, "flops_per_step": fl_step, "conv_us_per_step": tot_us, "layers": rows,
                "peak_source": pk["source"] + " bf16_tflops_sustained",
                "how": "each layer's forward / backward call captured in a CUDA graph and replayed 5x, CUDA events"}


WORKLOADS = {"gpt": GptWorkload, "mlp": MlpWorkload, "conv": ConvWorkload, "ddpm": DdpmWorkload}


def measure(wl, args, world, rank, local, K, W, dist, torch, b200, GradBucket, sample_clocks=True):
    """Warm up, capture, time K steps (device events, max over ranks), then the e2e leg. Returns a dict."""
    B, opt, params = wl.B, wl.opt, wl.params
    bucket = GradBucket(params) if (world > 1 and wl.dp) else None
    if bucket is not None:
        bucket.broadcast_parameters()
        opt.grad_scale = 1.0 / world

    def train_step(*inputs):
        opt.zero_grad()
        loss = wl.forward_loss(*inputs)
        loss.backward()  # with overlap on, each ~32 MB gradient chunk is all-reduced as soon as it is final
        if bucket is not None:
            # the one collective: sum of gradients over NVLink (NCCL). Each chunk's AdamW launch waits for that chunk only,
            # so the optimizer of the early chunks runs underneath the all-reduce of the last one
            bucket.all_reduce_and_step(opt)
        else:
            opt.step()
        return loss

    overlap = bucket is not None and not args.no_overlap
    hb(f"[{wl.name}] eager warm-up")
    for i in range(W):
        train_step(*wl.inputs)
        if overlap and i == 0:
            bucket.overlap_backward()  # live set known after one step: hook the chunked, overlapped all-reduce
    torch.cuda.synchronize()

    def all_ranks_ok(ok):
        """Every decision that changes the collective schedule is taken by ALL ranks together (a rank-local `except`
        that switched one rank to another all-reduce pattern would deadlock the others)."""
        if world == 1:
            return bool(ok)
        f = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        return bool(f.item())

    graphed, graph_err = None, None
    if not args.no_graph:
        for attempt in range(2):
            hb(f"[{wl.name}] graph capture (attempt {attempt}, overlap={overlap})")
            try:
                graphed = b200.GraphedStep(train_step, wl.inputs, optimizer=opt, warmup=2, pdl_retry=(world == 1))
            except Exception as e:  # report, never hide
                graph_err = f"{type(e).__name__}: {e}"[:300]
                graphed = None
                try:
                    torch.cuda.synchronize()
                except Exception:
                    pass
            if all_ranks_ok(graphed is not None):
                break
            if graph_err is None:
                graph_err = "capture failed on another rank"
            graphed = None
            hb(f"capture failed somewhere ({graph_err}); all ranks fall back together")
            if not overlap:
                break
            overlap = False  # retry the capture once, on every rank, with the plain end-of-backward all-reduce
            bucket = GradBucket(params)
            for p_ in params:
                p_._grad_ready = None
                p_._grad_buffer = None

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if wl.flush_l2 else None  # 2x L2

    def one_step():
        return graphed.replay() if graphed is not None else train_step(*wl.inputs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    hb(f"[{wl.name}] " + ("graph warm-up" if graphed is not None else "eager warm-up 2"))
    for _ in range(W):
        if flush is not None:
            flush.zero_()
        one_step()
    barrier()
    hb(f"[{wl.name}] timed region ({K} steps)")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]

    class _NoClocks:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            pass

        def summary(self):
            return None
    with (ClockSampler(local) if sample_clocks else _NoClocks()) as clk:
        barrier()
        wall0 = time.perf_counter()
        for i in range(K):
            if flush is not None:
                flush.zero_()
            ev[i][0].record()
            one_step()
            ev[i][1].record()
        barrier()
        wall = time.perf_counter() - wall0
        if flush is not None:
            dev_s = sum(a.elapsed_time(b) for a, b in ev) / 1e3  # the flush kernels sit between the per-step event pairs
        else:
            dev_s = ev[0][0].elapsed_time(ev[K - 1][1]) / 1e3    # first start -> last end: K whole steps back to back
        t = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s = float(t.item())
        if sample_clocks:
            # keep the GPU busy ~1 s longer so nvidia-smi (100 ms period) sees clocks under this load. The number of
            # extra steps is derived from the AGREED (max over ranks) step time: every step contains collectives, so
            # all ranks must run the same count (round 1 looped on each rank's own clock here and hung at 8 GPUs).
            hb(f"[{wl.name}] clock-sampling tail")
            for _ in range(int(min(2000, max(1, round(1.0 / max(dev_s / K, 1e-5)))))):
                one_step()
            barrier()
    hb(f"[{wl.name}] launch count (one eager step)")
    b200.reset_launch_count()
    train_step(*wl.inputs)
    torch.cuda.synchronize()
    launches_per_step = b200.launch_count()

    hb(f"[{wl.name}] e2e")
    n_host = len(wl.host)

    def e2e_step(i):
        hb_ = wl.host[i % n_host]
        if graphed is not None:
            graphed.load(*hb_)
            loss = graphed.replay()
        else:
            for dst, src in zip(wl.inputs, hb_):
                dst.data.copy_(src, non_blocking=True)
            loss = train_step(*wl.inputs)
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
    h2d = sum(h.numel() * h.element_size() for h in wl.host[0])
    return dict(value=B * world * K / dev_s, dev_s=dev_s, ms_per_step=dev_s / K * 1e3, wall_ms=wall / K * 1e3,
                e2e_value=B * world * K / float(t.item()), h2d=h2d, d2h=4, launches_per_step=launches_per_step,
                graphed=graphed is not None, graph_err=graph_err, overlap=overlap, bucket=bucket is not None,
                clocks=clk.summary(), final_loss=last_loss, handle=graphed)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import neunet
    import neunet.nn as nn
    from neunet import autograd, b200, optim
    from neunet.distributed import GradBucket

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a B200; there is no CPU fallback. Use --impl reference for the CPU arm.")
    torch.cuda.set_device(local)
    arm_watchdog(args.watchdog)
    hb(f"init (world {world})")
    if world > 1:
        import datetime
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # a lost rank / mismatched collective aborts after 2 minutes instead of spinning forever
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    b200.require_device()
    b200.set_precision("bf16")
    b200.manual_seed(0x5EED5EED + 7919 * rank)  # every rank draws its own dropout masks
    gemm_sms = args.gemm_sms if args.gemm_sms >= 0 else DEFAULT_DP_GEMM_SMS
    if world > 1 and gemm_sms > 0:
        b200.lib().nnb_set_sm_budget(int(gemm_sms))
    torch.manual_seed(1234 + rank)
    pk = peaks()
    cfg, label = workload_label(args.workload)
    W, K = max(args.warmup, 3), args.steps
    wl = WORKLOADS[args.workload](neunet, nn, optim, rank)
    m = measure(wl, args, world, rank, local, K, W, dist, torch, b200, GradBucket)

    # ---- the parity-precision mode (bf16x3 meets the 1e-4 north-star tolerance; bf16 is the throughput mode) -----------
    modes = {"bf16": {"value": m["value"], "ms_per_step": m["ms_per_step"],
                      "tolerance_vs_fp32": "2e-3..6e-3 max-norm per contraction (tests/test_gpu_parity.py)"}}
    if not args.no_x3 and args.workload in ("gpt", "ddpm"):
        hb("bf16x3 mode")
        m["handle"] = None
        try:
            b200.set_precision("bf16x3")
            wl3 = WORKLOADS[args.workload](neunet, nn, optim, rank)
            m3 = measure(wl3, args, world, rank, local, max(3, K // 2), 3, dist, torch, b200, GradBucket, sample_clocks=False)
            modes["bf16x3"] = {"value": m3["value"], "ms_per_step": m3["ms_per_step"],
                               "tolerance_vs_fp32": "<= 1e-4 max-norm (north-star bar; measured ~5e-6 per contraction)"}
            del wl3, m3
        except Exception as e:
            modes["bf16x3"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        finally:
            b200.set_precision("bf16")

    hb("roofline probes")
    try:
        roof = wl.roofline(pk, b200)
    except Exception as e:  # a failed probe must not lose the step measurements above
        roof = {"bound": None, "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                "error": f"{type(e).__name__}: {e}"[:300]}

    # ---- the other named configurations (N = 1 only; short runs) ------------------------------------------------------
    also = {}
    if world == 1 and not args.no_also and args.workload == "gpt":
        m["handle"] = None
        for name in ("mlp", "conv", "ddpm"):
            hb(f"also: {name}")
            try:
                _, lab = workload_label(name)
                w2 = WORKLOADS[name](neunet, nn, optim, rank)
                m2 = measure(w2, args, 1, 0, local, 10, 3, dist, torch, b200, GradBucket, sample_clocks=False)
                try:
                    r2 = w2.roofline(pk, b200)
                except Exception as e:
                    r2 = {"error": f"{type(e).__name__}: {e}"[:300]}
                also[name] = {"workload": lab, "value": m2["value"], "unit": "samples/s", "ms_per_step": m2["ms_per_step"],
                              "steps": 10, "e2e": {"value": m2["e2e_value"], "h2d_bytes_per_step": m2["h2d"], "d2h_bytes_per_step": 4},
                              "gpu_launches_per_step": m2["launches_per_step"], "graphed": m2["graphed"],
                              "roofline": {k: r2.get(k) for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "error") if k in r2}}
                del w2, m2
            except Exception as e:
                also[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()

    if rank == 0:
        hb("cpu baseline")
        use_all_host_threads()
        cpu, cpu_sample = cpu_run(args.workload, cfg, K, full=False)
        blas, cores = host_threads()
        if world == 1 and also:
            for name in list(also):
                if "error" in also[name] or name == "ddpm":
                    continue
                try:
                    c2, s2 = cpu_run(name, workload_label(name)[0], 5, full=False)
                    also[name]["cpu_baseline"] = {"value": c2["samples_per_s"], "unit": "samples/s", "kind": c2["kind"], "sample": s2}
                except Exception as e:
                    also[name]["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        line = {
            "metric": "training samples/sec", "value": m["value"], "unit": "samples/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": label, "global_batch": wl.B * world, "parallelism": f"dp{world}",
                       "l2": ("flushed (256 MiB write) between timed steps" if wl.flush_l2 else
                              "not flushed: the per-step working set (weights + optimizer state + activations, > 2 GB) exceeds the 126 MB L2"),
                       "gemm_sms": (gemm_sms if (world > 1 and gemm_sms > 0) else "all"),
                       "grad_allreduce": (None if not m["bucket"] else
                                          ("chunked (32 MB), overlapped with backward" if m["overlap"] else "one flat all-reduce after backward")),
                       "step_execution": "cuda-graph replay of the public-API step" if m["graphed"] else "eager",
                       "fusion": "deferred evaluation on (fused attention / epilogues)" if autograd.fusion_enabled() else "off (strictly eager ops)",
                       "graph_error": m["graph_err"], "precision": "bf16", "precision_modes": modes,
                       "host_wall_ms_per_step": m["wall_ms"], "timed_region_s": m["dev_s"], "also": also},
            "clocks": m["clocks"],
            "e2e": {"value": m["e2e_value"], "unit": "samples/s", "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"]},
            "gpu_launches": m["launches_per_step"] * K,
            "roofline": roof,
            "cpu_baseline": {"value": cpu["samples_per_s"], "unit": "samples/s", "cores": blas, "kind": cpu["kind"],
                             "sample": cpu_sample + f", NumPy/OpenBLAS, {cores} cores visible"},
            "final_loss": m["final_loss"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear-down: no collective after the timing all-reduces. Destroying the NCCL communicator while CUDA
        # graphs that captured its kernels are alive can block forever (seen on the 2-GPU box: both ranks hung
        # in destroy_process_group after printing), so drop the graph, drain the device and leave without it.
        m["handle"] = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gpt", choices=sorted(WORKLOADS),
                    help="gpt = BASELINE.json configs[3], the config the 1/2/4/8-GPU samples/s metric is quoted on (default); "
                         "mlp = configs[1]; conv = configs[2]; ddpm = configs[4]")
    ap.add_argument("--gemm-sms", type=int, default=-1,
                    help="N>1: SMs the persistent GEMM grids are sized for (rest is left to NCCL's CTAs); 0 = all, -1 = default")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: one flat all-reduce after backward instead of overlapped chunks")
    ap.add_argument("--watchdog", type=int, default=900,
                    help="seconds after which a still-running bench dumps all thread stacks and exits non-zero (0 = off)")
    ap.add_argument("--no-graph", action="store_true", help="time the eager public-API step instead of a CUDA-graph replay")
    ap.add_argument("--no-also", action="store_true", help="N=1 gpt: do not also measure mlp / conv / ddpm into config.also")
    ap.add_argument("--no-x3", action="store_true", help="do not also time the bf16x3 (parity-precision) mode")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
