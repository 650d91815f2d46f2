#!/usr/bin/env python
"""bench.py -- training samples/s of the neunet dense hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload gpt|mlp] [--impl ours|reference]

Workload "gpt" (default) = BASELINE.json configs[3], the config the headline metric ("training
samples/sec at 1/2/4/8 B200") is quoted on and that fits one GPU: examples/gpt.ipynb's model verbatim
(V=15000, d=512, 8 heads, d_ff=2048, 8 layers, dropout 0.1, Adam), T=64, 64 sequences per GPU, bf16
tensor-core contractions with fp32 accumulation, synthetic tokens, random-init weights. Data parallel
over N GPUs: batch sharded, gradients all-reduced with NCCL in chunks overlapped with backward.
Workload "mlp" = configs[1]: 2-layer MLP 784 -> 128 -> 10, batch 4096 per GPU, Linear fwd/bwd + fused
Swish + multi-tensor AdamW, CrossEntropy loss (layer init = reference's U(+-1/sqrt(in))).

One "step" = zero_grad, forward, loss, backward, (all-reduce,) optimizer.step on one batch.
  value : whole-job samples/s with the batch already resident in HBM; the step is replayed as a CUDA
          graph captured from the public neunet API (no Python between kernels), timed with CUDA
          events on the launching stream, max over ranks. gpt: K steps back to back (the per-step
          working set is far larger than L2); mlp: L2 flushed between the per-step event pairs.
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
for _p in (os.path.join(ROOT, "examples"),):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def ncu_traffic():
    """DRAM bytes per launch measured once with ncu (profiles/r1_gemm_classes_dram.json); None if absent."""
    path = os.path.join(ROOT, "profiles", "r1_gemm_classes_dram.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
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


def cpu_gpt_run(cfg, steps, warmup, batch=8, seed=0):
    """Oracle port of the GPT step (oracle/gpt_numpy.py) on a bounded sample: `batch` sequences."""
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
    return dict(samples_per_s=batch * steps / dt, ms_per_step=dt / steps * 1e3, batch=batch, steps=steps)


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


def workload_label(args):
    if args.workload == "gpt":
        c = GPT
        return c, (f"{c['name']} V={c['vocab']} d={c['d_model']} h={c['heads']} ff={c['d_ff']} L={c['layers']} T={c['seq']}, "
                   f"batch {c['batch']}/GPU (BASELINE.json configs[3], examples/gpt.ipynb dims)")
    c = MLP
    return c, f"{c['name']} batch {c['batch']}/GPU, Linear fwd/bwd + fused Swish + fused AdamW (BASELINE.json configs[1])"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg, label = workload_label(args)
    use_all_host_threads()
    steps = max(args.steps, 1)
    if args.workload == "gpt":
        steps = min(steps, 5)
        r = cpu_gpt_run(cfg, steps=steps, warmup=1, batch=8)
        sample = f"{steps} steps of 8 sequences x T={cfg['seq']} (oracle/gpt_numpy.py, dropout off -- the CPU arm does slightly LESS work than the GPU arm; cost is linear in batch)"
    else:
        r = cpu_mlp_run(cfg, steps=steps, warmup=max(args.warmup, 1))
        sample = f"{steps} steps of batch {cfg['batch']} (oracle/restated.py)"
    blas, cores = host_threads()
    line = {
        "impl": "reference", "metric": "training samples/sec", "value": r["samples_per_s"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": label + " -- reference CPU path (NumPy/OpenBLAS oracle port)", "global_batch": r["batch"]},
        "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": blas, "kind": "port",
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


class MlpWorkload:
    """BASELINE configs[1]: LinearSwish(784,128) -> Linear(128,10) -> CrossEntropy, AdamW."""

    def __init__(self, neunet, nn, optim, rank):
        c = self.cfg = MLP
        self.B = c["batch"]
        np.random.seed(0)  # identical init on every rank
        self.l1 = nn.LinearSwish(c["d_in"], c["d_hid"]).to("cuda")
        self.l2 = nn.Linear(c["d_hid"], c["d_out"]).to("cuda")
        self.params = self.l1.parameters() + self.l2.parameters()
        self.opt = optim.AdamW(self.params, lr=c["lr"])
        self.loss_fn = nn.CrossEntropyLoss()
        import torch
        rng = np.random.RandomState(1000 + rank)
        self.host = [(torch.from_numpy(rng.randn(self.B, c["d_in"]).astype(np.float32)).pin_memory(),
                      torch.from_numpy(rng.randint(0, c["d_out"], self.B).astype(np.int32)).pin_memory()) for _ in range(8)]
        self.inputs = [neunet.tensor(self.host[0][0].numpy(), device="cuda"),
                       neunet.tensor(self.host[0][1].numpy(), dtype=np.int32, device="cuda")]

    def forward_loss(self, x, y):
        return self.loss_fn(self.l2(self.l1(x)), y)

    flush_l2 = True  # the whole working set (~20 MB) fits the 126 MB L2: flush between timed steps

    def roofline(self, pk, b200):
        """Dominant kernel = layer-1 forward GEMM (4096 x 784 x 128 + bias + Swish, Z and O written)."""
        c = self.cfg
        B, K, N = c["batch"], c["d_in"], c["d_hid"]
        us, nl = b200.probe_linear_gemm(B, K, N, form=0, with_bias=True, swish=True, rounds=5)
        alg = B * K * 2 + N * K * 2 + 2 * B * N * 4
        ach = alg / (us * 1e-6) / 1e9
        tr = (ncu_traffic() or {}).get("mlp_layer1_fwd")
        return {"bound": "hbm", "kernel": "gemm_tcgen05_kernel: Linear-1 forward 4096x784x128 + bias + Swish epilogue",
                "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": (tr["dram_read"] + tr["dram_write"]) if tr else None,
                "traffic_source": tr["source"] if tr else None,
                "us_per_launch": us, "algorithmic_bytes": alg, "flops_per_launch": 2 * B * K * N,
                "peak_source": pk["source"] + " hbm_gbs",
                "how": "graph-paced launches over rotating operand sets > L2, CUDA events (nnb_probe_linear_gemm)"}

    def cpu(self):
        r = cpu_mlp_run(self.cfg, steps=150, warmup=3)
        return r, f"{r['steps']} steps of batch {r['batch']} (oracle/restated.py)"


class GptWorkload:
    """BASELINE configs[3]: examples/gpt.ipynb model (examples/models.py), Adam, CrossEntropy(ignore pad)."""

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

    flush_l2 = False  # weights + Adam state (0.65 GB) + activations (> 2 GB) >> 126 MB L2: no flush needed

    def linear_shapes(self):
        """(K, N, launches per step) of every nn.Linear on the path (gpt cell 2-7)."""
        c = self.cfg
        d, ff, V, L = c["d_model"], c["d_ff"], c["vocab"], c["layers"]
        return [(d, d, 4 * L, "wq/wk/wv/fc"), (d, ff, L, "ffn.fc_1"), (ff, d, L, "ffn.fc_2"), (d, V, 1, "fc_out")]

    def roofline(self, pk, b200):
        """Dominant kernel = gemm_tcgen05_kernel. One step launches it in 12 Linear shape x form classes
        (4 layer shapes x fwd/dgrad/wgrad); every class is timed GPU-paced on rotating operand sets > L2 and
        `achieved` = (sum of algorithmic 2MKN over all Linear GEMM launches of a step) / (sum of their times)."""
        c = self.cfg
        M_ = c["batch"] * c["seq"]
        tot_fl, tot_us, rows = 0.0, 0.0, []
        for K, N, n, name in self.linear_shapes():
            for form, fname in enumerate(("fwd", "dgrad", "wgrad")):
                us, nl = b200.probe_linear_gemm(M_, K, N, form=form, with_bias=(form == 0), rounds=3)
                fl = 2.0 * M_ * K * N
                tot_fl += fl * n
                tot_us += us * n
                rows.append({"layer": name, "form": fname, "M": M_, "K": K, "N": N, "per_step": n, "us": round(us, 2),
                             "tflops": round(fl / us * 1e-6, 1), "kernels": nl})
        ach = tot_fl / tot_us * 1e-6
        ngemm = sum(r["per_step"] for r in rows)
        top = max(rows, key=lambda r: r["us"] * r["per_step"])
        return {"bound": "tensor", "kernel": "gemm_tcgen05_kernel: all nn.Linear fwd/dgrad/wgrad launches of one step "
                                             f"({ngemm} GEMMs, M={M_})",
                "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops_sustained"],
                "traffic": (ncu_traffic() or {}).get("traffic_per_launch_weighted"),
                "traffic_note": "DRAM read+write bytes per launch, weighted over the same GEMM classes (profiles/r1_gemm_classes_dram.md; "
                                "algorithmic operand+output bytes per launch: %.0f)" % ((ncu_traffic() or {}).get("algorithmic_bytes_per_launch_weighted", float("nan"))),
                "us_per_launch": tot_us / ngemm,
                "flops_per_step": tot_fl, "gemm_us_per_step": tot_us,
                "largest_class": f"{top['layer']} {top['form']} {top['M']}x{top['K']}x{top['N']}: {top['us']} us, {top['tflops']} TFLOP/s",
                "classes": rows,
                "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                "how": "graph-paced launches over rotating operand sets > L2, CUDA events (nnb_probe_linear_gemm)"}

    def cpu(self):
        r = cpu_gpt_run(self.cfg, steps=3, warmup=1, batch=8)
        return r, f"{r['steps']} steps of 8 sequences x T={self.cfg['seq']} (oracle/gpt_numpy.py, dropout off -- the CPU arm does slightly LESS work than the GPU arm; cost is linear in batch)"


def run_ours(args):
    import torch
    import torch.distributed as dist

    import neunet
    import neunet.nn as nn
    from neunet import b200, optim
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
    gemm_sms = args.gemm_sms if args.gemm_sms >= 0 else DEFAULT_DP_GEMM_SMS
    if world > 1 and gemm_sms > 0:
        b200.lib().nnb_set_sm_budget(int(gemm_sms))
    torch.manual_seed(1234 + rank)
    pk = peaks()
    _, label = workload_label(args)
    wl = (GptWorkload if args.workload == "gpt" else MlpWorkload)(neunet, nn, optim, rank)
    B, opt, params = wl.B, wl.opt, wl.params

    bucket = GradBucket(params) if world > 1 else None
    if bucket is not None:
        bucket.broadcast_parameters()
        opt.grad_scale = 1.0 / world

    def train_step(*inputs):
        opt.zero_grad()
        loss = wl.forward_loss(*inputs)
        loss.backward()  # with overlap on, each ~32 MB gradient chunk is all-reduced as soon as it is final
        if bucket is not None:
            bucket.all_reduce()  # the one collective: sum of gradients over NVLink (NCCL); waits for the chunks
        opt.step()
        return loss

    # ---- warm-up (eager) then capture the whole step as a CUDA graph ------------------------------
    W = max(args.warmup, 3)
    overlap = bucket is not None and not args.no_overlap
    hb("eager warm-up")
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
            hb(f"graph capture (attempt {attempt}, overlap={overlap})")
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
            # retry the capture once, on every rank, with the plain end-of-backward all-reduce
            overlap = False
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

    hb("graph warm-up" if graphed is not None else "eager warm-up 2")
    for _ in range(W):
        if flush is not None:
            flush.zero_()
        one_step()
    barrier()
    hb("timed region")

    # ---- timed: K steps, each bracketed by CUDA events, L2 flushed in between ---------------------
    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    with ClockSampler(local) as clk:
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
        # keep the GPU busy ~1 s longer so nvidia-smi (100 ms period) sees clocks under this load. The number of
        # extra steps is derived from the AGREED (max over ranks) step time: every step contains collectives, so
        # all ranks must run the same count (round 1 looped on each rank's own clock here and hung at 8 GPUs).
        hb("clock-sampling tail")
        for _ in range(int(min(2000, max(1, round(1.0 / max(dev_s / K, 1e-5)))))):
            one_step()
        barrier()
    # launches per step: count one eager step (graph replays do not pass through the counter)
    hb("launch count (one eager step)")
    b200.reset_launch_count()
    train_step(*wl.inputs)
    torch.cuda.synchronize()
    launches_per_step = b200.launch_count()
    value = B * world * K / dev_s

    hb("e2e")
    # ---- e2e: host batches through the public API, H2D + D2H inside the timed region --------------
    n_host = len(wl.host)

    def e2e_step(i):
        hb = wl.host[i % n_host]
        if graphed is not None:
            graphed.load(*hb)
            loss = graphed.replay()
        else:
            for dst, src in zip(wl.inputs, hb):
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
    e2e_value = B * world * K / float(t.item())
    h2d = sum(h.numel() * h.element_size() for h in wl.host[0])
    d2h = 4

    hb("roofline probes")
    try:
        roof = wl.roofline(pk, b200)
    except Exception as e:  # a failed probe must not lose the step measurements above
        roof = {"bound": None, "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                "error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        hb("cpu baseline")
        use_all_host_threads()
        cpu, cpu_sample = wl.cpu()
        blas, cores = host_threads()
        line = {
            "metric": "training samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": dev_s / K * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": label, "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2": ("flushed (256 MiB write) between timed steps" if flush is not None else
                              "not flushed: per-step working set (weights + Adam state 0.65 GB, activations > 2 GB) exceeds the 126 MB L2"),
                       "gemm_sms": (gemm_sms if (world > 1 and gemm_sms > 0) else "all"),
                       "grad_allreduce": (None if bucket is None else
                                          ("chunked (32 MB), overlapped with backward" if overlap else "one flat all-reduce after backward")),
                       "step_execution": "cuda-graph replay of the public-API step" if graphed is not None else "eager",
                       "graph_error": graph_err, "precision": b200.get_precision(),
                       "host_wall_ms_per_step": wall / K * 1e3},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_step * K,
            "roofline": roof,
            "cpu_baseline": {"value": cpu["samples_per_s"], "unit": "samples/s", "cores": blas, "kind": "port",
                             "sample": cpu_sample + f", NumPy/OpenBLAS, {cores} cores visible"},
            "final_loss": last_loss,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear-down: no collective after the timing all-reduces. Destroying the NCCL communicator while CUDA
        # graphs that captured its kernels are alive can block forever (seen on the 2-GPU box: both ranks hung
        # in destroy_process_group after printing), so drop the graph, drain the device and leave without it.
        graphed = None
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
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gpt", choices=["mlp", "gpt"],
                    help="gpt = BASELINE.json configs[3], the config the 1/2/4/8-GPU samples/s metric is quoted on (default); "
                         "mlp = configs[1]")
    ap.add_argument("--gemm-sms", type=int, default=-1,
                    help="N>1: SMs the persistent GEMM grids are sized for (rest is left to NCCL's CTAs); 0 = all, -1 = default")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: one flat all-reduce after backward instead of overlapped chunks")
    ap.add_argument("--watchdog", type=int, default=600,
                    help="seconds after which a still-running bench dumps all thread stacks and exits non-zero (0 = off)")
    ap.add_argument("--no-graph", action="store_true", help="time the eager public-API step instead of a CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
