#!/usr/bin/env python
"""One eager training step of a bench workload between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...` launch lists and `--set full` captures (never a bench number).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_gpt.csv python scripts/profile_step.py --workload gpt
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (adds the package paths)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="gpt", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--batch", type=int, default=0)
    args = ap.parse_args()
    import torch
    import neunet
    import neunet.nn as nn
    from neunet import b200, optim
    torch.cuda.set_device(0)
    b200.require_device()
    b200.set_precision("bf16")
    if args.batch:
        bench.workload_label(args.workload)[0]["batch"] = args.batch
    wl = bench.WORKLOADS[args.workload](neunet, nn, optim, 0)

    def step():
        wl.opt.zero_grad()
        loss = wl.forward_loss(*wl.inputs)
        loss.backward()
        wl.opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
