"""Batch-sharded data parallelism: the ONE collective of the training step.

The reference has no distributed code at all; this is the B200-native addition the
north-star asks for: one process per GPU (``torch.distributed`` / NCCL over NVLink 5 + NVSwitch for
the plumbing), identical parameters on every rank, a per-rank shard of the batch, and a single
sum all-reduce over a FLAT gradient bucket per step. The 1/world_size averaging is not a separate
pass: it is folded into the multi-tensor Adam(W) kernel (``optimizer.grad_scale``).

Parameters whose ``.grad`` is None (e.g. the GPT example's never-used ``cross_attn``) are skipped
exactly as ``optim.py:21-22`` skips them, so they cost no bandwidth.
"""
from __future__ import annotations

import numpy as np


def is_initialized():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized()
    except Exception:
        return False


def world_size():
    import torch.distributed as dist
    return dist.get_world_size() if is_initialized() else 1


def rank():
    import torch.distributed as dist
    return dist.get_rank() if is_initialized() else 0


class GradBucket:
    """Flat fp32 bucket holding every parameter's gradient; ``all_reduce()`` packs the current
    ``param.grad`` arrays into it, runs one NCCL (or gloo, on CPU) all-reduce and points each
    ``param.grad`` at its slice of the reduced bucket."""

    def __init__(self, params):
        self.params = list(params)
        self.sizes = [int(np.prod(p.shape)) for p in self.params]
        self.offsets = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        self.device = self.params[0].device if self.params else "cpu"
        total = int(self.offsets[-1])
        if self.device == "cuda":
            import torch
            self.flat = torch.zeros(total, dtype=torch.float32, device="cuda")
        else:
            self.flat = np.zeros(total, dtype=np.float32)
        self._live = None

    def _slice(self, i):
        return self.flat[int(self.offsets[i]): int(self.offsets[i + 1])]

    def broadcast_parameters(self, src=0):
        """Make every rank start from rank `src`'s parameters."""
        if world_size() == 1:
            return
        import torch
        import torch.distributed as dist
        for p in self.params:
            if self.device == "cuda":
                dist.broadcast(p.data, src)
            else:
                t = torch.from_numpy(np.ascontiguousarray(p.data))
                dist.broadcast(t, src)
                p.data[...] = t.numpy()

    def all_reduce(self):
        """Sum gradients over all ranks. Live set = parameters with a gradient on THIS rank; it must
        be the same on every rank (it is, for replicated models)."""
        live = [i for i, p in enumerate(self.params) if p.grad is not None]
        if not live:
            return
        if self.device == "cuda":
            import torch
            srcs = [self.params[i].grad.reshape(-1) for i in live]
            dsts = [self._slice(i) for i in live]
            torch._foreach_copy_(dsts, srcs)
            # contiguous run of live slices -> one collective over [lo, hi)
            lo, hi = int(self.offsets[live[0]]), int(self.offsets[live[-1] + 1])
            if len(live) != live[-1] - live[0] + 1:  # holes (grad None): zero them so the sum is unaffected
                for i in range(live[0], live[-1] + 1):
                    if self.params[i].grad is None:
                        self._slice(i).zero_()
            if world_size() > 1:
                import torch.distributed as dist
                dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM)
            for i in live:
                self.params[i].grad = self._slice(i).reshape(tuple(self.params[i].shape))
        else:
            for i in live:
                self._slice(i)[...] = np.asarray(self.params[i].grad, dtype=np.float32).reshape(-1)
            lo, hi = int(self.offsets[live[0]]), int(self.offsets[live[-1] + 1])
            for i in range(live[0], live[-1] + 1):
                if self.params[i].grad is None:
                    self._slice(i)[...] = 0
            if world_size() > 1:
                import torch
                import torch.distributed as dist
                t = torch.from_numpy(self.flat[lo:hi])
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            for i in live:
                self.params[i].grad = self._slice(i).reshape(tuple(self.params[i].shape))
