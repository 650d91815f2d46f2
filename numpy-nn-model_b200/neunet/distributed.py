"""Batch-sharded data parallelism: the ONE collective of the training step.

The reference has no distributed code at all; this is the B200-native addition the
north-star asks for: one process per GPU (``torch.distributed`` / NCCL over NVLink 5 + NVSwitch for
the plumbing), identical parameters on every rank, a per-rank shard of the batch, and a single
sum all-reduce over a FLAT gradient bucket per step. The 1/world_size averaging is not a separate
pass: it is folded into the multi-tensor Adam(W) kernel (``optimizer.grad_scale``).

Parameters whose ``.grad`` is None (e.g. the GPT example's never-used ``cross_attn``) are skipped
exactly as ``optim.py:21-22`` skips them, so they cost no bandwidth.

Overlap: ``GradBucket.overlap_backward()`` splits the flat bucket into ~32 MB chunks in REVERSE
parameter order and registers ``_grad_ready`` hooks (``Tensor.backward`` fires one when the last tape
node that feeds a leaf has run). A chunk's all-reduce is issued asynchronously the moment its last
gradient is final, so NCCL runs on its own stream underneath the remaining wgrad/dgrad kernels;
``all_reduce()`` then only waits and handles whatever was not launched (first step, dead parameters).
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
    ``param.grad`` arrays into it, runs the NCCL (or gloo, on CPU) sum all-reduce and points each
    ``param.grad`` at its slice of the reduced bucket."""

    def __init__(self, params, chunk_bytes=32 << 20, native_comm=None):
        """native_comm: a ``neunet.b200.NativeComm`` -- the all-reduce then goes through the library's own
        ``nnb_comm_allreduce_sum`` (NCCL behind the C-ABI) instead of ``torch.distributed``."""
        self.native_comm = native_comm
        self.params = list(params)
        self.sizes = [int(np.prod(p.shape)) for p in self.params]
        self.offsets = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        self.device = self.params[0].device if self.params else "cpu"
        self.chunk_bytes = int(chunk_bytes)
        total = int(self.offsets[-1])
        if self.device == "cuda":
            import torch
            self.flat = torch.zeros(total, dtype=torch.float32, device="cuda")
        else:
            self.flat = np.zeros(total, dtype=np.float32)
        self._chunks = None      # [(first_param, last_param_exclusive)] in launch (reverse) order
        self._chunk_of = {}      # param index -> chunk index
        self._pending = []       # per chunk: gradients still missing this step
        self._launched = []      # per chunk: async work handle (or True) once issued this step
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._defer = 0          # > 0 inside `accumulate()`: ready-hooks do not launch anything

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

    # ---- packing + collective over a run of parameters [a, b) --------------------------------------
    def _pack(self, a, b):
        live = [i for i in range(a, b) if self.params[i].grad is not None]
        if not live:
            return live
        if self.device == "cuda":
            import torch
            # gradients a layer already wrote into its bucket slice (Parameter._grad_buffer) need no copy
            todo = [i for i in live if self.params[i].grad.data_ptr() != self._slice(i).data_ptr()]
            if todo:
                srcs = [self.params[i].grad.reshape(-1) for i in todo]
                dsts = [self._slice(i) for i in todo]
                torch._foreach_copy_(dsts, srcs)
        else:
            for i in live:
                self._slice(i)[...] = np.asarray(self.params[i].grad, dtype=np.float32).reshape(-1)
        for i in range(live[0], live[-1] + 1):  # holes (grad None) inside the run: zero so the sum is unaffected
            if self.params[i].grad is None:
                if self.device == "cuda":
                    self._slice(i).zero_()
                else:
                    self._slice(i)[...] = 0
        return live

    def _reduce(self, live, async_op=False):
        """One sum all-reduce over the contiguous run covering `live`; returns the work handle."""
        if not live or world_size() == 1:
            return None
        import torch.distributed as dist
        lo, hi = int(self.offsets[live[0]]), int(self.offsets[live[-1] + 1])
        if self.device == "cuda" and self.native_comm is not None:
            return self.native_comm.all_reduce_(self.flat[lo:hi], async_op=async_op)
        if self.device == "cuda":
            return dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=async_op)
        import torch
        return dist.all_reduce(torch.from_numpy(self.flat[lo:hi]), op=dist.ReduceOp.SUM, async_op=async_op)

    def _adopt(self, live):
        for i in live:
            self.params[i].grad = self._slice(i).reshape(tuple(self.params[i].shape))

    # ---- overlap with backward --------------------------------------------------------------------
    def overlap_backward(self, live_only=True):
        """Register ready-hooks so chunks are reduced while backward is still running. Call after one
        ordinary step: with `live_only` the chunks are built from the parameters that received a
        gradient in that step (dead parameters would otherwise keep a chunk from ever completing)."""
        idx = [i for i, p in enumerate(self.params) if (p.grad is not None or not live_only)]
        self._chunks, self._chunk_of = [], {}
        run, run_bytes = [], 0
        for i in reversed(idx):  # backward produces gradients roughly in reverse parameter order
            run.append(i)
            run_bytes += self.sizes[i] * 4
            if run_bytes >= self.chunk_bytes:
                self._chunks.append(sorted(run))
                run, run_bytes = [], 0
        if run:
            self._chunks.append(sorted(run))
        for c, members in enumerate(self._chunks):
            for i in members:
                self._chunk_of[i] = c
        for i, p in enumerate(self.params):
            p._grad_ready = self._on_ready if i in self._chunk_of else None
            # layers that can take an output destination write their gradient straight into the bucket
            p._grad_buffer = (self._slice(i).reshape(tuple(p.shape))
                              if (i in self._chunk_of and self.device == "cuda") else None)
        self._reset_step()

    def _reset_step(self):
        if self._chunks is not None:
            self._pending = [len(m) for m in self._chunks]
            self._launched = [None] * len(self._chunks)

    def accumulate(self):
        """Context manager for steps that call ``backward()`` MORE THAN ONCE before the optimizer (gradient
        accumulation, GAN-style real/fake losses): chunks are not launched from the ready-hooks, so every backward
        accumulates into ``param.grad`` as usual and ``all_reduce()`` reduces everything afterwards.
        Without it an overlapped chunk is reduced as soon as the FIRST backward has produced it; a second
        backward touching the same parameters raises instead of silently dropping its gradient."""
        bucket = self

        class _Ctx:
            def __enter__(self):
                bucket._defer += 1
                return bucket

            def __exit__(self, *exc):
                bucket._defer -= 1

        return _Ctx()

    def _on_ready(self, param):
        i = self._index.get(id(param))
        c = self._chunk_of.get(i)
        if c is None or self._defer:
            return
        if self._launched[c] is not None:
            raise RuntimeError(
                "GradBucket.overlap_backward(): a second backward() reached a parameter whose gradient chunk is "
                "already being all-reduced. Overlap mode supports one backward() per optimizer step; wrap steps "
                "with several backward() calls in `with bucket.accumulate():` (reduces after the last one).")
        self._pending[c] -= 1
        if self._pending[c] == 0:
            members = self._chunks[c]
            live = self._pack(members[0], members[-1] + 1)
            self._launched[c] = (live, self._reduce(live, async_op=True))

    def all_reduce(self):
        """Sum gradients over all ranks. Live set = parameters with a gradient on THIS rank; it must
        be the same on every rank (it is, for replicated models)."""
        if self._chunks is None:
            live = self._pack(0, len(self.params))
            self._reduce(live)
            self._adopt(live)
            return
        done = set()
        for c, st in enumerate(self._launched):
            if st is None:  # chunk never completed during backward: do it now, in chunk order on every rank
                members = self._chunks[c]
                live = self._pack(members[0], members[-1] + 1)
                st = (live, self._reduce(live, async_op=True))
            live, work = st
            if work is not None:
                work.wait()
            self._adopt(live)
            done.update(live)
        rest = [i for i, p in enumerate(self.params) if p.grad is not None and i not in done and i not in self._chunk_of]
        if rest:  # parameters that were dead when the chunks were built but have a gradient now
            live = self._pack(rest[0], rest[-1] + 1)
            self._reduce(live)
            self._adopt(live)
        self._reset_step()

    def all_reduce_and_step(self, optimizer):
        """``all_reduce()`` followed by ``optimizer.step()``, with the optimizer of every gradient chunk launched as soon as
        THAT chunk's all-reduce has landed: chunks are reduced in the order backward produced them, so the Adam(W) update of
        the early chunks (the last layers) runs underneath the collective of the late ones (the embedding, final only when
        backward ends) instead of behind it. Same arithmetic as the two separate calls. Falls back to them whenever the
        overlapped path does not apply (no chunks yet, CPU, an optimizer over a different parameter list, stray gradients)."""
        ok = (self._chunks is not None and self.device == "cuda" and hasattr(optimizer, "step_ranges")
              and len(optimizer.params) == len(self.params) and all(a is b for a, b in zip(optimizer.params, self.params)))
        if ok:
            covered = set(self._chunk_of)
            ok = not any(p.grad is not None and i not in covered for i, p in enumerate(self.params))
        if not ok:
            self.all_reduce()
            optimizer.step()
            return
        states = []
        for c, st in enumerate(self._launched):
            if st is None:  # chunk never completed during backward: launch it now, in chunk order on every rank
                members = self._chunks[c]
                live = self._pack(members[0], members[-1] + 1)
                st = (live, self._reduce(live, async_op=True))
            states.append(st)
            self._adopt(st[0])  # pointers only: the optimizer's gradient table is the bucket slices

        def waiter(work):
            return (lambda: work.wait()) if work is not None else None

        ranges, lo_prev = [], len(self.params)
        for c, (live, work) in enumerate(states):  # launch order = reverse parameter order
            members = self._chunks[c]
            lo = members[0]
            ranges.append((lo, lo_prev - lo, waiter(work)))  # up to the previous chunk: dead parameters in between are skipped
            lo_prev = lo
        if lo_prev > 0:
            ranges.append((0, lo_prev, None))
        optimizer.step_ranges(ranges)
        self._reset_step()
