"""N-rank NCCL data-parallel step on real GPUs == single-GPU step on the concatenated batch.

Needs >= 2 CUDA devices (skipped otherwise; the driver's 1-GPU test box skips it, `gpurun --gpus 2` runs it).
Each rank trains the same small GPT-shaped stack (Linear / RMSNorm / Swish / CrossEntropy) on its own batch shard
through `neunet.distributed.GradBucket` -- first the flat all-reduce, then the chunked all-reduce overlapped
with backward -- with the 1/world average folded into Adam (`grad_scale`). Because the loss is a per-shard mean
and the shards have equal size, the parameters after every step must equal those of one process that sees
the whole batch. CPU/gloo coverage of the same logic: tests/test_distributed_cpu.py.
"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(device):
    import neunet.nn as nn
    np.random.seed(7)
    layers = [nn.Linear(64, 128), nn.RMSNorm(128), nn.Swish(), nn.Linear(128, 96), nn.Linear(96, 32)]
    return [l.to(device) for l in layers]


def _params(layers):
    ps = []
    for l in layers:
        ps += l.parameters()
    return ps


def _step(layers, opt, x, y, bucket=None, device="cuda", chunked_step=False):
    import neunet
    import neunet.nn as nn
    opt.zero_grad()
    h = neunet.tensor(x, device=device)
    for l in layers:
        h = l(h)
    loss = nn.CrossEntropyLoss()(h, neunet.tensor(y, dtype=np.int32, device=device))
    loss.backward()
    if bucket is not None and chunked_step:
        bucket.all_reduce_and_step(opt)  # the optimizer of each chunk right behind that chunk's all-reduce
        return loss
    if bucket is not None:
        bucket.all_reduce()
    opt.step()
    return loss


def _worker(rank, world, port, xs, ys, ret, native=False):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import datetime
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=60))
    from neunet import b200
    from neunet.distributed import GradBucket
    from neunet.optim import Adam
    b200.set_precision("bf16x3")
    layers = _build("cuda")
    params = _params(layers)
    comm = b200.NativeComm(rank, world) if native else None  # nnb_comm_*: NCCL behind the C-ABI instead of torch.distributed
    bucket = GradBucket(params, chunk_bytes=16 << 10, native_comm=comm)  # several chunks per step
    bucket.broadcast_parameters(0)
    opt = Adam(params, lr=1e-2, eps=1e-3)  # large eps: near-zero gradients must not turn round-off into sign flips
    opt.grad_scale = 1.0 / world
    for t in range(4):
        _step(layers, opt, xs[t][rank], ys[t][rank], bucket, chunked_step=(t == 3))
        if t == 1:
            bucket.overlap_backward()  # steps 2, 3: chunked all-reduce launched from the ready-hooks
    torch.cuda.synchronize()
    ret[rank] = [p.data.cpu().numpy().copy() for p in params]
    dist.barrier()
    if comm is not None:
        comm.destroy()
    dist.destroy_process_group()


@pytest.mark.parametrize("native", [False, True], ids=["torch.distributed", "nnb_comm"])
@pytest.mark.parametrize("world", [2])
def test_nccl_ranks_equal_single_gpu_on_concatenated_batch(world, native):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} CUDA devices")
    import torch.multiprocessing as mp
    from neunet import b200
    from neunet.optim import Adam
    rng = np.random.RandomState(0)
    per = 48
    xs = [[rng.randn(per, 64).astype(np.float32) for _ in range(world)] for _ in range(4)]
    ys = [[rng.randint(0, 32, per).astype(np.int32) for _ in range(world)] for _ in range(4)]
    # single-GPU run on the concatenated batch
    torch.cuda.set_device(0)
    b200.set_precision("bf16x3")
    layers = _build("cuda")
    params = _params(layers)
    opt = Adam(params, lr=1e-2, eps=1e-3)  # large eps: near-zero gradients must not turn round-off into sign flips
    for t in range(4):
        _step(layers, opt, np.concatenate(xs[t]), np.concatenate(ys[t]))
    want = [p.data.cpu().numpy() for p in params]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, xs, ys, ret, native)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("a rank did not finish (collective hang?)")
        assert p.exitcode == 0
    for r in range(world):
        for got, ref in zip(ret[r], want):
            # max-norm relative error; bf16x3 contractions + a different summation split of the batch
            assert np.abs(got - ref).max() <= 2e-4 * max(np.abs(ref).max(), 1e-3)
    for got, other in zip(ret[0], ret[1]):
        np.testing.assert_array_equal(got, other)  # replicas stay bit-identical
