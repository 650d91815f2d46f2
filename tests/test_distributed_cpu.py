"""world_size-2 data-parallel test on CPU (gloo): two processes train the same MLP on different
batch shards with the flat-bucket all-reduce + grad_scale=1/world; both must end with identical
parameters equal to a single-process run on the concatenated batch."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT

torch = pytest.importorskip("torch")
mp = pytest.importorskip("torch.multiprocessing")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(seed=0):
    import neunet.nn as nn
    np.random.seed(seed)
    return nn.Linear(12, 8), nn.Linear(8, 4)


def _step(l1, l2, opt, x, y, bucket=None, combined=False):
    import neunet
    import neunet.nn as nn
    opt.zero_grad()
    loss = nn.CrossEntropyLoss()(l2(nn.Swish()(l1(neunet.tensor(x)))), neunet.tensor(y, dtype=np.int32))
    loss.backward()
    if bucket is not None and combined:
        bucket.all_reduce_and_step(opt)  # on CPU: falls back to all_reduce() + step() (the chunk-wise optimizer is a device path)
        return float(loss.data)
    if bucket is not None:
        bucket.all_reduce()
    opt.step()
    return float(loss.data)


def _worker(rank, world, port, xs, ys, out, overlap=False):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neunet.distributed import GradBucket
    from neunet.optim import AdamW
    l1, l2 = _make(seed=rank)          # different init per rank on purpose ...
    params = l1.parameters() + l2.parameters()
    bucket = GradBucket(params, chunk_bytes=64)   # tiny chunks: several async all-reduces per step
    bucket.broadcast_parameters(0)     # ... broadcast makes them identical
    opt = AdamW(params, lr=1e-2)
    opt.grad_scale = 1.0 / world
    for t in range(3):
        _step(l1, l2, opt, xs[rank], ys[rank], bucket, combined=(t == 2))
        if overlap and t == 0:
            bucket.overlap_backward()   # chunks reduced from Tensor.backward's ready-hooks from now on
            assert len(bucket._chunks) >= 2
    if overlap:
        assert all(p.grad is None or np.shares_memory(p.grad, bucket.flat) for p in params)
    out[rank] = [p.data.copy() for p in params]
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [False, True])
def test_two_rank_data_parallel_matches_single_process(overlap):
    rng = np.random.RandomState(0)
    xs = [rng.randn(6, 12).astype(np.float32) for _ in range(2)]
    ys = [rng.randint(0, 4, 6).astype(np.int32) for _ in range(2)]
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, xs, ys, out, overlap), nprocs=2, join=True)
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a, b)          # replicas stay bit-identical
    # single process on the concatenated batch: mean loss over 12 = mean of the two shard means
    from neunet.optim import AdamW
    l1, l2 = _make(seed=0)
    params = l1.parameters() + l2.parameters()
    opt = AdamW(params, lr=1e-2)
    for t in range(3):
        _step(l1, l2, opt, np.concatenate(xs), np.concatenate(ys))
    for a, p in zip(out[0], params):
        np.testing.assert_allclose(a, p.data, rtol=2e-4, atol=2e-6)


def test_bucket_skips_params_without_grad():
    from neunet import Tensor
    from neunet.distributed import GradBucket
    ps = [Tensor(np.ones(3), requires_grad=True), Tensor(np.ones(2), requires_grad=True), Tensor(np.ones(4), requires_grad=True)]
    ps[0].grad, ps[2].grad = np.full(3, 2.0, np.float32), np.full(4, 3.0, np.float32)
    b = GradBucket(ps)
    b.all_reduce()
    assert ps[1].grad is None and ps[0].grad.sum() == 6 and ps[2].grad.sum() == 12
    assert np.shares_memory(ps[0].grad, b.flat)


def test_ready_hooks_fire_once_per_leaf_after_its_last_use():
    """Tensor.backward calls `_grad_ready` when a leaf's gradient is final -- also for a leaf used twice."""
    import neunet
    from neunet import Tensor
    w = Tensor(np.ones((3, 3)), requires_grad=True)
    v = Tensor(np.ones((3, 3)), requires_grad=True)
    seen = []
    w._grad_ready = lambda t: seen.append(("w", t.grad.copy()))
    v._grad_ready = lambda t: seen.append(("v", t.grad.copy()))
    x = Tensor(np.arange(9.0).reshape(3, 3), requires_grad=False)
    y = neunet.matmul(neunet.matmul(x, w), w) * v     # w feeds two tape nodes
    y.backward()
    names = [n for n, _ in seen]
    assert sorted(names) == ["v", "w"]
    for n, g in seen:                                  # the hook saw the FINAL gradient
        np.testing.assert_array_equal(g, (w if n == "w" else v).grad)
