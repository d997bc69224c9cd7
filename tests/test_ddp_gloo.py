"""World-size-2 gloo run of the batch-sharding helpers on CPU: sharded DDP gradients == full-batch gradients."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from mop_b200 import ddp
    r, _, w = ddp.init("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.GELU(), torch.nn.Linear(16, 5))
    net = ddp.wrap(model)
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(10, 12, generator=g), torch.randint(0, 5, (10,), generator=g)
    lo, hi = ddp.shard_bounds(10, rank, world)
    loss = torch.nn.functional.cross_entropy(net(X[lo:hi]), Y[lo:hi], reduction="sum") / 10 * world
    loss.backward()  # DDP averages over ranks: sum-over-shard/10*world averaged == full-batch mean loss grad
    slow = ddp.max_over_ranks(float(rank + 1), "cpu")
    if rank == 0:
        torch.save({"grads": [p.grad.clone() for p in model.parameters()], "slow": slow}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradients_match_full_batch(tmp_path):
    from mop_b200 import ddp
    assert [ddp.shard_bounds(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.GELU(), torch.nn.Linear(16, 5))
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(10, 12, generator=g), torch.randint(0, 5, (10,), generator=g)
    torch.nn.functional.cross_entropy(model(X), Y).backward()
    for a, p in zip(got["grads"], model.parameters()):
        assert torch.allclose(a, p.grad, atol=1e-6)
    assert got["slow"] == 2.0


def _worker_flat(rank, world, port, out, bucket_mb=0.0, pack=False):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from mop_b200 import ddp
    ddp.init("gloo")
    torch.manual_seed(rank)   # different initial weights per rank: the constructor broadcasts rank 0's
    model = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.GELU(), torch.nn.Linear(16, 5))
    flat = ddp.FlatGradAllReduce(model, bucket_mb=bucket_mb, pack=pack)
    if bucket_mb > 0:
        assert len(flat.buckets) >= 2 and flat.buckets[0][1] == flat.flat.numel() and flat.buckets[-1][0] == 0
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(10, 12, generator=g), torch.randint(0, 5, (10,), generator=g)
    lo, hi = ddp.shard_bounds(10, rank, world)
    for _ in range(2):   # the second pass checks zero(): gradients must not accumulate across steps
        flat.zero()
        loss = torch.nn.functional.cross_entropy(model(X[lo:hi]), Y[lo:hi], reduction="sum") / 10 * world
        loss.backward()
        flat.reduce()
    assert all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in model.parameters())   # still views of the flat buffer
    if rank == 0:
        torch.save({"grads": [p.grad.clone() for p in model.parameters()], "w": [p.detach().clone() for p in model.parameters()]}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_matches_full_batch(tmp_path):
    """FlatGradAllReduce (one flat buffer, one all-reduce; the multi-GPU bench path) == full-batch gradients."""
    out = str(tmp_path / "f.pt")
    mp.spawn(_worker_flat, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.GELU(), torch.nn.Linear(16, 5))
    for a, p in zip(got["w"], model.parameters()):
        assert torch.equal(a, p.detach())
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(10, 12, generator=g), torch.randint(0, 5, (10,), generator=g)
    torch.nn.functional.cross_entropy(model(X), Y).backward()
    for a, p in zip(got["grads"], model.parameters()):
        assert torch.allclose(a, p.grad, atol=1e-6)


def test_bucketed_overlapped_allreduce_matches_full_batch(tmp_path):
    """FlatGradAllReduce(bucket_mb > 0): buckets in reverse parameter order, all-reduces launched from post-accumulate-grad hooks
    during the backward (the ViT-B/16 path: 346 MB of gradients) == full-batch gradients, on both steps."""
    out = str(tmp_path / "fb.pt")
    mp.spawn(_worker_flat, args=(2, _free_port(), out, 0.0002), nprocs=2, join=True)   # ~50-element buckets: two buckets
    got = torch.load(out)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.GELU(), torch.nn.Linear(16, 5))
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(10, 12, generator=g), torch.randint(0, 5, (10,), generator=g)
    torch.nn.functional.cross_entropy(model(X), Y).backward()
    for a, p in zip(got["grads"], model.parameters()):
        assert torch.allclose(a, p.grad, atol=1e-6)


def test_packed_gradient_allreduce_matches_full_batch(tmp_path):
    """FlatGradAllReduce(pack=True), the CUDA-graph bench path at N > 1: autograd keeps fresh gradient tensors (no add_ per parameter),
    pack() concatenates them into the flat buffer, bind() points p.grad at the reduced slices == full-batch gradients, on both steps."""
    out = str(tmp_path / "fp.pt")
    mp.spawn(_worker_flat, args=(2, _free_port(), out, 0.0, True), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.GELU(), torch.nn.Linear(16, 5))
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(10, 12, generator=g), torch.randint(0, 5, (10,), generator=g)
    torch.nn.functional.cross_entropy(model(X), Y).backward()
    for a, p in zip(got["grads"], model.parameters()):
        assert torch.allclose(a, p.grad, atol=1e-6)
