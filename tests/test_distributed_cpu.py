"""World-size-2 gloo test of the data-parallel host logic: batch sharding and the bucketed
gradient all-reduce used after backward.  Runs on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metatransformer_b200.distributed import GradBucketReducer, allreduce_gradients, shard_range


def test_shard_range_partitions_batch():
    for gb in (1, 7, 8, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_range(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Linear(32, 4))
    model[1].weight.requires_grad_(False)          # frozen parameter: must be skipped
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(5, 16, generator=g)
    model(x).square().mean().backward()
    local = [p.grad.flatten().tolist() for p in model.parameters() if p.grad is not None]
    n = allreduce_gradients(model.parameters(), world, bucket_bytes=1024)
    out = [p.grad.flatten().tolist() for p in model.parameters() if p.grad is not None]
    q.put((rank, n, local, out))      # plain lists: tensors sent through a Queue race with the sender's exit
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_allreduce_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, n0, l0, o0), (_, n1, l1, o1) = res
    assert n0 == n1 and n0 >= 2                    # several 1 KB buckets
    assert len(o0) == 3                            # weight0, bias0, bias1 (frozen weight has no grad)
    for a, b, x, y in zip(l0, l1, o0, o1):
        a, b, x, y = (torch.tensor(t) for t in (a, b, x, y))
        assert torch.allclose(x, (a + b) / 2, atol=1e-6) and torch.allclose(x, y)


def _worker_reducer(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
    model[2].weight.requires_grad_(False)          # frozen parameter: no bucket slot
    red = GradBucketReducer(model.parameters(), world, bucket_bytes=600)      # several small buckets
    outs = []
    for step in range(2):                          # two steps: buckets are persistent and re-armed
        g = torch.Generator().manual_seed(100 + 10 * step + rank)
        x = torch.randn(5, 16, generator=g)
        red.zero()
        model(x).square().mean().backward()
        n = red.finish()
        outs.append((n, [p.grad.flatten().tolist() for p in model.parameters() if p.requires_grad]))
    # reference: same two steps, local gradients of both ranks averaged
    refs = []
    for step in range(2):
        acc = None
        for r in range(world):
            m2 = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
            m2.load_state_dict(model.state_dict())
            m2[2].weight.requires_grad_(False)
            g = torch.Generator().manual_seed(100 + 10 * step + r)
            m2(torch.randn(5, 16, generator=g)).square().mean().backward()
            gr = [p.grad for p in m2.parameters() if p.requires_grad]
            acc = gr if acc is None else [a + b for a, b in zip(acc, gr)]
        refs.append([(a / world).flatten().tolist() for a in acc])
    q.put((rank, len(red.buckets), outs, refs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_bucketed_overlapped_reducer_world2_gloo():
    """GradBucketReducer: .grad views into persistent flat buckets, one async all-reduce per bucket issued from the
    post-accumulate hooks during backward, averaged in finish()."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_reducer, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, nb, outs, refs in res:
        assert nb >= 2
        for (n, grads), ref in zip(outs, refs):
            assert n == nb
            for a, b in zip(grads, ref):
                assert torch.allclose(torch.tensor(a), torch.tensor(b), atol=1e-6)
    for a, b in zip(res[0][2][1][1], res[1][2][1][1]):
        assert a == b                              # both ranks hold identical averaged gradients
