"""Data-parallel plumbing for the encoder path: one process per GPU, torch.distributed (NCCL over
NVLink on the GPU box, gloo in CPU tests).  The forward needs no collective — ranks are pure replicas over a batch
shard; only trainable-parameter gradients are all-reduced after backward (DDP semantics,
Video/run_class_finetuning.py:739-741).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (Video/utils.py:295-298)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_process_group(backend: Optional[str] = None) -> Tuple[int, int, int]:
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of the samples owned by `rank` (first `global_batch % world` ranks get +1)."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world: int, bucket_bytes: int = 25 << 20,
                        group=None) -> int:
    """Average `.grad` of the trainable parameters over ranks with flat-bucket all-reduces (sum, then / world).
    Returns the number of collectives issued.  With a frozen encoder only the tokenizer / head gradients exist
    (0.79 MB audio conv, 6.3 MB video tubelets: one bucket)."""
    grads: List[torch.Tensor] = [p.grad for p in params if p.requires_grad and p.grad is not None]
    if world <= 1 or not grads:
        return 0
    n_coll = 0
    bucket: List[torch.Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size, n_coll
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_coll += 1
        bucket, size = [], 0

    for g in grads:
        if bucket and (size + g.numel() * g.element_size() > bucket_bytes or g.dtype != bucket[0].dtype):
            flush()
        bucket.append(g)
        size += g.numel() * g.element_size()
    flush()
    return n_coll


def max_over_ranks(value: float, device: torch.device, world: int) -> float:
    if world <= 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
