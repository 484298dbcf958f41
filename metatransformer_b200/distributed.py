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


class GradBucketReducer:
    """Bucketed, overlapped gradient all-reduce for the fine-tune / trainable-tokenizer steps (DDP semantics of
    Video/run_class_finetuning.py:739-741, without wrapping the model).

    * Persistent flat buckets: every trainable parameter's `.grad` is a VIEW into one of a few flat fp32 (or bf16) buffers,
      filled in reverse registration order (= the order backward produces them), so a bucket is reduced in place — no
      torch.cat, no copy-back.
    * A post-accumulate hook per parameter counts arrivals; when the last gradient of a bucket has been written the bucket's
      all-reduce is issued asynchronously (NCCL runs it on its own stream after an event on the compute stream), i.e. while
      the backward of the earlier Blocks is still running.  `finish()` waits for the outstanding collectives and divides by
      the world size.  Between steps call `zero()` (or zero the grads in place); never set `.grad = None`.
    Buckets are sized for launch latency and overlap (NVSwitch gives every peer full bandwidth; there is no per-link
    sizing to do): one bucket per ~`bucket_bytes` of gradients, a single bucket for the 0.8-6 MB frozen-encoder cases.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], world: int, bucket_bytes: int = 64 << 20, group=None,
                 grad_dtype: Optional[torch.dtype] = None):
        self.world, self.group = world, group
        self.params = [p for p in params if p.requires_grad]
        self.buckets: List[torch.Tensor] = []
        self._bucket_of = {}
        self._pending: List[int] = []
        self._expected: List[int] = []
        self._works = []
        self._hooks = []
        self.collectives_issued = 0
        cur: List[torch.nn.Parameter] = []
        size = 0
        groups: List[List[torch.nn.Parameter]] = []
        for p in reversed(self.params):
            nbytes = p.numel() * (torch.empty((), dtype=grad_dtype or p.dtype).element_size())
            if cur and (size + nbytes > bucket_bytes or (grad_dtype or p.dtype) != (grad_dtype or cur[0].dtype)
                        or p.device != cur[0].device):
                groups.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            groups.append(cur)
        for bi, grp in enumerate(groups):
            dt = grad_dtype or grp[0].dtype
            flat = torch.zeros(sum(p.numel() for p in grp), dtype=dt, device=grp[0].device)
            off = 0
            for p in grp:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
                self._bucket_of[p] = bi
            self.buckets.append(flat)
            self._expected.append(len(grp))
            self._pending.append(len(grp))
        if world > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        bi = self._bucket_of[p]
        flat = self.buckets[bi]
        if p.grad.data_ptr() < flat.data_ptr() or p.grad.data_ptr() >= flat.data_ptr() + flat.numel() * flat.element_size():
            raise RuntimeError("GradBucketReducer: a parameter's .grad was replaced (set to None / reassigned); zero the "
                               "gradients in place (reducer.zero()) so that they stay views of the flat buckets")
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.collectives_issued += 1

    def finish(self) -> int:
        """Wait for this step's collectives, average, re-arm for the next step.  Returns the number of collectives.
        Buckets whose gradients did not all arrive (unused parameters) are reduced here, after the backward."""
        n = 0
        if self.world > 1:
            for bi, left in enumerate(self._pending):
                if left != 0:
                    self._works.append(dist.all_reduce(self.buckets[bi], op=dist.ReduceOp.SUM, group=self.group,
                                                       async_op=True))
                    self.collectives_issued += 1
            for w in self._works:
                w.wait()
            n = len(self._works)
            for flat in self.buckets:
                flat.div_(self.world)
        self._works = []
        self._pending = list(self._expected)
        return n

    def zero(self) -> None:
        for flat in self.buckets:
            flat.zero_()

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


def max_over_ranks(value: float, device: torch.device, world: int) -> float:
    if world <= 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
