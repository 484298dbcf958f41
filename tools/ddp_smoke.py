"""2+ rank NCCL smoke of the training pattern the reference uses (Audio/src/models/ast_models.py:66-70 under DDP,
Video/run_class_finetuning.py:739-741): frozen encoder, trainable tokenizer, per-rank batch shard, gradient all-reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/ddp_smoke.py
Checks on every rank that the all-reduced tokenizer gradients equal the mean of the per-rank gradients (gathered
independently) and are identical across ranks.  Prints one line per rank.
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import metatransformer_b200 as mt  # noqa: E402
from metatransformer_b200.distributed import allreduce_gradients, init_process_group, shard_range  # noqa: E402


def main():
    rank, local_rank, world = init_process_group("nccl")
    dev = torch.device("cuda", local_rank)
    if "--ddp" in sys.argv and not dist.is_initialized():     # single-process debugging of the DDP wrapper
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    torch.manual_seed(0)                                  # identical weights on every rank
    finetune = "--finetune" in sys.argv      # encoder weights trainable too: 2 x 7.1 M parameters -> several buckets
    enc = mt.build_encoder("base", depth=2).to(dev)
    enc = enc.train() if finetune else enc.eval().requires_grad_(False)
    tok = mt.Data2Seq("image", 768).to(dev)
    global_batch = 8
    g = torch.Generator().manual_seed(123)
    images = torch.rand(global_batch, 3, 224, 224, generator=g)
    lo, hi = shard_range(global_batch, rank, world)
    y = enc(tok(images[lo:hi].to(dev)))
    loss = y.float().square().mean() * (hi - lo) / global_batch * world   # so that the rank-mean equals the global mean
    loss.backward()
    params = [p for p in list(tok.parameters()) + list(enc.parameters()) if p.requires_grad]
    local = [p.grad.clone() for p in params]
    if "--ddp" in sys.argv:
        # same step through torch's DistributedDataParallel: parameters are inputs of the Block / tokenizer autograd
        # nodes, so DDP's bucket hooks fire as each Block's backward finishes (all-reduce overlapped with the rest)
        import torch.nn as nn
        for p in params:
            p.grad = None
        ddp = nn.parallel.DistributedDataParallel(nn.Sequential(tok, enc), device_ids=[local_rank],
                                                  bucket_cap_mb=8, broadcast_buffers=False)
        y = ddp(images[lo:hi].to(dev))
        (y.float().square().mean() * (hi - lo) / global_batch * world).backward()
        n_coll = -1
    else:
        n_coll = allreduce_gradients(params, world)
    # independent check: gather every rank's local gradients and average them
    ok = True
    for p, lg in zip(params, local):
        bucket = [torch.empty_like(lg) for _ in range(world)]
        if world > 1:
            dist.all_gather(bucket, lg)
        else:
            bucket = [lg]
        mean = torch.stack(bucket).mean(0)
        # (the DDP pass is a second backward: split-K / dQ atomics make it equal only to fp32 summation order)
        tol = dict(rtol=2e-2, atol=1e-5) if "--ddp" in sys.argv else dict(rtol=1e-4, atol=1e-6)
        ok &= bool(torch.allclose(p.grad, mean, **tol))
    nbytes = sum(p.numel() * p.element_size() for p in params)
    print(f"rank {rank}/{world}: finetune={finetune} grad bytes {nbytes / 1e6:.1f} MB shard [{lo},{hi}) loss {loss.item():.5f} collectives {n_coll} "
          f"grad_norm {params[0].grad.norm().item():.5f} allreduce_ok={ok}", flush=True)
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
