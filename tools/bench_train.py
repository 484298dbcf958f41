#!/usr/bin/env python
"""Training-step timing, 1..8 GPUs (data parallel, one process per GPU, NCCL): the only collective this path has is the
gradient all-reduce of the trainable parameters (Video/run_class_finetuning.py:739-741).

    python tools/bench_train.py [--mode frozen|finetune] [--config c2|c4] [--steps 10]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_train.py --mode finetune

  frozen   : frozen Base-16 encoder + trainable tokenizer (the reference's AST / X-Ray recipe) — the all-reduce carries
             0.8 MB (audio conv, c4) / 2.4 MB (image conv, c2): latency-bound, one bucket
  finetune : every encoder parameter trains — 340 MB of fp32 gradients per step: bandwidth-bound, bucketed and issued from
             gradient hooks while the backward of the earlier Blocks still runs (metatransformer_b200.distributed.
             GradBucketReducer); `--no-overlap` reduces after the backward instead (the exposed-communication baseline)
Prints one JSON line (rank 0): per-step ms (CUDA events, max over ranks) for forward, backward (incl. the overlapped
collectives' launch) and the exposed wait in finish(), tokens/s over all ranks, gradient bytes, buckets.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="finetune", choices=["frozen", "finetune"])
    ap.add_argument("--config", default="c2", choices=["c2", "c4"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the BASELINE config's)")
    ap.add_argument("--bucket-mb", type=int, default=64)
    ap.add_argument("--no-overlap", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import metatransformer_b200 as mt
    from metatransformer_b200.distributed import GradBucketReducer, allreduce_gradients, init_process_group

    rank, local_rank, world = init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(0)
    enc = mt.build_encoder("base", depth=args.depth).to(dev)
    if args.mode == "finetune":
        enc.train()
    else:
        enc.eval().requires_grad_(False)
    if args.config == "c2":
        B, N = args.batch or 256, 196
        tok = mt.Data2Seq("image", 768, token_dtype=torch.bfloat16).to(dev)
        g = torch.Generator().manual_seed(100 + rank)
        data = torch.rand(B, 3, 224, 224, generator=g).to(dev)
        feed = lambda: tok(data)                                               # noqa: E731
    else:
        B, N = args.batch or 64, 512
        tok = mt.data2seq.Acoustic.PatchEmbed(embed_dim=768, fstride=16, tstride=16, token_dtype=torch.bfloat16).to(dev)
        g = torch.Generator().manual_seed(100 + rank)
        data = torch.randn(B, 1, 128, 1024, generator=g).to(dev)
        feed = lambda: tok(data)                                               # noqa: E731
    params = [p for p in list(tok.parameters()) + list(enc.parameters()) if p.requires_grad]
    grad_bytes = sum(p.numel() * 4 for p in params)
    red = None if args.no_overlap else GradBucketReducer(params, world, bucket_bytes=args.bucket_mb << 20)

    def step():
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        if red is not None:
            red.zero()
        else:
            for p in params:
                p.grad = None
        e[0].record()
        y = enc(feed())
        loss = y.float().square().mean()
        e[1].record()
        loss.backward()
        e[2].record()
        n = red.finish() if red is not None else allreduce_gradients(params, world, bucket_bytes=args.bucket_mb << 20)
        e[3].record()
        return e, n

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    acc = [0.0, 0.0, 0.0]
    ncoll = 0
    for _ in range(args.steps):
        e, ncoll = step()
        torch.cuda.synchronize()
        for i in range(3):
            acc[i] += e[i].elapsed_time(e[i + 1])
    ms = torch.tensor([a / args.steps for a in acc], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    fw, bw, ar = [float(v) for v in ms.tolist()]
    if rank == 0:
        print(json.dumps({
            "what": "training step", "mode": args.mode, "config": args.config, "n_gpus": world, "per_gpu_batch": B,
            "seq_len": N, "depth": args.depth, "overlap": red is not None, "forward_ms": fw, "backward_ms": bw,
            "exposed_allreduce_ms": ar, "step_ms": fw + bw + ar, "tokens_per_s": world * B * N / ((fw + bw + ar) * 1e-3),
            "grad_bytes": grad_bytes, "collectives_per_step": ncoll,
            "buckets": None if red is None else len(red.buckets), "timing": "CUDA events, max over ranks"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
