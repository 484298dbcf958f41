#!/usr/bin/env python
"""Per-kernel microbenchmarks at a BASELINE shape: each kernel alone, back to back for >= `--seconds` (sustained clocks),
rotating buffers larger than L2, CUDA events.  Works from any checkout of the package (run it with that checkout as cwd):

    python tools/bench_kernels.py [--config c2] [--seconds 1.0] [--only attention,qkv,proj,fc1,fc2]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.getcwd())
CONFIGS = {"c2": (256, 196, 768, 12), "c3": (128, 256, 1024, 16), "c4": (64, 512, 768, 12), "c5": (32, 1568, 1024, 16)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--only", default="attention,qkv,proj,fc1,fc2")
    args = ap.parse_args()
    import torch
    import metatransformer_b200 as mt
    from metatransformer_b200 import ops
    B, N, d, H = CONFIGS[args.config]
    M = B * N
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    NB = 3

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).bfloat16().to(dev)

    res = {}

    def timeit(name, fn, flops=None, nbytes=None):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        per = e0.elapsed_time(e1) / 20
        iters = max(20, int(args.seconds * 1e3 / per))
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        r = {"us": round(us, 2), "iters": iters}
        if flops:
            r["tflops"] = round(flops / us / 1e6, 1)
        if nbytes:
            r["gbs"] = round(nbytes / us / 1e3, 1)
        res[name] = r

    only = set(args.only.split(","))
    if "attention" in only:
        qkv = [rnd(B, N, 3 * d) for _ in range(NB)]
        out = torch.empty(B, N, d, dtype=torch.bfloat16, device=dev)
        timeit("attention", lambda i: ops.attention(qkv[i % NB], H, out=out), flops=4.0 * M * N * d, nbytes=4.0 * M * d * 2)
        del qkv
    x = [rnd(M, d) for _ in range(NB)]
    if "qkv" in only:
        w = rnd(3 * d, d, scale=0.02)
        bias = torch.zeros(3 * d, device=dev)
        cs = torch.zeros(3 * d, device=dev)
        stats = torch.zeros(M, 2, device=dev)
        stats[:, 1] = 1
        y = torch.empty(M, 3 * d, dtype=torch.bfloat16, device=dev)
        timeit("qkv (LN fold)", lambda i: ops.linear_ln(x[i % NB], w, bias, cs, stats, out=y), flops=2.0 * M * 3 * d * d)
        del y
    if "fc1" in only:
        w = rnd(4 * d, d, scale=0.02)
        bias = torch.zeros(4 * d, device=dev)
        cs = torch.zeros(4 * d, device=dev)
        stats = torch.zeros(M, 2, device=dev)
        stats[:, 1] = 1
        y = torch.empty(M, 4 * d, dtype=torch.bfloat16, device=dev)
        timeit("fc1 (LN fold + GELU)", lambda i: ops.linear_ln(x[i % NB], w, bias, cs, stats, gelu=True, out=y),
               flops=2.0 * M * 4 * d * d)
        del y
    if "proj" in only:
        w = rnd(d, d, scale=0.02)
        bias = torch.zeros(d, device=dev)
        r = [rnd(M, d) for _ in range(NB)]
        y = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
        timeit("proj (+residual, bf16 stream)", lambda i: ops.linear(x[i % NB], w, bias, residual=r[i % NB], out=y),
               flops=2.0 * M * d * d, nbytes=3.0 * M * d * 2)
        try:
            rl = [rnd(M, d, scale=0.004) for _ in range(NB)]
            ylo = torch.empty_like(y)
            timeit("proj (+residual, split stream)",
                   lambda i: ops.linear(x[i % NB], w, bias, residual=r[i % NB], residual_lo=rl[i % NB], out=y, out_lo=ylo),
                   flops=2.0 * M * d * d, nbytes=5.0 * M * d * 2)
        except TypeError:
            pass
        del r, y
    if "fc2" in only:
        w = rnd(d, 4 * d, scale=0.02)
        bias = torch.zeros(d, device=dev)
        hid = [rnd(M, 4 * d) for _ in range(2)]
        r = [rnd(M, d) for _ in range(NB)]
        y = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
        timeit("fc2 (+residual, bf16 stream)", lambda i: ops.linear(hid[i % 2], w, bias, residual=r[i % NB], out=y),
               flops=2.0 * M * 4 * d * d)
        try:
            rl = [rnd(M, d, scale=0.004) for _ in range(NB)]
            ylo = torch.empty_like(y)
            timeit("fc2 (+residual, split stream)",
                   lambda i: ops.linear(hid[i % 2], w, bias, residual=r[i % NB], residual_lo=rl[i % NB], out=y, out_lo=ylo),
                   flops=2.0 * M * 4 * d * d)
        except TypeError:
            pass
    print(json.dumps({"config": args.config, "cwd": os.getcwd(), "results": res}))


if __name__ == "__main__":
    main()
