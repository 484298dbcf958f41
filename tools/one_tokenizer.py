#!/usr/bin/env python
"""Two calls of a Data2Seq conv tokenizer at a BASELINE shape — the target of ncu captures of `im2row_kernel` and the
tokenizer GEMM:   ncu --set full -k regex:'im2row|gemm_bf16' -s 2 -c 2 python tools/one_tokenizer.py --config c5
(the first call packs the weights; launches 2-3 of the filter are the second call's gather and GEMM)."""
import argparse
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c5", choices=["c1", "c4", "c5"])
    ap.add_argument("--iters", type=int, default=2)
    args = ap.parse_args()
    import torch
    import metatransformer_b200 as mt
    if args.config == "c1":
        tok = mt.Data2Seq("image", 768, token_dtype=torch.bfloat16).cuda().eval()
        x = torch.rand(256, 3, 224, 224, device="cuda")
        nbytes = x.numel() * 4 + 256 * 196 * 768 * 2 + 768 * 768 * 2
    elif args.config == "c4":
        tok = mt.data2seq.Acoustic.PatchEmbed(embed_dim=768, fstride=16, tstride=16, token_dtype=torch.bfloat16).cuda().eval()
        x = torch.randn(64, 1, 128, 1024, device="cuda")
        nbytes = x.numel() * 4 + 64 * 512 * 768 * 2 + 768 * 256 * 2
    else:
        tok = mt.data2seq.Video.PatchEmbed(embed_dim=1024, token_dtype=torch.bfloat16).cuda().eval()
        x = torch.randn(32, 3, 16, 224, 224, device="cuda")
        nbytes = x.numel() * 4 + 32 * 1568 * 1024 * 2 + 1024 * 1536 * 2
    with torch.no_grad():
        y = tok(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters - 1):
            y = tok(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / max(args.iters - 1, 1)
    print(f"{args.config}: tokens {tuple(y.shape)} {y.dtype}, {ms * 1e3:.1f} us/call, algorithmic {nbytes / 1e6:.1f} MB -> "
          f"{nbytes / ms / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
