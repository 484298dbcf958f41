#!/usr/bin/env python
"""Two plain (no CUDA graph) forwards of a shallow fused encoder at a BASELINE shape — the target of ncu captures:
    ncu --set full -k regex:'gemm_bf16|attention_fwd' -s 15 -c 5 python tools/one_forward.py --config c2
With depth 2 the filtered launches 15-19 are layer 1's qkv (LayerNorm statistics from the fc2 epilogue), attention, proj,
fc1, fc2 of the SECOND forward (weights packed, function attributes set)."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
CONFIGS = {"c2": ("base", 256, 196), "c3": ("large", 128, 256), "c4": ("base", 64, 512), "c5": ("large", 32, 1568)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--depth", type=int, default=2)
    args = ap.parse_args()
    import torch
    import metatransformer_b200 as mt
    from metatransformer_b200 import synthetic
    scale, B, N = CONFIGS[args.config]
    dim = 768 if scale == "base" else 1024
    enc = mt.build_encoder(scale, depth=args.depth)
    enc.load_state_dict(synthetic.make_encoder_state_dict(dim, args.depth, seed=0), strict=True)
    enc = enc.cuda().eval()
    fused = mt.FusedEncoder(enc)
    x = synthetic.make_tokens(B, N, dim, seed=1234).bfloat16().cuda()
    with torch.no_grad():
        for _ in range(2):
            y = fused(x)
    torch.cuda.synchronize()
    print("ok", tuple(y.shape), float(y.float().abs().mean()))


if __name__ == "__main__":
    main()
