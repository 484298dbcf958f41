"""Throughput of the BASELINE.json configurations C2..C5 on one GPU (per-GPU shapes), encoder and tokenizer timed
separately with CUDA events.  Prints one JSON line per configuration (for BASELINE.md section 4).

    python tools/bench_configs.py [C2 C3 C4 C5]
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import metatransformer_b200 as mt  # noqa: E402
from metatransformer_b200 import synthetic as orc  # noqa: E402  (seeded weights + FLOP formula)

CFG = {
    "C2": dict(scale="base", B=256, N=196, tok=None),
    "C3": dict(scale="large", B=128, N=256, tok=None),
    "C4": dict(scale="base", B=64, N=512, tok="audio"),      # fbank (64, 1024, 128), Conv2d(1,768,16,s=16)
    "C5": dict(scale="large", B=32, N=1568, tok="video"),    # clip (32, 3, 16, 224, 224), Conv3d tubelets
}


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main(names):
    peaks = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())
    for name in names:
        c = CFG[name]
        cfg = mt.ENCODER_CONFIGS[c["scale"]]
        d, depth, heads = cfg["dim"], cfg["depth"], cfg["num_heads"]
        enc = mt.build_encoder(c["scale"])
        enc.load_state_dict(orc.make_encoder_state_dict(d, depth, seed=0), strict=True)
        enc = enc.cuda().eval()
        fused = mt.FusedEncoder(enc, use_cuda_graph=True)
        B, N = c["B"], c["N"]
        g = torch.Generator().manual_seed(1234)
        xs = [torch.randn(B, N, d, generator=g).bfloat16().cuda() for _ in range(3)]
        y = torch.empty_like(xs[0])
        i = [0]

        def step():
            fused(xs[i[0] % 3], out=y)
            i[0] += 1

        with torch.no_grad():
            ms = timeit(step, 20)
            flops = orc.encoder_flops(B, N, d, depth)
            line = {"cfg": name, "model": c["scale"], "B": B, "N": N, "d": d, "ms_per_fwd": ms,
                    "tokens_per_s": B * N / ms * 1e3, "tflops": flops / ms / 1e9,
                    "frac_of_sustained_peak": flops / ms / 1e9 / peaks["bf16_tflops_sustained"]}
            if c["tok"] == "audio":
                tok = mt.data2seq.Acoustic.PatchEmbed(embed_dim=d, fstride=16, tstride=16, token_dtype=torch.bfloat16).cuda().eval()
                fb = torch.randn(B, 1024, 128, generator=g).cuda()
                fn = lambda: tok(fb.unsqueeze(1).transpose(2, 3).contiguous())  # noqa: E731
                assert fn().shape == (B, 512, d)
                in_bytes, K = fb.numel() * 4, 256
            elif c["tok"] == "video":
                tok = mt.data2seq.Video.PatchEmbed(embed_dim=d, token_dtype=torch.bfloat16).cuda().eval()
                clip = torch.randn(B, 3, 16, 224, 224, generator=g).cuda()
                fn = lambda: tok(clip)  # noqa: E731
                assert fn().shape == (B, 1568, d)
                in_bytes, K = clip.numel() * 4, 1536
            else:
                fn = None
            if fn is not None:
                tms = timeit(fn, 10)
                alg_bytes = in_bytes + B * N * d * 2 + (K * d + d) * 2
                line.update({"tokenizer_ms": tms, "tokenizer_alg_GBps": alg_bytes / tms / 1e6,
                             "tokenizer_frac_of_hbm_peak": alg_bytes / tms / 1e6 / peaks["hbm_gbs"],
                             "e2e_tokens_per_s": B * N / (ms + tms) * 1e3})
        print(json.dumps(line), flush=True)
        del enc, fused, xs, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main(sys.argv[1:] or ["C2", "C3", "C4", "C5"])
