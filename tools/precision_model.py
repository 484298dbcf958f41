#!/usr/bin/env python
"""CPU emulation of the B200 path's precision model (which tensors are rounded to bf16, which stay fp32), used to
choose the residual-stream format before spending GPU time:  python tools/precision_model.py [base|large] [N] [B]

Emulated pipeline per Block (fp32 accumulation everywhere, like the tcgen05 kernels):
  A operand of qkv / fc1 = bf16(x)  (LayerNorm folded: stats from the fp32 values, W' = bf16(W * gamma))
  qkv, attention output, MLP hidden = bf16;  P = bf16(exp2(...)) for the P.V product
  residual stream = "bf16" (round after every residual add) or "split" (hi + lo bf16 pair = 16 mantissa bits)
Compared against the fp32 oracle on unrounded and on bf16-rounded weights.  Development tool (imports oracle/).
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import encoder_oracle as orc  # noqa: E402


def r16(t):
    return t.bfloat16().float()


def split(t):
    hi = r16(t)
    return hi + r16(t - hi)


def emulate(x, sd, depth, heads, stream, eps=1e-5):
    rs = {"bf16": r16, "split": split, "fp32": lambda t: t}[stream]
    x = rs(x)
    for i in range(depth):
        g = lambda k: sd[f"{i}." + k]  # noqa: E731
        for branch in ("attn", "mlp"):
            gamma, beta = (g("norm1.weight"), g("norm1.bias")) if branch == "attn" else (g("norm2.weight"), g("norm2.bias"))
            lin = "attn.qkv" if branch == "attn" else "mlp.fc1"
            w, b = g(lin + ".weight"), g(lin + ".bias")
            wf = r16(w * gamma)                       # folded weight, rounded
            colsum = wf.sum(1)
            bf = b + w @ beta
            mean = x.mean(-1, keepdim=True)
            var = ((x - mean) ** 2).mean(-1, keepdim=True)
            rstd = torch.rsqrt(var + eps)
            acc = r16(x) @ wf.t()
            y = rstd * (acc - mean * colsum) + bf
            if branch == "attn":
                B, N, C = x.shape
                dh = C // heads
                qkv = r16(y).reshape(B, N, 3, heads, dh).permute(2, 0, 3, 1, 4)
                q, k, v = qkv[0], qkv[1], qkv[2]
                s = (q @ k.transpose(-2, -1)) * dh ** -0.5
                p = (s - s.max(-1, keepdim=True).values).exp()
                l = p.sum(-1, keepdim=True)
                o = r16((r16(p) @ v) / l).transpose(1, 2).reshape(B, N, C)
                out = o @ r16(g("attn.proj.weight")).t() + g("attn.proj.bias")
            else:
                hid = r16(orc.gelu_erf(y))
                out = hid @ r16(g("mlp.fc2.weight")).t() + g("mlp.fc2.bias")
            x = rs(x + out)
    return x


def main():
    scale = sys.argv[1] if len(sys.argv) > 1 else "base"
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 196
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    dim, heads, depth = (768, 12, 12) if scale == "base" else (1024, 16, 24)
    torch.manual_seed(0)
    sd = orc.make_encoder_state_dict(dim, depth, seed=0)
    sd16 = {k: (r16(v) if v.dim() == 2 else v) for k, v in sd.items()}
    x = r16(orc.make_tokens(B, N, dim, seed=1234))
    with torch.no_grad():
        ref = orc.encoder(x.double(), {k: v.double() for k, v in sd.items()}, depth, heads).float()
        ref16 = orc.encoder(x.double(), {k: v.double() for k, v in sd16.items()}, depth, heads).float()
        for stream in ("bf16", "split", "fp32"):
            got = emulate(x, sd, depth, heads, stream)
            for name, r in (("fp32 weights", ref), ("bf16-rounded weights", ref16)):
                e = got - r
                print(f"{scale} L={depth} N={N} stream={stream:5s} vs oracle on {name:20s}: rel_l2={float(e.norm() / r.norm()):.3e} "
                      f"max_abs/max|ref|={float(e.abs().max() / max(1.0, float(r.abs().max()))):.3e} (max|ref|={float(r.abs().max()):.2f})")
            gb = r16(got)
            e = gb - ref
            print(f"    ... output rounded to bf16: rel_l2={float(e.norm() / ref.norm()):.3e} "
                  f"max_abs/max|ref|={float(e.abs().max() / max(1.0, float(ref.abs().max()))):.3e}")


if __name__ == "__main__":
    main()
