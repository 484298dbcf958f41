"""Seeded synthetic weights / tokens and the algorithmic FLOP count used by bench.py and the tools (SURVEY.md §8d:
"Synthetic inputs").  Kept in the package so that nothing on the measured path imports `oracle/`; the oracle has its own
copy of the generators and tests/test_api_surface.py checks that the two produce identical tensors.
"""
from __future__ import annotations

from typing import Dict

import torch


def make_encoder_state_dict(dim: int, depth: int, mlp_ratio: float = 4.0, seed: int = 0, qkv_bias: bool = True,
                            dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Flat state-dict of nn.Sequential(Block...) with the key layout of the published encoder checkpoints
    ("{i}.attn.qkv.weight", ...): clipped-normal(std .02) weights, N(0,.02) biases, LayerNorm gamma 1+N(0,.02) and beta
    N(0,.02), so that every bias / affine path carries signal."""
    g = torch.Generator().manual_seed(seed)
    hidden = int(dim * mlp_ratio)
    sd: Dict[str, torch.Tensor] = {}

    def w(*shape):
        return (torch.randn(*shape, generator=g) * 0.02).clamp_(-0.04, 0.04).to(dtype)

    def b(n):
        return (torch.randn(n, generator=g) * 0.02).to(dtype)

    for i in range(depth):
        p = f"{i}."
        sd[p + "norm1.weight"] = (1.0 + torch.randn(dim, generator=g) * 0.02).to(dtype)
        sd[p + "norm1.bias"] = b(dim)
        sd[p + "attn.qkv.weight"] = w(3 * dim, dim)
        if qkv_bias:
            sd[p + "attn.qkv.bias"] = b(3 * dim)
        sd[p + "attn.proj.weight"] = w(dim, dim)
        sd[p + "attn.proj.bias"] = b(dim)
        sd[p + "norm2.weight"] = (1.0 + torch.randn(dim, generator=g) * 0.02).to(dtype)
        sd[p + "norm2.bias"] = b(dim)
        sd[p + "mlp.fc1.weight"] = w(hidden, dim)
        sd[p + "mlp.fc1.bias"] = b(hidden)
        sd[p + "mlp.fc2.weight"] = w(dim, hidden)
        sd[p + "mlp.fc2.bias"] = b(dim)
    return sd


def make_tokens(B: int, N: int, dim: int, seed: int = 1234, dtype=torch.float32) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, N, dim, generator=g).to(dtype)


def encoder_flops(B: int, N: int, dim: int, depth: int, mlp_ratio: float = 4.0) -> float:
    """Algorithmic forward FLOPs B*N*L*(24 d^2 + 4 N d) for mlp_ratio 4."""
    hidden = dim * mlp_ratio
    per_token = 2 * (3 * dim * dim + dim * dim + 2 * dim * hidden) + 4 * N * dim
    return float(B) * N * depth * per_token
