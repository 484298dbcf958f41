"""Shared forward of the conv tokenizers: im2row gather kernel + tcgen05 GEMM with the bias in the epilogue."""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from .. import ops
from ..vision_transformer import _packed


def conv_tokens(module: nn.Module, conv: nn.Module, x5: torch.Tensor, kernel: Sequence[int], stride: Sequence[int],
                token_dtype: Optional[torch.dtype]) -> torch.Tensor:
    """x5 [B,C,T,H,W] -> tokens [B, Tp*Hp*Wp, d] (contiguous; the reference returns a transposed view,
    Data2Seq/Image.py:26).  conv.weight [d, C, (kt,) kh, kw] is used as the [d, K] GEMM operand."""
    if not x5.is_cuda:
        raise RuntimeError("Data2Seq tokenizers run on B200 GPUs only; there is no CPU fallback")
    if torch.is_grad_enabled() and (x5.requires_grad or conv.weight.requires_grad):
        raise NotImplementedError("tokenizer backward (SURVEY.md §8a E8) is not implemented yet; use torch.no_grad()")
    if x5.dtype not in (torch.float32, torch.bfloat16):
        x5 = x5.float()
    x5 = x5.contiguous()
    d = conv.weight.shape[0]
    w2 = _packed(module, "proj.weight2d", conv.weight, torch.bfloat16).view(d, -1)
    b = _packed(module, "proj.bias", conv.bias, torch.float32)
    rows, (Tp, Hp, Wp) = ops.im2row(x5, kernel, stride)
    tok = ops.linear(rows, w2, b)
    out_dtype = token_dtype or (x5.dtype if x5.dtype != torch.float64 else torch.float32)
    if out_dtype != torch.bfloat16:
        tok = ops.cast(tok, out_dtype)
    return tok.view(x5.shape[0], Tp * Hp * Wp, d)
