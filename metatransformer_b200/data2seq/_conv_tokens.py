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
    if torch.is_grad_enabled() and x5.requires_grad:
        raise NotImplementedError("gradient with respect to the tokenizer INPUT is not implemented (the reference "
                                  "pipelines never need it); weight / bias gradients are")
    if torch.is_grad_enabled() and (conv.weight.requires_grad or (conv.bias is not None and conv.bias.requires_grad)):
        return _ConvTokensFn.apply(conv.weight, conv.bias, module, x5, tuple(kernel), tuple(stride), token_dtype)
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


class _ConvTokensFn(torch.autograd.Function):
    """Trainable tokenizer under a (frozen or not) encoder (Audio/src/models/ast_models.py:66-70): forward = im2row +
    GEMM, backward = weight gradient dW = dTokens^T @ rows (TN GEMM on transposed copies) and bias gradient."""

    @staticmethod
    def forward(ctx, weight, bias, module, x5, kernel, stride, token_dtype):
        if x5.dtype not in (torch.float32, torch.bfloat16):
            x5 = x5.float()
        x5 = x5.contiguous()
        d = weight.shape[0]
        w2 = _packed(module, "proj.weight2d", weight, torch.bfloat16).view(d, -1)
        b = _packed(module, "proj.bias", bias, torch.float32)
        rows, (Tp, Hp, Wp) = ops.im2row(x5, kernel, stride)
        tok = ops.linear(rows, w2, b)
        out_dtype = token_dtype or x5.dtype
        ctx.rows, ctx.wshape, ctx.has_bias, ctx.wdtype = rows, weight.shape, bias is not None, weight.dtype
        ctx.bdtype = bias.dtype if bias is not None else None
        if out_dtype != torch.bfloat16:
            tok = ops.cast(tok, out_dtype)
        return tok.view(x5.shape[0], Tp * Hp * Wp, d)

    @staticmethod
    def backward(ctx, g):
        d = ctx.wshape[0]
        g2 = g.reshape(-1, d)
        g2 = g2.contiguous() if g2.dtype == torch.bfloat16 else ops.cast(g2.contiguous(), torch.bfloat16)
        dw = ops.linear_wgrad(g2, ctx.rows)                       # [d, K] fp32 (split-K tcgen05 GEMM over the tokens)
        dw = dw if dw.dtype == ctx.wdtype else ops.cast(dw, ctx.wdtype)
        db = None
        if ctx.has_bias:
            db = ops.colsum(g2)
            db = db if ctx.bdtype == torch.float32 else ops.cast(db, ctx.bdtype)
        ctx.rows = None
        return dw.view(ctx.wshape), db, None, None, None, None, None
