"""Shared forward of the conv tokenizers: im2row gather kernel + tcgen05 GEMM with the bias in the epilogue.

Optional fusion of the step right after the tokenizer: `pos_embed` (1, T, d) is added in the GEMM
epilogue when there are no prefix tokens (Video/models/modeling_finetune.py:434-436), or together with the cls / dist
`prefix_tokens` (1, k, d) in one assembly pass (Audio/src/models/ast_models.py:157-159).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from .. import ops
from ..vision_transformer import _packed


def _table(module: nn.Module, name: str, t: Optional[torch.Tensor], rows: int, d: int) -> Optional[torch.Tensor]:
    """(1, rows, d) / (rows, d) parameter -> cached contiguous bf16 [rows, d]."""
    if t is None:
        return None
    if t.shape[-1] != d or t.numel() != rows * d:
        raise ValueError(f"{name}: expected {rows} x {d} values, got shape {tuple(t.shape)}")
    return _packed(module, name, t, torch.bfloat16).view(rows, d)


def _forward_impl(module, weight, bias, pos, prefix, x5, kernel, stride, token_dtype):
    if x5.dtype not in (torch.float32, torch.bfloat16):
        x5 = x5.float()
    x5 = x5.contiguous()
    d = weight.shape[0]
    w2 = _packed(module, "proj.weight2d", weight, torch.bfloat16).view(d, -1)
    b = _packed(module, "proj.bias", bias, torch.float32)
    rows, (Tp, Hp, Wp) = ops.im2row(x5, kernel, stride)
    B, n = x5.shape[0], Tp * Hp * Wp
    k = 0 if prefix is None else prefix.numel() // d
    pos2 = _table(module, "pos_embed", pos, n + k, d)
    pre2 = _table(module, "prefix_tokens", prefix, k, d)
    if k == 0 and pos2 is not None:
        tok = ops.linear(rows, w2, b, residual=pos2, residual_period=n).view(B, n, d)
    else:
        tok = ops.linear(rows, w2, b).view(B, n, d)
        if k > 0:
            tok = ops.assemble_tokens(tok, pre2, pos2)
    out_dtype = token_dtype or (x5.dtype if x5.dtype != torch.float64 else torch.float32)
    if out_dtype != torch.bfloat16:
        tok = ops.cast(tok, out_dtype)
    return tok, rows, k


def conv_tokens(module: nn.Module, conv: nn.Module, x5: torch.Tensor, kernel: Sequence[int], stride: Sequence[int],
                token_dtype: Optional[torch.dtype], pos_embed: Optional[torch.Tensor] = None,
                prefix_tokens: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x5 [B,C,T,H,W] -> tokens [B, k + Tp*Hp*Wp, d] (contiguous; the reference returns a transposed view,
    Data2Seq/Image.py:26).  conv.weight [d, C, (kt,) kh, kw] is used as the [d, K] GEMM operand."""
    if not x5.is_cuda:
        raise RuntimeError("Data2Seq tokenizers run on B200 GPUs only; there is no CPU fallback")
    if torch.is_grad_enabled() and x5.requires_grad:
        raise NotImplementedError("gradient with respect to the tokenizer INPUT is not implemented (the reference "
                                  "pipelines never need it); weight / bias / pos_embed / prefix-token gradients are")
    params = (conv.weight, conv.bias, pos_embed, prefix_tokens)
    if torch.is_grad_enabled() and any(p is not None and p.requires_grad for p in params):
        return _ConvTokensFn.apply(conv.weight, conv.bias, pos_embed, prefix_tokens, module, x5, tuple(kernel),
                                   tuple(stride), token_dtype)
    return _forward_impl(module, conv.weight, conv.bias, pos_embed, prefix_tokens, x5, kernel, stride, token_dtype)[0]


class _ConvTokensFn(torch.autograd.Function):
    """Trainable tokenizer under a (frozen or not) encoder (Audio/src/models/ast_models.py:66-70): forward = im2row +
    GEMM, backward = weight gradient dW = dTokens^T @ rows (split-K tcgen05 GEMM on transposed copies), bias gradient,
    and the batch-summed gradient of the positional table / prefix tokens."""

    @staticmethod
    def forward(ctx, weight, bias, pos, prefix, module, x5, kernel, stride, token_dtype):
        tok, rows, k = _forward_impl(module, weight, bias, pos, prefix, x5, kernel, stride, token_dtype)
        ctx.rows, ctx.k, ctx.wshape, ctx.wdtype = rows, k, weight.shape, weight.dtype
        ctx.bdtype = bias.dtype if bias is not None else None
        ctx.pos_meta = None if pos is None else (pos.shape, pos.dtype)
        ctx.pre_meta = None if prefix is None else (prefix.shape, prefix.dtype)
        return tok

    @staticmethod
    def backward(ctx, g):
        d, k = ctx.wshape[0], ctx.k
        B, T, _ = g.shape
        g3 = g.contiguous()
        g3 = g3 if g3.dtype == torch.bfloat16 else ops.cast(g3, torch.bfloat16)
        dpos = dpre = None
        if (ctx.pos_meta is not None and ctx.needs_input_grad[2]) or (ctx.pre_meta is not None and ctx.needs_input_grad[3]):
            gsum = ops.colsum(g3.view(B, T * d))                          # fp32 [T*d]: gradient summed over the batch
            if ctx.pos_meta is not None and ctx.needs_input_grad[2]:
                dpos = gsum.view(ctx.pos_meta[0])
                dpos = dpos if ctx.pos_meta[1] == torch.float32 else ops.cast(dpos, ctx.pos_meta[1])
            if ctx.pre_meta is not None and ctx.needs_input_grad[3]:
                dpre = gsum[:k * d].contiguous().view(ctx.pre_meta[0])
                dpre = dpre if ctx.pre_meta[1] == torch.float32 else ops.cast(dpre, ctx.pre_meta[1])
        g2 = (g3[:, k:, :].contiguous() if k > 0 else g3).view(-1, d)     # strided copy only when prefix rows exist
        dw = db = None
        if ctx.needs_input_grad[0]:
            dw = ops.linear_wgrad(g2, ctx.rows)                           # [d, K] fp32
            dw = (dw if dw.dtype == ctx.wdtype else ops.cast(dw, ctx.wdtype)).view(ctx.wshape)
        if ctx.bdtype is not None and ctx.needs_input_grad[1]:
            db = ops.colsum(g2)
            db = db if ctx.bdtype == torch.float32 else ops.cast(db, ctx.bdtype)
        ctx.rows = None
        return dw, db, dpos, dpre, None, None, None, None, None
