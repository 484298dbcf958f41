"""The step after the encoder stack, on the sm_100a kernels: final LayerNorm fused with the pooling, then the linear head.

Every reference fork ends its encoder the same way with a different pooling:

    pool="cls"       norm(x)[:, 0]                      Video/models/modeling_finetune.py:449,
                                                        Hyper-spectrum/metatransformer.py:163-165 (mlp_head = LN + Linear)
    pool="mean"      fc_norm(x.mean(1))                 Video/models/modeling_finetune.py:446-447 (use_mean_pooling)
    pool="cls_dist"  (norm(x)[:, 0] + norm(x)[:, 1])/2  Audio/src/models/ast_models.py:163-164

`ClassifierHead` keeps the parameter names the forks use (`norm`, `head`), trains through autograd (input gradient for
the encoder below it, LayerNorm and Linear parameter gradients) and has no eager fallback.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .vision_transformer import _packed

__all__ = ["ClassifierHead", "pool_norm"]

_MODES = {"cls": ops.POOL_FIRST, "mean": ops.POOL_MEAN, "cls_dist": ops.POOL_FIRST_TWO}


def _bf16_3d(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("metatransformer_b200 heads run on B200 GPUs only; there is no CPU fallback")
    if x.dim() != 3:
        raise ValueError(f"expected (B, N, C) tokens, got shape {tuple(x.shape)}")
    x = x.contiguous()
    return x if x.dtype == torch.bfloat16 else ops.cast(x, torch.bfloat16)


def _padded_head(module: nn.Module, lin: nn.Linear):
    """bf16 [Np, d] weight / fp32 [Np] bias with the class count rounded up to the GEMM's 32-column granule, and the
    [d, Np] transpose for the input gradient; rebuilt when the parameters change."""
    cache = module.__dict__.setdefault("_mt_pack_cache", {})
    params = (lin.weight, lin.bias)
    key = tuple(None if q is None else (q.data_ptr(), q._version, q.device, q.dtype) for q in params)
    hit = cache.get("head")
    if hit is not None and hit[0] == key:
        return hit[1]
    C, d = lin.weight.shape
    Np = (C + 31) // 32 * 32
    w = torch.zeros((Np, d), dtype=torch.float32, device=lin.weight.device)          # one-time weight packing
    w[:C] = lin.weight.detach().float()
    b = torch.zeros(Np, dtype=torch.float32, device=lin.weight.device)
    if lin.bias is not None:
        b[:C] = lin.bias.detach().float()
    wb = ops.cast(w, torch.bfloat16)
    packed = (wb, b, ops.transpose(wb))
    cache["head"] = (key, packed)
    return packed


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, module, norm_w, norm_b, head_w, head_b):
        xb = _bf16_3d(x.detach())
        norm, mode = module.norm, _MODES[module.pool]
        g = _packed(module, "norm.weight", norm.weight, torch.float32)
        b = _packed(module, "norm.bias", norm.bias, torch.float32)
        feat = ops.pool_layernorm(xb, g, b, mode, norm.eps)                           # [B, d] bf16
        ctx.module, ctx.xb, ctx.feat, ctx.in_dtype = module, xb, feat, x.dtype
        if head_w is None:
            return feat if x.dtype == torch.bfloat16 else ops.cast(feat, x.dtype)
        wb, bias, _ = _padded_head(module, module.head)
        logits = ops.linear(feat, wb, bias)                                           # [B, Np]
        logits = logits if x.dtype == torch.bfloat16 else ops.cast(logits, x.dtype)
        return logits[:, :head_w.shape[0]]

    @staticmethod
    def backward(ctx, gy):
        module, xb, feat = ctx.module, ctx.xb, ctx.feat
        norm, mode = module.norm, _MODES[module.pool]
        B, N, d = xb.shape
        dhw = dhb = None
        if isinstance(module.head, nn.Linear):
            C = module.head.weight.shape[0]
            wb, _, wbt = _padded_head(module, module.head)
            Np = wb.shape[0]
            gp = torch.zeros((B, Np), dtype=torch.bfloat16, device=xb.device)          # class padding of the logits grad
            gp[:, :C] = gy
            dfeat = ops.linear(gp, wbt)                                               # [B, d] = dY W
            if ctx.needs_input_grad[4]:
                dhw = ops.linear_wgrad(gp, feat)[:C]
                dhw = dhw if dhw.dtype == module.head.weight.dtype else ops.cast(dhw.contiguous(), module.head.weight.dtype)
            if module.head.bias is not None and ctx.needs_input_grad[5]:
                dhb = ops.colsum(gp)[:C]
                dhb = dhb if dhb.dtype == module.head.bias.dtype else ops.cast(dhb.contiguous(), module.head.bias.dtype)
        else:
            dfeat = gy.contiguous()
            dfeat = dfeat if dfeat.dtype == torch.bfloat16 else ops.cast(dfeat, torch.bfloat16)
        want_params = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        dx, dg, db = ops.pool_layernorm_bwd(dfeat, xb, _packed(module, "norm.weight", norm.weight, torch.float32), mode,
                                            norm.eps, param_grads=want_params)
        if ctx.needs_input_grad[0]:
            dx = dx if ctx.in_dtype == torch.bfloat16 else ops.cast(dx, ctx.in_dtype)
        else:
            dx = None
        if want_params:
            dg = dg if dg.dtype == norm.weight.dtype else ops.cast(dg, norm.weight.dtype)
            db = db if db.dtype == norm.bias.dtype else ops.cast(db, norm.bias.dtype)
        ctx.xb = ctx.feat = None
        return dx, None, dg, db, dhw, dhb


class ClassifierHead(nn.Module):
    """norm -> pool -> head.  `num_classes=0` returns the pooled, normalised feature (B, dim)."""

    def __init__(self, dim: int, num_classes: int, pool: str = "cls", norm_layer=nn.LayerNorm):
        super().__init__()
        if pool not in _MODES:
            raise ValueError(f"pool must be one of {sorted(_MODES)}, got {pool!r}")
        self.pool = pool
        self.norm = norm_layer(dim)
        if not isinstance(self.norm, nn.LayerNorm) or not self.norm.elementwise_affine:
            raise NotImplementedError("ClassifierHead: only an affine nn.LayerNorm has a B200 kernel")
        self.head = nn.Linear(dim, num_classes) if num_classes > 0 else nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        lin = self.head if isinstance(self.head, nn.Linear) else None
        if self.pool == "cls_dist" and x.shape[1] < 2:
            raise ValueError("pool='cls_dist' needs at least two tokens")
        return _HeadFn.apply(x, self, self.norm.weight, self.norm.bias, None if lin is None else lin.weight,
                             None if lin is None else lin.bias)


def pool_norm(x: torch.Tensor, norm: nn.LayerNorm, pool: str = "cls") -> torch.Tensor:
    """Functional, inference-only form: the pooled + normalised (B, dim) feature of tokens x (B, N, dim)."""
    if torch.is_grad_enabled() and (x.requires_grad or norm.weight.requires_grad):
        raise NotImplementedError("pool_norm is forward-only; use ClassifierHead for training")
    xb = _bf16_3d(x)
    feat = ops.pool_layernorm(xb, norm.weight.detach().float().contiguous(), norm.bias.detach().float().contiguous(),
                              _MODES[pool], norm.eps)
    return feat if x.dtype == torch.bfloat16 else ops.cast(feat, x.dtype)
