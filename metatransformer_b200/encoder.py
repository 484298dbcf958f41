"""Whole-encoder fast path: one C call (`mt_encoder_forward`) for nn.Sequential(*[Block]*L).

`build_encoder("base" | "large")` returns exactly the module the reference README builds (README.md:125-147), made of
this package's drop-in Blocks; `FusedEncoder(seq)` wraps such a Sequential, packs its weights to bf16 once and runs
all L blocks back to back on the current stream with a cached workspace (no per-op Python, no torch allocations).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import BlockWeights, check
from .vision_transformer import Block, _packed, _layernorm_eps

ENCODER_CONFIGS = {
    # README.md:125-134 / :138-147
    "base": dict(dim=768, num_heads=12, depth=12),
    "large": dict(dim=1024, num_heads=16, depth=24),
}


def build_encoder(scale: str = "base", **overrides) -> nn.Sequential:
    cfg = dict(ENCODER_CONFIGS[scale])
    cfg.update(overrides)
    depth = cfg.pop("depth")
    return nn.Sequential(*[
        Block(dim=cfg["dim"], num_heads=cfg["num_heads"], mlp_ratio=4., qkv_bias=True, norm_layer=nn.LayerNorm,
              act_layer=nn.GELU) for _ in range(depth)])


class FusedEncoder(nn.Module):
    """Runs a Sequential of `Block`s through the C-ABI whole-encoder driver (inference / frozen-encoder forward)."""

    def __init__(self, blocks: nn.Sequential, fold_layernorm: bool = True, use_cuda_graph: bool = False,
                 precision: str = "auto"):
        super().__init__()
        if precision not in ("auto", "bf16", "fp32"):
            raise ValueError("FusedEncoder: precision must be 'auto', 'bf16' or 'fp32'")
        # "auto" follows the reference's dtype rules (see Block.precision): fp32 tensors outside torch.autocast take the
        # fp32-accurate path (3xTF32 GEMMs, 1e-4 gate), everything else the bf16 tensor-core path (1e-2 gate)
        self.precision = precision
        self.fold_layernorm = fold_layernorm
        # replay the 84 / 168 launches of one forward from a captured CUDA graph (keyed by the buffers involved);
        # only pays off when the caller re-uses its input / output buffers
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}
        if len(blocks) == 0 or not all(isinstance(b, Block) for b in blocks):
            raise TypeError("FusedEncoder needs a non-empty nn.Sequential of metatransformer_b200 Blocks")
        self.blocks = blocks
        self._handle = None
        self._handle_key = None
        self._keepalive = None
        self._ws = None

    # -- handle management --------------------------------------------------------------------------------------
    def _weights_key(self):
        return tuple((p.data_ptr(), p._version, str(p.device), p.dtype) for p in self.blocks.parameters())

    def _ensure_handle(self):
        key = (self._weights_key(), self.fold_layernorm)
        if self._handle is not None and key == self._handle_key:
            return
        self._destroy()
        lib = _lib.load()
        b0 = self.blocks[0]
        dim = b0.norm1.normalized_shape[0]
        heads = b0.attn.num_heads
        hidden = b0.mlp.fc1.out_features
        eps = _layernorm_eps(b0.norm1)
        arr = (BlockWeights * len(self.blocks))()
        keep = []
        for i, blk in enumerate(self.blocks):
            if blk.attn.num_heads != heads or blk.mlp.fc1.out_features != hidden or _layernorm_eps(blk.norm1) != eps \
                    or _layernorm_eps(blk.norm2) != eps or abs(blk.attn.scale - (dim // heads) ** -0.5) > 1e-12:
                raise NotImplementedError("FusedEncoder: all blocks must share one configuration and the default scale")

            def P(mod, name, param, dt):
                t = _packed(mod, name, param, dt)
                keep.append(t)
                return None if t is None else t.data_ptr()

            w = arr[i]
            if self.fold_layernorm:
                wq, csq, bq = blk._folded("fold.qkv", blk.attn.qkv, blk.norm1)
                w1, cs1, b1 = blk._folded("fold.fc1", blk.mlp.fc1, blk.norm2)
                keep.extend([wq, csq, bq, w1, cs1, b1])
                w.qkv_w, w.qkv_b, w.qkv_colsum = wq.data_ptr(), bq.data_ptr(), csq.data_ptr()
                w.fc1_w, w.fc1_b, w.fc1_colsum = w1.data_ptr(), b1.data_ptr(), cs1.data_ptr()
            else:
                w.ln1_w = P(blk, "norm1.weight", blk.norm1.weight, torch.float32)
                w.ln1_b = P(blk, "norm1.bias", blk.norm1.bias, torch.float32)
                w.qkv_w = P(blk.attn, "qkv.weight", blk.attn.qkv.weight, torch.bfloat16)
                w.qkv_b = P(blk.attn, "qkv.bias", blk.attn.qkv.bias, torch.float32)
                w.ln2_w = P(blk, "norm2.weight", blk.norm2.weight, torch.float32)
                w.ln2_b = P(blk, "norm2.bias", blk.norm2.bias, torch.float32)
                w.fc1_w = P(blk.mlp, "fc1.weight", blk.mlp.fc1.weight, torch.bfloat16)
                w.fc1_b = P(blk.mlp, "fc1.bias", blk.mlp.fc1.bias, torch.float32)
            w.proj_w = P(blk.attn, "proj.weight", blk.attn.proj.weight, torch.bfloat16)
            w.proj_b = P(blk.attn, "proj.bias", blk.attn.proj.bias, torch.float32)
            w.fc2_w = P(blk.mlp, "fc2.weight", blk.mlp.fc2.weight, torch.bfloat16)
            w.fc2_b = P(blk.mlp, "fc2.bias", blk.mlp.fc2.bias, torch.float32)
        h = ctypes.c_void_p()
        check(lib.mt_encoder_create(ctypes.byref(h), dim, heads, hidden, len(self.blocks), float(eps), arr),
              "mt_encoder_create")
        self._handle, self._handle_key, self._keepalive = h, key, keep
        self.dim = dim

    def _destroy(self):
        if self._handle is not None:
            _lib.load().mt_encoder_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    @property
    def launches_per_forward(self) -> int:
        self._ensure_handle()
        return _lib.load().mt_encoder_launches_per_forward(self._handle)

    # -- forward --------------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("FusedEncoder: CUDA (B200) tensors only; there is no CPU fallback")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.blocks.parameters())):
            raise NotImplementedError("FusedEncoder is forward-only; wrap the call in torch.no_grad()")
        if self.training and any(m.p > 0 for m in self.blocks.modules() if isinstance(m, nn.Dropout)) or any(
                getattr(b, "_is_stochastic", lambda: False)() for b in self.blocks):
            raise NotImplementedError("FusedEncoder: dropout / drop-path in training mode is not supported; call the "
                                      "Blocks (nn.Sequential) for training")
        B, N, C = x.shape
        if self.precision == "fp32" or (self.precision == "auto" and x.dtype == torch.float32
                                        and not torch.is_autocast_enabled()):
            x2 = x.reshape(-1, C).float()
            x2 = x2 if x2.is_contiguous() else x2.contiguous()
            for blk in self.blocks:
                x2 = blk._forward_fp32(x2, B, N)
            y = x2.view(B, N, C)
            if out is not None:
                out.copy_(y)
                return out
            return y if x.dtype == torch.float32 else ops.cast(y, x.dtype)
        self._ensure_handle()
        lib = _lib.load()
        if C != self.dim:
            raise ValueError(f"FusedEncoder: expected last dim {self.dim}, got {C}")
        xin = x.contiguous()
        if xin.dtype not in (torch.bfloat16, torch.float32):
            xin = ops.cast(xin, torch.bfloat16)          # fp16 callers: bf16 has the range, the stream is split inside
        io_dt = xin.dtype                                  # bf16 or fp32, handed to the driver as is
        if out is not None:
            if out.dtype != io_dt or tuple(out.shape) != tuple(xin.shape) or out.device != xin.device \
                    or not out.is_contiguous():
                raise ValueError(f"FusedEncoder: out must be a contiguous {io_dt} tensor of shape {tuple(xin.shape)} on "
                                 f"{xin.device} (got {out.dtype}, {tuple(out.shape)}, {out.device})")
            y = out
        else:
            y = torch.empty_like(xin)
        xb = xin
        dt_code = _lib.MT_DT_BF16 if io_dt == torch.bfloat16 else _lib.MT_DT_F32
        need = ctypes.c_size_t()
        check(lib.mt_encoder_workspace_bytes(self._handle, B, N, ctypes.byref(need)), "mt_encoder_workspace_bytes")
        if self._ws is None or self._ws.numel() < need.value or self._ws.device != x.device:
            self._ws = torch.empty(need.value, dtype=torch.uint8, device=x.device)
            self._graphs.clear()

        def launch():
            check(lib.mt_encoder_forward_ex(self._handle, xb.data_ptr(), dt_code, y.data_ptr(), dt_code, B, N,
                                            self._ws.data_ptr(), self._ws.numel(),
                                            torch.cuda.current_stream().cuda_stream), "mt_encoder_forward_ex")

        with torch.cuda.device(x.device):
            if not self.use_cuda_graph or xb is not x:       # a per-call staging copy would be re-captured every call
                launch()
            else:
                key = (xb.data_ptr(), y.data_ptr(), B, N, io_dt, self._handle_key)
                g = self._graphs.get(key)
                if g is None:
                    launch()                         # warm-up outside capture (function attributes, lazy module load)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        launch()
                    if len(self._graphs) > 16:
                        self._graphs.clear()
                    self._graphs[key] = g
                g.replay()
        return y if y.dtype == x.dtype else ops.cast(y, x.dtype)
