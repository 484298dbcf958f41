"""timm-compatible `Mlp`, `Attention`, `Block` whose forward runs on hand-written sm_100a kernels (libmt_b200.so).

Drop-in for `from timm.models.vision_transformer import Block` as the reference uses it (README.md:116-149 and the
eleven call sites listed in SURVEY.md §2.1 row 2): same constructor signature as timm 0.4.12 (plus timm 0.8's
`init_values=None` for Tabular/README.md:31), same attribute / parameter names, hence the same state-dict keys
("{i}.norm1.weight", "{i}.attn.qkv.weight", ...) so `encoder.load_state_dict(ckpt, strict=True)`
works on the published checkpoints.  Reference math: PointCloud/openpoints/models/layers/attention.py:12-58,
mlp.py:11-35 (verbatim timm copies).

Precision model (documented in DESIGN.md): GEMM / attention operands bf16, fp32 accumulation, fp32 LayerNorm and
softmax statistics; inside a Block the residual stream is a split hi + lo pair of bf16 tensors (16 mantissa bits).
bf16 / fp16 inputs give a bf16-rounded Block output in the input dtype; fp32 inputs keep the split stream across
Blocks (fp32 in, fp32 out, like the reference under autocast: fp32 residual, reduced-precision matmul operands).  There is NO CPU or eager fallback: CPU tensors, non-B200 devices or a missing
extension raise.
"""
from __future__ import annotations

from functools import partial
from typing import Optional

import torch
import torch.nn as nn

from . import ops

__all__ = ["Mlp", "Attention", "Block", "DropPath"]


class DropPath(nn.Module):
    """Stochastic depth per sample (timm.models.layers.DropPath).  Identity in eval mode / at rate 0."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        return _DropPathFn.apply(x, self.drop_prob)


class _DropPathFn(torch.autograd.Function):
    """Stand-alone stochastic depth: out = x * bernoulli(keep)[sample] / keep (timm.models.layers.drop_path), one pass of
    the mt_dropout kernel with p = 0 and a per-sample row scale; the backward applies the same scale to the gradient.
    (Inside Block the same kernel also performs the residual add.)"""

    @staticmethod
    def forward(ctx, x, drop_prob):
        if not x.is_cuda:
            raise RuntimeError("DropPath: CUDA (B200) tensors only; there is no CPU fallback")
        keep = 1.0 - drop_prob
        B = x.shape[0]
        rs = torch.empty(B, dtype=torch.float32, device=x.device).bernoulli_(keep).div_(keep)
        ctx.rs, ctx.rows_per = rs, x.numel() // B // x.shape[-1]
        return _DropPathFn._scale(x, rs, ctx.rows_per)

    @staticmethod
    def _scale(t, rs, rows_per):
        t2 = t.reshape(-1, t.shape[-1])
        t2 = t2 if t2.is_contiguous() else t2.contiguous()
        tb = t2 if t2.dtype == torch.bfloat16 else ops.cast(t2, torch.bfloat16)
        out = ops.dropout(tb, 0.0, 0, 0, row_scale=rs, rows_per_scale=rows_per)
        out = out if t.dtype == torch.bfloat16 else ops.cast(out, t.dtype)
        return out.view(t.shape)

    @staticmethod
    def backward(ctx, g):
        return _DropPathFn._scale(g, ctx.rs, ctx.rows_per), None


def _philox_reserve(device: torch.device, n_elements: int):
    """(seed, offset) for a mask over n_elements, drawn from torch's CUDA generator of `device` so that
    torch.cuda.manual_seed / torch.utils.checkpoint's RNG stashing govern these masks exactly like nn.Dropout's."""
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed, offset = gen.initial_seed(), gen.get_offset()
    gen.set_offset(offset + (n_elements // 8 + 4) // 4 * 4)
    return seed, offset


# ------------------------------------------------------------------------------------------------------------------
# parameter packing: bf16 matrices / fp32 vectors derived from the nn.Parameters, rebuilt lazily whenever the
# parameter storage, version or device changes (nn.DataParallel replicas, .to(), load_state_dict, optimizer steps)
# ------------------------------------------------------------------------------------------------------------------
def _packed(module: nn.Module, name: str, param: Optional[torch.Tensor], dtype: torch.dtype) -> Optional[torch.Tensor]:
    if param is None:
        return None
    cache = module.__dict__.setdefault("_mt_pack_cache", {})
    key = (param.data_ptr(), param._version, param.device, param.dtype)
    hit = cache.get(name)
    if hit is not None and hit[0] == key:
        return hit[1]
    src = param.detach()
    if not src.is_contiguous():
        src = src.contiguous()
    packed = src if src.dtype == dtype else ops.cast(src, dtype)
    cache[name] = (key, packed)
    return packed


_HD = 64   # head width of the tcgen05 attention kernels


def _cached(module: nn.Module, name: str, sources, build):
    """Derived tensor cached on the (storage, version, device, dtype) of its source parameters."""
    cache = module.__dict__.setdefault("_mt_pack_cache", {})
    key = tuple(None if q is None else (q.data_ptr(), q._version, q.device, q.dtype) for q in sources)
    hit = cache.get(name)
    if hit is not None and hit[0] == key:
        return hit[1]
    val = build()
    cache[name] = (key, val)
    return val


def _pad_qkv_rows(t: Optional[torch.Tensor], H: int, dh: int) -> Optional[torch.Tensor]:
    """[3*H*dh, ...] -> [3*H*64, ...] with zero rows for the padded lanes of every head (one-time weight preparation).
    head_dim < 64 (Graph fork: 768 / 32 heads = 24) then runs on the head_dim-64 tensor-core kernels, forward and
    backward: zero q / k lanes add nothing to q.k, zero v lanes give zero output lanes, which the zero proj columns ignore."""
    if t is None or dh == _HD:
        return t
    out = t.new_zeros((3, H, _HD) + tuple(t.shape[1:]))
    out[:, :, :dh] = t.reshape((3, H, dh) + tuple(t.shape[1:]))
    return out.reshape((3 * H * _HD,) + tuple(t.shape[1:]))


def _pad_proj_cols(w: torch.Tensor, H: int, dh: int) -> torch.Tensor:
    if dh == _HD:
        return w
    out = w.new_zeros((w.shape[0], H, _HD))
    out[:, :, :dh] = w.reshape(w.shape[0], H, dh)
    return out.reshape(w.shape[0], H * _HD)


def _check_input(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{what}: input is on {x.device}; metatransformer_b200 runs on B200 GPUs only and has no CPU "
                           "fallback (use oracle/ for CPU reference numbers)")
    if x.dim() != 3:
        raise ValueError(f"{what}: expected (B, N, C) input, got shape {tuple(x.shape)}")
    if torch.is_grad_enabled() and x.requires_grad and what != "Block":
        raise NotImplementedError(f"{what}: autograd is implemented at Block granularity only (SURVEY.md §8a E8); "
                                  "call the enclosing Block or use torch.no_grad()")


def _to_bf16_2d(x: torch.Tensor) -> torch.Tensor:
    x2 = x.reshape(-1, x.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()          # e.g. the transposed view Image.PatchEmbed returns (Data2Seq/Image.py:26)
    return x2 if x2.dtype == torch.bfloat16 else ops.cast(x2, torch.bfloat16)


def _from_bf16(y2: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    y = y2 if like.dtype == torch.bfloat16 else ops.cast(y2, like.dtype)
    return y.reshape(like.shape[:-1] + (y2.shape[-1],))


class Mlp(nn.Module):
    """timm 0.4.12 Mlp: fc2(drop(act(fc1(x)))) with one shared `drop` (mlp.py:11-35 in the OpenPoints copy)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        if not isinstance(self.act, nn.GELU) or getattr(self.act, "approximate", "none") != "none":
            raise NotImplementedError("Mlp: only act_layer=nn.GELU (exact erf) has a fused B200 epilogue")

    def _check_train(self):
        if self.training and self.drop.p > 0:
            raise NotImplementedError("Mlp: dropout p > 0 in training mode is built at Block granularity only")

    def _forward_2d(self, h: torch.Tensor, residual: Optional[torch.Tensor]) -> torch.Tensor:
        w1 = _packed(self, "fc1.weight", self.fc1.weight, torch.bfloat16)
        b1 = _packed(self, "fc1.bias", self.fc1.bias, torch.float32)
        w2 = _packed(self, "fc2.weight", self.fc2.weight, torch.bfloat16)
        b2 = _packed(self, "fc2.bias", self.fc2.bias, torch.float32)
        hid = ops.linear(h, w1, b1, gelu=True)
        return ops.linear(hid, w2, b2, residual=residual)

    def forward(self, x):
        _check_input(x, "Mlp")
        self._check_train()
        return _from_bf16(self._forward_2d(_to_bf16_2d(x), None), x)


class Attention(nn.Module):
    """timm 0.4.12 Attention (attention.py:12-38): qkv Linear -> softmax(q k^T * scale) v -> proj Linear."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        if dim % num_heads != 0:
            raise ValueError("Attention: dim must be divisible by num_heads")

    def _check_train(self):
        if self.training and self.attn_drop.p > 0:
            raise NotImplementedError("Attention: attn_drop > 0 in training mode is built at Block granularity only "
                                      "(the hash mask is generated inside the attention kernels)")
        if self.training and self.proj_drop.p > 0:
            raise NotImplementedError("Attention: proj_drop > 0 in training mode is built at Block granularity only")

    @property
    def head_dim(self) -> int:
        return self.qkv.in_features // self.num_heads

    def _pads(self) -> bool:
        """head_dim < 64: the packed weights are zero-padded to 64 lanes per head (tensor-core kernels, with backward)"""
        return self.head_dim < _HD

    def _qkv_packed(self):
        if not self._pads():
            return (_packed(self, "qkv.weight", self.qkv.weight, torch.bfloat16),
                    _packed(self, "qkv.bias", self.qkv.bias, torch.float32))
        H, dh = self.num_heads, self.head_dim
        return _cached(self, "pad.qkv", (self.qkv.weight, self.qkv.bias), lambda: (
            ops.cast(_pad_qkv_rows(self.qkv.weight.detach().float().contiguous(), H, dh), torch.bfloat16),
            None if self.qkv.bias is None else _pad_qkv_rows(self.qkv.bias.detach().float(), H, dh).contiguous()))

    def _proj_packed(self):
        bp = _packed(self, "proj.bias", self.proj.bias, torch.float32)
        if not self._pads():
            return _packed(self, "proj.weight", self.proj.weight, torch.bfloat16), bp
        H, dh = self.num_heads, self.head_dim
        wp = _cached(self, "pad.proj", (self.proj.weight,), lambda: ops.cast(
            _pad_proj_cols(self.proj.weight.detach().float().contiguous(), H, dh).contiguous(), torch.bfloat16))
        return wp, bp

    def _unpad_index(self, device):
        """positions of the real (3, H, dh) rows inside the padded (3, H, 64) layout, and of the (H, dh) columns"""
        H, dh = self.num_heads, self.head_dim
        cache = self.__dict__.setdefault("_mt_pack_cache", {})
        hit = cache.get("pad.idx")
        if hit is None or hit[0].device != device:
            lane = torch.arange(dh, device=device)
            cols = (torch.arange(H, device=device)[:, None] * _HD + lane[None, :]).reshape(-1)
            rows = (torch.arange(3, device=device)[:, None] * (H * _HD) + cols[None, :]).reshape(-1)
            hit = (rows, cols)
            cache["pad.idx"] = hit
        return hit

    def _forward_2d(self, h: torch.Tensor, B: int, N: int, residual: Optional[torch.Tensor]) -> torch.Tensor:
        wq, bq = self._qkv_packed()
        wp, bp = self._proj_packed()
        qkv = ops.linear(h, wq, bq)                                   # (M, 3C), packed (3, H, dh)
        o = ops.attention(qkv.view(B, N, -1), self.num_heads, self.scale)   # (B, N, C)
        return ops.linear(o.view(B * N, -1), wp, bp, residual=residual)

    def forward(self, x):
        _check_input(x, "Attention")
        self._check_train()
        B, N, C = x.shape
        return _from_bf16(self._forward_2d(_to_bf16_2d(x), B, N, None), x)


def _layernorm_eps(norm: nn.Module) -> float:
    if not isinstance(norm, nn.LayerNorm) or not norm.elementwise_affine:
        raise NotImplementedError("Block: only norm_layer=nn.LayerNorm (affine) has a B200 kernel")
    return norm.eps


class Block(nn.Module):
    """timm Block: x = x + drop_path(attn(norm1(x))); x = x + drop_path(mlp(norm2(x)))  (attention.py:41-58).

    Seven kernel launches per forward: LN, GEMM(qkv+bias), attention, GEMM(proj+bias+residual), LN,
    GEMM(fc1+bias+GELU), GEMM(fc2+bias+residual).
    """

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, init_values=None):
        super().__init__()
        if init_values is not None:
            raise NotImplementedError("Block: layer-scale (init_values) is not part of the Meta-Transformer encoder")
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop)
        _layernorm_eps(self.norm1)
        _layernorm_eps(self.norm2)

    # LayerNorm folding (default): norm1 / norm2 are applied inside the qkv / fc1 GEMM epilogues from per-row
    # statistics, the normalised activations never touch HBM.  Set to False for the plain LN-kernel + GEMM sequence.
    fold_layernorm = True

    # Arithmetic of the inference forward:
    #   "bf16": bf16 tensor-core operands, fp32 accumulation, split residual stream (1e-2 gate) — the throughput path;
    #   "fp32": fp32-accurate 3xTF32 tensor-core GEMMs + fp32 LayerNorm / attention (1e-4 gate), ~6-8x slower;
    #   "auto": what the reference's own dtype rules give — fp32 tensors outside torch.autocast run the fp32 path
    #           (X-Ray/train.py:52-60), everything else (bf16 / fp16 tensors, or any tensor under autocast:
    #           Audio/src/models/ast_models.py:145, Video/engine_for_finetuning.py:97) runs the bf16 path.
    # Training (autograd) always uses the bf16 kernels.
    precision = "auto"

    def _use_fp32(self, x: torch.Tensor) -> bool:
        if self.precision == "fp32":
            return True
        return self.precision == "auto" and x.dtype == torch.float32 and not torch.is_autocast_enabled()

    def _forward_fp32(self, x2: torch.Tensor, B: int, N: int) -> torch.Tensor:
        """fp32-accurate Block forward on an fp32 [M, C] residual stream (csrc/mt_fp32.cu)."""
        attn, mlp = self.attn, self.mlp

        def pair(mod, name, w):
            return _cached(mod, "f32pair." + name, (w,), lambda: ops.split_tf32(w.detach().float().contiguous()))

        def vec(mod, name, v):
            return _packed(mod, name, v, torch.float32)

        hb, hs = ops.layernorm_f32_split(x2, vec(self, "norm1.weight", self.norm1.weight),
                                         vec(self, "norm1.bias", self.norm1.bias), self.norm1.eps)
        wqb, wqs = pair(attn, "qkv", attn.qkv.weight)
        qkv = ops.gemm_f32x3(hb, hs, wqb, wqs, bias=vec(attn, "qkv.bias", attn.qkv.bias))
        ob, os_ = ops.attention_f32(qkv.view(B, N, -1), attn.num_heads, attn.scale)
        wpb, wps = pair(attn, "proj", attn.proj.weight)
        C = x2.shape[1]
        x_mid = ops.gemm_f32x3(ob.view(B * N, C), os_.view(B * N, C), wpb, wps, bias=vec(attn, "proj.bias", attn.proj.bias),
                               residual=x2)
        hb, hs = ops.layernorm_f32_split(x_mid, vec(self, "norm2.weight", self.norm2.weight),
                                         vec(self, "norm2.bias", self.norm2.bias), self.norm2.eps)
        w1b, w1s = pair(mlp, "fc1", mlp.fc1.weight)
        hidb, hids = ops.gemm_f32x3(hb, hs, w1b, w1s, bias=vec(mlp, "fc1.bias", mlp.fc1.bias), gelu=True, split_out=True)
        w2b, w2s = pair(mlp, "fc2", mlp.fc2.weight)
        return ops.gemm_f32x3(hidb, hids, w2b, w2s, bias=vec(mlp, "fc2.bias", mlp.fc2.bias), residual=x_mid)

    def _folded(self, name: str, lin: nn.Linear, norm: nn.LayerNorm):
        """(W*gamma bf16, colsum, b + W beta) for Linear(LayerNorm(.)); cached like `_packed`."""
        cache = self.__dict__.setdefault("_mt_pack_cache", {})
        params = (lin.weight, lin.bias, norm.weight, norm.bias)
        key = tuple(None if q is None else (q.data_ptr(), q._version, q.device, q.dtype) for q in params)
        hit = cache.get(name)
        if hit is not None and hit[0] == key:
            return hit[1]
        w = lin.weight.detach().contiguous()
        folded = ops.fold_layernorm(w, _packed(self, name + ".g", norm.weight, torch.float32),
                                    _packed(self, name + ".b", norm.bias, torch.float32),
                                    _packed(self, name + ".bias", lin.bias, torch.float32))
        cache[name] = (key, folded)
        return folded

    def _folded_qkv(self):
        """(W*gamma, colsum, b + W beta) of qkv(norm1(.)), zero-padded to 64 lanes per head when head_dim < 64"""
        attn = self.attn
        folded = self._folded("fold.qkv", attn.qkv, self.norm1)
        if not attn._pads():
            return folded
        H, dh = attn.num_heads, attn.head_dim
        return _cached(self, "fold.qkv.pad", (attn.qkv.weight, attn.qkv.bias, self.norm1.weight, self.norm1.bias), lambda: (
            _pad_qkv_rows(folded[0], H, dh).contiguous(), _pad_qkv_rows(folded[1], H, dh).contiguous(),
            _pad_qkv_rows(folded[2], H, dh).contiguous()))

    def _forward_2d(self, x2: torch.Tensor, B: int, N: int, x2_lo: Optional[torch.Tensor] = None,
                    want_lo: bool = False):
        """(hi, lo | None) of the Block output.  The residual stream is a split pair of bf16 tensors hi + lo (16 mantissa
        bits, see mt_gemm.cu): x2_lo is the low half of the input (None for a bf16 input), the low half of the output is
        produced when want_lo; inside the Block x_mid always carries its low half."""
        if not self.fold_layernorm:
            g1 = _packed(self, "norm1.weight", self.norm1.weight, torch.float32)
            b1 = _packed(self, "norm1.bias", self.norm1.bias, torch.float32)
            g2 = _packed(self, "norm2.weight", self.norm2.weight, torch.float32)
            b2 = _packed(self, "norm2.bias", self.norm2.bias, torch.float32)
            h = ops.layernorm(x2, g1, b1, self.norm1.eps)
            x_mid = self.attn._forward_2d(h, B, N, residual=x2)
            h = ops.layernorm(x_mid, g2, b2, self.norm2.eps, out=h)
            return self.mlp._forward_2d(h, residual=x_mid), None
        attn, mlp = self.attn, self.mlp
        wq, csq, bq = self._folded_qkv()
        stats = ops.row_stats(x2, self.norm1.eps, x_lo=x2_lo)
        qkv = ops.linear_ln(x2, wq, bq, csq, stats)
        o = ops.attention(qkv.view(B, N, -1), attn.num_heads, attn.scale)
        mid_lo = torch.empty_like(x2)
        wp, bp = attn._proj_packed()
        x_mid = ops.linear(o.view(B * N, -1), wp, bp, residual=x2, residual_lo=x2_lo, out_lo=mid_lo)
        w1, cs1, b1 = self._folded("fold.fc1", mlp.fc1, self.norm2)
        stats = ops.row_stats(x_mid, self.norm2.eps, out=stats, x_lo=mid_lo)
        hid = ops.linear_ln(x_mid, w1, b1, cs1, stats, gelu=True)
        out_lo = torch.empty_like(x2) if want_lo else None
        out = ops.linear(hid, _packed(mlp, "fc2.weight", mlp.fc2.weight, torch.bfloat16),
                         _packed(mlp, "fc2.bias", mlp.fc2.bias, torch.float32), residual=x_mid, residual_lo=mid_lo,
                         out_lo=out_lo)
        return out, out_lo

    def _transposed(self, name: str, w: torch.Tensor, *sources: Optional[torch.Tensor]) -> torch.Tensor:
        """bf16 [in, out] copy of a packed bf16 [out, in] weight: the B operand of the dgrad GEMM  dX = dY @ W.
        Cached on the SOURCE parameters' (storage, version, device, dtype) — the same key `_packed` / `_folded` use — so
        the transpose is rebuilt whenever they are: the packed tensor itself comes from torch.empty (version always 0,
        addresses recycled by the caching allocator) and cannot tell two weights apart."""
        cache = self.__dict__.setdefault("_mt_pack_cache", {})
        key = tuple(None if q is None else (q.data_ptr(), q._version, q.device, q.dtype) for q in sources)
        hit = cache.get(name)
        if hit is not None and hit[0] == key:
            return hit[1]
        wt = w.t().contiguous()           # one-time weight preparation (plumbing, not on the per-step path)
        cache[name] = (key, wt)
        return wt

    def _is_stochastic(self) -> bool:
        p_path = self.drop_path.drop_prob if isinstance(self.drop_path, DropPath) else 0.0
        return self.training and (self.attn.proj_drop.p > 0 or self.mlp.drop.p > 0 or p_path > 0
                                  or self.attn.attn_drop.p > 0)

    @staticmethod
    def _replay(g: torch.Tensor, p: float, key, rs, N: int) -> torch.Tensor:
        """Apply a forward dropout / drop-path mask to a gradient (identity when that site had neither)."""
        if p == 0 and rs is None:
            return g
        return ops.dropout(g, p, *key, row_scale=rs, rows_per_scale=N)

    def _drop_config(self, B: int, device: torch.device):
        """Training-time stochastic pieces of this Block, or None: nn.Dropout after proj / inside and after the Mlp
        (attention.py:37, mlp.py:31-34) and per-sample DropPath on both residual branches (attention.py:56-57)."""
        if not self.training:
            return None
        p_proj, p_mlp, p_attn = self.attn.proj_drop.p, self.mlp.drop.p, self.attn.attn_drop.p
        p_path = self.drop_path.drop_prob if isinstance(self.drop_path, DropPath) else 0.0
        if p_proj == 0 and p_mlp == 0 and p_path == 0 and p_attn == 0:
            return None
        # "plain": only the attention dropout is active — the residual GEMMs keep their fused residual adds
        cfg = {"p_proj": p_proj, "p_mlp": p_mlp, "p_attn": p_attn, "seed_attn": 0, "rs1": None, "rs2": None,
               "plain": p_proj == 0 and p_mlp == 0 and p_path == 0}
        if p_path > 0:
            keep = 1.0 - p_path
            rs = torch.empty((2, B), dtype=torch.float32, device=device).bernoulli_(keep).div_(keep)   # timm DropPath
            cfg["rs1"], cfg["rs2"] = rs[0], rs[1]
        return cfg

    def _forward_train(self, x2: torch.Tensor, B: int, N: int):
        """Forward that keeps what the backward needs (qkv, attention output + lse, the pre-GELU activation)."""
        attn, mlp = self.attn, self.mlp
        drop = self._drop_config(B, x2.device)
        wq, csq, bq = self._folded_qkv()
        stats1 = ops.row_stats(x2, self.norm1.eps)
        qkv = ops.linear_ln(x2, wq, bq, csq, stats1)
        if drop is not None and drop["p_attn"] > 0:
            # attention dropout (attention.py:33): hash mask keyed on a 64-bit draw from torch's CUDA generator, replayed
            # by the backward kernel
            sd_, off_ = _philox_reserve(x2.device, B * attn.num_heads * N * N)
            drop["seed_attn"] = (sd_ ^ (off_ * 0x9E3779B97F4A7C15)) & (2 ** 64 - 1)
        o, lse2 = ops.attention_lse(qkv.view(B, N, -1), attn.num_heads, attn.scale,
                                    drop_p=0.0 if drop is None else drop["p_attn"],
                                    seed=0 if drop is None else drop["seed_attn"])
        wp, bp = attn._proj_packed()
        plain = drop is None or drop["plain"]
        if plain:
            x_mid = ops.linear(o.view(B * N, -1), wp, bp, residual=x2)
        else:
            branch = ops.linear(o.view(B * N, -1), wp, bp)
            drop["k1"] = _philox_reserve(x2.device, branch.numel())
            x_mid = ops.dropout(branch, drop["p_proj"], *drop["k1"], residual=x2, row_scale=drop["rs1"],
                                rows_per_scale=N, out=branch)
        w1, cs1, b1 = self._folded("fold.fc1", mlp.fc1, self.norm2)
        stats2 = ops.row_stats(x_mid, self.norm2.eps)
        u = ops.linear_ln(x_mid, w1, b1, cs1, stats2, gelu=False)
        hid = ops.gelu(u)
        w2 = _packed(mlp, "fc2.weight", mlp.fc2.weight, torch.bfloat16)
        b2 = _packed(mlp, "fc2.bias", mlp.fc2.bias, torch.float32)
        if plain:
            out = ops.linear(hid, w2, b2, residual=x_mid)
        else:
            if drop["p_mlp"] > 0:
                drop["kh"] = _philox_reserve(x2.device, hid.numel())
                hid = ops.dropout(hid, drop["p_mlp"], *drop["kh"], out=hid)
            branch = ops.linear(hid, w2, b2)
            drop["k2"] = _philox_reserve(x2.device, branch.numel())
            out = ops.dropout(branch, drop["p_mlp"], *drop["k2"], residual=x_mid, row_scale=drop["rs2"],
                              rows_per_scale=N, out=branch)
        return out, (x2, stats1, qkv, o, lse2, x_mid, stats2, u, drop)

    def _backward_2d(self, g2: torch.Tensor, saved, B: int, N: int) -> torch.Tensor:
        """dX of the Block for frozen weights (dgrad only): 4 GEMMs + attention backward + 2 LayerNorm backwards."""
        x2, stats1, qkv, o, lse2, x_mid, stats2, u, drop = saved
        attn, mlp = self.attn, self.mlp
        M = B * N
        # gradient entering the mlp branch (after its dropout / stochastic depth, replayed from the saved Philox keys)
        plain = drop is None or drop["plain"]
        gb2 = g2 if plain else self._replay(g2, drop["p_mlp"], drop["k2"], drop["rs2"], N)
        w2t = self._transposed("t.fc2", _packed(mlp, "fc2.weight", mlp.fc2.weight, torch.bfloat16), mlp.fc2.weight)
        du = ops.linear_gelu_grad(gb2, w2t, u)                                  # dL/du = (g W2) * gelu'(u)
        if drop is not None and drop["p_mlp"] > 0:
            du = ops.dropout(du, drop["p_mlp"], *drop["kh"], out=du)            # the mask sits between gelu and fc2
        w1f = self._folded("fold.fc1", mlp.fc1, self.norm2)[0]
        dxh2 = ops.linear(du, self._transposed("t.fc1f", w1f, mlp.fc1.weight, self.norm2.weight))                  # dL/dxhat (gamma already folded in)
        g1 = ops.layernorm_bwd(dxh2, x_mid, stats2, None, add=g2)               # + residual branch
        gb1 = g1 if plain else self._replay(g1, drop["p_proj"], drop["k1"], drop["rs1"], N)
        wpt = self._transposed("t.proj", attn._proj_packed()[0], attn.proj.weight)
        do = ops.linear(gb1, wpt)
        dqkv = ops.attention_bwd(qkv.view(B, N, -1), o, do.view(B, N, -1), lse2, attn.num_heads, attn.scale,
                                 drop_p=0.0 if drop is None else drop["p_attn"],
                                 seed=0 if drop is None else drop["seed_attn"])
        wqf = self._folded_qkv()[0]
        dxh1 = ops.linear(dqkv.view(M, -1), self._transposed("t.qkvf", wqf, attn.qkv.weight, self.norm1.weight))
        dx = ops.layernorm_bwd(dxh1, x2, stats1, None, add=g1)
        return dx, (gb2, du, gb1, dqkv.view(M, -1))

    _PARAM_ORDER = ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight",
                    "attn.proj.bias", "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight",
                    "mlp.fc2.bias")

    def _param_list(self):
        named = dict(self.named_parameters())
        return [named.get(k) for k in self._PARAM_ORDER]

    def _weight_grads(self, saved, grads):
        """Fine-tune mode: gradients of all 12 parameters, in `_PARAM_ORDER`.
        Every Linear's dW = dY^T X runs as the TN GEMM on transposed copies (fp32 accumulation over the tokens, bf16
        result); the two LayerNorm-folded Linears are un-folded by mt_ln_fold_bwd."""
        x2, stats1, qkv, o, lse2, x_mid, stats2, u, drop = saved
        g2, du, g1, dqkv = grads          # g2 / g1: gradients entering the fc2 / proj branches (dropout already replayed)
        attn, mlp = self.attn, self.mlp
        d = x2.shape[1]
        cache = self.__dict__.setdefault("_mt_pack_cache", {})
        if "ones" not in cache or cache["ones"][0].device != x2.device or cache["ones"][0].numel() != d:
            cache["ones"] = (torch.ones(d, device=x2.device), torch.zeros(d, device=x2.device))
        ones, zeros = cache["ones"]
        f32 = lambda p_: _packed(self, "f32." + str(id(p_)), p_, torch.float32)  # noqa: E731
        # --- mlp.fc2: input gelu(u) ---
        hid = ops.gelu(u)
        if drop is not None and drop["p_mlp"] > 0:
            hid = ops.dropout(hid, drop["p_mlp"], *drop["kh"], out=hid)
        dw2 = ops.linear_wgrad(g2, hid)
        db2 = ops.colsum(g2)
        del hid
        # --- mlp.fc1 (norm2 folded): input xhat2 ---
        xh = ops.layernorm(x_mid, ones, zeros, self.norm2.eps)
        dwf = ops.linear_wgrad(du, xh)
        dyb1 = ops.colsum(du)
        dw1, dg2, dbt2 = ops.ln_fold_bwd(dwf, mlp.fc1.weight.detach().contiguous(), f32(self.norm2.weight),
                                         f32(self.norm2.bias), dyb1)
        # --- attn.proj: input o ---
        dwp = ops.linear_wgrad(g1, o.view(o.shape[0] * o.shape[1], -1))
        dbp = ops.colsum(g1)
        if attn._pads():                       # drop the zero-padded head lanes again (weight-gradient re-layout)
            rows_idx, cols_idx = attn._unpad_index(x2.device)
            dwp = dwp.index_select(1, cols_idx).contiguous()
        # --- attn.qkv (norm1 folded): input xhat1 ---
        xh = ops.layernorm(x2, ones, zeros, self.norm1.eps, out=xh)
        dwf = ops.linear_wgrad(dqkv, xh)
        dybq = ops.colsum(dqkv)
        if attn._pads():
            dwf = dwf.index_select(0, rows_idx).contiguous()
            dybq = dybq.index_select(0, rows_idx).contiguous()
        dwq, dg1, dbt1 = ops.ln_fold_bwd(dwf, attn.qkv.weight.detach().contiguous(), f32(self.norm1.weight),
                                         f32(self.norm1.bias), dybq)
        return [dg1, dbt1, dwq, dybq if attn.qkv.bias is not None else None, dwp, dbp, dg2, dbt2, dw1, dyb1, dw2, db2]

    def forward(self, x, pos=None):
        """`pos` (extension): a (1, N, C) / (N, C) table added to x before the Block — the PointCloud fork's
        `x = block(x + pos_embed)` loop (PointCloud/openpoints/models/backbone/metatransformer.py:161-163) as one fused
        pass instead of an eager add per layer; d pos = the batch sum of d x."""
        if pos is not None:
            x = _AddPosFn.apply(x, pos)
        _check_input(x, "Block")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return _BlockFn.apply(x, self, *[p for p in self._param_list() if p is not None])
        B, N, C = x.shape
        if self._is_stochastic():                               # train mode without autograd: masks still apply
            return _from_bf16(self._forward_train(_to_bf16_2d(x), B, N)[0], x)
        if self._use_fp32(x):
            x2 = x.reshape(-1, C).float()
            y = self._forward_fp32(x2 if x2.is_contiguous() else x2.contiguous(), B, N).view(B, N, C)
            return y if x.dtype == torch.float32 else ops.cast(y, x.dtype)
        if x.dtype == torch.float32 and self.fold_layernorm and x.numel() % 4 == 0:
            # fp32 caller (autocast-style: fp32 residual stream, bf16 tensor-core operands): the stream enters and leaves
            # the Block as a split hi + lo pair, so a Sequential of Blocks accumulates its 2L residual adds in fp32
            x2 = x.reshape(-1, C)
            hi, lo = ops.split_f32(x2 if x2.is_contiguous() else x2.contiguous())
            out_hi, out_lo = self._forward_2d(hi, B, N, lo, want_lo=True)
            return ops.merge_f32(out_hi, out_lo).view(B, N, C)
        return _from_bf16(self._forward_2d(_to_bf16_2d(x), B, N)[0], x)


class _AddPosFn(torch.autograd.Function):
    """x (B, N, C) + pos (N, C) broadcast over the batch, one pass of mt_assemble_tokens (bf16 result)."""

    @staticmethod
    def forward(ctx, x, pos):
        if not x.is_cuda:
            raise RuntimeError("Block: CUDA (B200) tensors only; there is no CPU fallback")
        B, N, C = x.shape
        if pos.numel() != N * C:
            raise ValueError(f"Block: pos must hold N*C = {N * C} values, got shape {tuple(pos.shape)}")
        ctx.meta = (x.dtype, pos.dtype, pos.shape)
        xb = _to_bf16_2d(x).view(B, N, C)
        pb = pos.detach().reshape(N, C).contiguous()
        pb = pb if pb.dtype == torch.bfloat16 else ops.cast(pb, torch.bfloat16)
        return ops.assemble_tokens(xb, None, pb)

    @staticmethod
    def backward(ctx, g):
        xdt, pdt, pshape = ctx.meta
        B, N, C = g.shape
        gb = g.contiguous()
        gb = gb if gb.dtype == torch.bfloat16 else ops.cast(gb, torch.bfloat16)
        dpos = None
        if ctx.needs_input_grad[1]:
            dpos = ops.colsum(gb.view(B, N * C)).view(pshape)
            dpos = dpos if pdt == torch.float32 else ops.cast(dpos, pdt)
        dx = g if g.dtype == xdt else ops.cast(gb, xdt)
        return dx, dpos


class _BlockFn(torch.autograd.Function):
    """Autograd node of one frozen Block: forward and input-gradient both run on the sm_100a kernels."""

    @staticmethod
    def forward(ctx, x, block, *params):
        B, N, C = x.shape
        out2, saved = block._forward_train(_to_bf16_2d(x.detach()), B, N)
        ctx.block, ctx.saved, ctx.shape, ctx.in_dtype = block, saved, (B, N, C), x.dtype
        ctx.params = params
        return _from_bf16(out2, x)

    @staticmethod
    def backward(ctx, g):
        if ctx.saved is None:
            raise RuntimeError("Block: backward through the same graph a second time is not supported (the saved "
                               "activations are released after the first pass; retain_graph=True has no effect here)")
        B, N, C = ctx.shape
        g2 = _to_bf16_2d(g)
        dx2, inter = ctx.block._backward_2d(g2, ctx.saved, B, N)
        pgrads = [None] * len(ctx.params)
        if any(ctx.needs_input_grad[2:]):
            full = ctx.block._weight_grads(ctx.saved, inter)
            present = [gr for gr, prm in zip(full, ctx.block._param_list()) if prm is not None]
            for i, (gr, prm) in enumerate(zip(present, ctx.params)):
                if ctx.needs_input_grad[2 + i] and gr is not None:
                    gr = gr.view(prm.shape)
                    pgrads[i] = gr if gr.dtype == prm.dtype else ops.cast(gr.contiguous(), prm.dtype)
        ctx.saved = None
        dx = None
        if ctx.needs_input_grad[0]:
            dx = (dx2 if ctx.in_dtype == torch.bfloat16 else ops.cast(dx2, ctx.in_dtype)).view(B, N, C)
        return (dx, None, *pgrads)
