"""Functional wrappers: torch CUDA tensors in, one C-ABI kernel launch each, torch tensors out.

PyTorch is used for device memory and streams only; every FLOP of these ops runs in libmt_b200.so.
All functions raise (`MtError` / `ValueError`) instead of falling back.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import MT_DT_BF16, MT_DT_F16, MT_DT_F32, check

_DT = {torch.bfloat16: MT_DT_BF16, torch.float32: MT_DT_F32, torch.float16: MT_DT_F16}
_device_checked = set()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError("metatransformer_b200 ops need CUDA tensors on a B200 (no CPU fallback exists)")
        if not t.is_contiguous():
            raise ValueError("metatransformer_b200 ops need contiguous tensors")
    dev = next(t.device.index for t in tensors if t is not None)
    if dev not in _device_checked:
        check(_lib.load().mt_device_ok(dev), "mt_device_ok")
        _device_checked.add(dev)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def cast(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """dtype conversion on device (mt_cast)."""
    _require_cuda(x)
    if x.dtype == dtype:
        return x
    out = torch.empty_like(x, dtype=dtype)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_cast(x.data_ptr(), _DT[x.dtype], out.data_ptr(), _DT[dtype], x.numel(), _stream()), "mt_cast")
    return out


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, gelu: bool = False, cta_group: int = 0,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = act(x @ weight.T + bias) (+ residual).  x [M,K] bf16, weight [N,K] bf16, bias [N] fp32, residual [M,N] bf16."""
    _require_cuda(x, weight, bias, residual, out)
    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise ValueError("linear: x and weight must be bf16")
    if bias is not None and bias.dtype != torch.float32:
        raise ValueError("linear: bias must be fp32")
    if residual is not None and residual.dtype != torch.bfloat16:
        raise ValueError("linear: residual must be bf16")
    M, K = x.shape
    N, K2 = weight.shape
    if K != K2:
        raise ValueError(f"linear: K mismatch {K} vs {K2}")
    if residual is not None and tuple(residual.shape) != (M, N):
        raise ValueError("linear: residual shape mismatch")
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_gemm_bf16(x.data_ptr(), weight.data_ptr(), _ptr(bias), _ptr(residual), out.data_ptr(),
                                       M, N, K, 1 if gelu else 0, cta_group, _stream()), "mt_gemm_bf16")
    return out


def row_stats(x: torch.Tensor, eps: float = 1e-5, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-row LayerNorm statistics of x [M,d] (bf16/fp32): fp32 [M,2] = (mean, rstd)."""
    _require_cuda(x, out)
    M, d = x.shape
    if out is None:
        out = torch.empty((M, 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_row_stats(x.data_ptr(), _DT[x.dtype], out.data_ptr(), M, d, float(eps), _stream()),
              "mt_row_stats")
    return out


def fold_layernorm(weight: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bias: Optional[torch.Tensor]):
    """One-time weight preparation for Linear(LayerNorm(x)): returns (W*gamma as bf16 [N,K], colsum [N], b + W beta [N])."""
    _require_cuda(weight, gamma, beta, bias)
    if weight.dtype not in (torch.float32, torch.bfloat16):
        weight = weight.float()
    N, K = weight.shape
    wf = torch.empty((N, K), dtype=torch.bfloat16, device=weight.device)
    cs = torch.empty(N, dtype=torch.float32, device=weight.device)
    bf = torch.empty(N, dtype=torch.float32, device=weight.device)
    with torch.cuda.device(weight.device):
        check(_lib.load().mt_fold_layernorm(weight.data_ptr(), _DT[weight.dtype], gamma.data_ptr(), beta.data_ptr(),
                                            _ptr(bias), wf.data_ptr(), cs.data_ptr(), bf.data_ptr(), N, K, _stream()),
              "mt_fold_layernorm")
    return wf, cs, bf


def linear_ln(x: torch.Tensor, w_folded: torch.Tensor, bias_folded: torch.Tensor, colsum: torch.Tensor,
              stats: torch.Tensor, gelu: bool = False, cta_group: int = 0,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = act(LayerNorm(x) @ W.T + b) with the LayerNorm folded into the GEMM (see fold_layernorm / row_stats)."""
    _require_cuda(x, w_folded, bias_folded, colsum, stats, out)
    M, K = x.shape
    N = w_folded.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_gemm_bf16_ln(x.data_ptr(), w_folded.data_ptr(), bias_folded.data_ptr(), colsum.data_ptr(),
                                          stats.data_ptr(), out.data_ptr(), M, N, K, 1 if gelu else 0, cta_group,
                                          _stream()), "mt_gemm_bf16_ln")
    return out


def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm over the last dim of x [M,d] (bf16 or fp32) -> bf16; weight/bias fp32."""
    _require_cuda(x, weight, bias, out)
    if weight.dtype != torch.float32 or bias.dtype != torch.float32:
        raise ValueError("layernorm: weight/bias must be fp32")
    M, d = x.shape
    if out is None:
        out = torch.empty((M, d), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_layernorm(x.data_ptr(), _DT[x.dtype], weight.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                       M, d, float(eps), _stream()), "mt_layernorm")
    return out


def attention(qkv: torch.Tensor, num_heads: int, scale: Optional[float] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qkv [B,N,3*H*dh] (packed as (3,H,dh)) bf16 -> softmax(q k^T * scale) v as [B,N,H*dh] bf16."""
    _require_cuda(qkv, out)
    if qkv.dtype != torch.bfloat16:
        raise ValueError("attention: qkv must be bf16")
    B, N, C3 = qkv.shape
    dh = C3 // (3 * num_heads)
    if 3 * num_heads * dh != C3:
        raise ValueError("attention: last dim must be 3*H*dh")
    if scale is None:
        scale = dh ** -0.5
    if out is None:
        out = torch.empty((B, N, num_heads * dh), dtype=torch.bfloat16, device=qkv.device)
    with torch.cuda.device(qkv.device):
        check(_lib.load().mt_attention_fwd(qkv.data_ptr(), out.data_ptr(), B, N, num_heads, dh, float(scale), _stream()),
              "mt_attention_fwd")
    return out


def im2row(x: torch.Tensor, kernel: Sequence[int], stride: Sequence[int]) -> Tuple[torch.Tensor, Tuple[int, int, int]]:
    """x [B,C,T,H,W] (fp32/bf16) -> rows [B*Tp*Hp*Wp, C*kt*kh*kw] bf16 and (Tp,Hp,Wp)."""
    _require_cuda(x)
    B, C, T, H, W = x.shape
    kt, kh, kw = kernel
    st, sh, sw = stride
    Tp, Hp, Wp = (T - kt) // st + 1, (H - kh) // sh + 1, (W - kw) // sw + 1
    rows = torch.empty((B * Tp * Hp * Wp, C * kt * kh * kw), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_im2row(x.data_ptr(), _DT[x.dtype], rows.data_ptr(), B, C, T, H, W, kt, kh, kw, st, sh, sw,
                                    _stream()), "mt_im2row")
    return rows, (Tp, Hp, Wp)
