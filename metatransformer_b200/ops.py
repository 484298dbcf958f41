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
           out: Optional[torch.Tensor] = None, residual_period: int = 0,
           residual_lo: Optional[torch.Tensor] = None, out_lo: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = act(x @ weight.T + bias) (+ residual).  x [M,K] bf16, weight [N,K] bf16, bias [N] fp32, residual [M,N] bf16 —
    or, with residual_period = n > 0, a [n,N] table of which output row m adds row m % n (positional embedding).
    Split residual stream: `residual_lo` is the low half of a (hi, lo) bf16 pair and `out_lo` (bf16 [M,N], written)
    receives the low half of the sum; the returned tensor is the high half."""
    _require_cuda(x, weight, bias, residual, out, residual_lo, out_lo)
    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise ValueError("linear: x and weight must be bf16")
    if bias is not None and bias.dtype != torch.float32:
        raise ValueError("linear: bias must be fp32")
    if any(t is not None and t.dtype != torch.bfloat16 for t in (residual, residual_lo, out_lo)):
        raise ValueError("linear: residual / residual_lo / out_lo must be bf16")
    M, K = x.shape
    N, K2 = weight.shape
    if K != K2:
        raise ValueError(f"linear: K mismatch {K} vs {K2}")
    if residual is not None and tuple(residual.shape) != ((residual_period or M), N):
        raise ValueError("linear: residual shape mismatch")
    if (residual_lo is not None or out_lo is not None) and (residual is None or gelu or residual_period):
        raise ValueError("linear: the split (hi, lo) stream needs a plain residual GEMM")
    if any(t is not None and tuple(t.shape) != (M, N) for t in (residual_lo, out_lo)):
        raise ValueError("linear: residual_lo / out_lo shape mismatch")
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        if residual_period > 0 or residual_lo is not None or out_lo is not None:
            o = _lib.GemmOpts(bias=_ptr(bias), residual=_ptr(residual), act=1 if gelu else 0, cta_group=cta_group,
                              residual_period=residual_period, residual_lo=_ptr(residual_lo), y_lo=_ptr(out_lo))
            check(_lib.load().mt_gemm_bf16_ex(x.data_ptr(), weight.data_ptr(), out.data_ptr(), M, N, K, o, _stream()),
                  "mt_gemm_bf16_ex")
        else:
            check(_lib.load().mt_gemm_bf16(x.data_ptr(), weight.data_ptr(), _ptr(bias), _ptr(residual), out.data_ptr(),
                                           M, N, K, 1 if gelu else 0, cta_group, _stream()), "mt_gemm_bf16")
    return out


def split_f32(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 -> (hi, lo) bf16 pair with hi + lo = x to 16 mantissa bits (the residual-stream representation)."""
    _require_cuda(x)
    if x.dtype != torch.float32 or x.numel() % 4:
        raise ValueError("split_f32: fp32 tensor with numel % 4 == 0")
    hi = torch.empty_like(x, dtype=torch.bfloat16)
    lo = torch.empty_like(x, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_split_f32(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream()), "mt_split_f32")
    return hi, lo


def merge_f32(hi: torch.Tensor, lo: torch.Tensor) -> torch.Tensor:
    """(hi, lo) bf16 pair -> fp32 hi + lo."""
    _require_cuda(hi, lo)
    if hi.dtype != torch.bfloat16 or lo.dtype != torch.bfloat16 or hi.shape != lo.shape or hi.numel() % 4:
        raise ValueError("merge_f32: two bf16 tensors of the same shape, numel % 4 == 0")
    out = torch.empty_like(hi, dtype=torch.float32)
    with torch.cuda.device(hi.device):
        check(_lib.load().mt_merge_f32(hi.data_ptr(), lo.data_ptr(), out.data_ptr(), hi.numel(), _stream()), "mt_merge_f32")
    return out


def row_stats(x: torch.Tensor, eps: float = 1e-5, out: Optional[torch.Tensor] = None,
              x_lo: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-row LayerNorm statistics of x [M,d] (bf16/fp32; or of the split pair x + x_lo): fp32 [M,2] = (mean, rstd)."""
    _require_cuda(x, out, x_lo)
    M, d = x.shape
    if out is None:
        out = torch.empty((M, 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        if x_lo is not None:
            check(_lib.load().mt_row_stats_split(x.data_ptr(), x_lo.data_ptr(), out.data_ptr(), M, d, float(eps), _stream()),
                  "mt_row_stats_split")
        else:
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


def attention_lse(qkv: torch.Tensor, num_heads: int, scale: Optional[float] = None, drop_p: float = 0.0, seed: int = 0):
    """attention() that also returns lse2 [B,H,N] (fp32 log2-sum-exp of the scaled scores) for the backward.  drop_p > 0:
    attention dropout with the hash mask of (seed, batch, head, query, key) (attention_bwd replays it)."""
    _require_cuda(qkv)
    B, N, C3 = qkv.shape
    dh = C3 // (3 * num_heads)
    if scale is None:
        scale = dh ** -0.5
    out = torch.empty((B, N, num_heads * dh), dtype=torch.bfloat16, device=qkv.device)
    lse2 = torch.empty((B, num_heads, N), dtype=torch.float32, device=qkv.device)
    with torch.cuda.device(qkv.device):
        check(_lib.load().mt_attention_fwd_ex(qkv.data_ptr(), out.data_ptr(), lse2.data_ptr(), B, N, num_heads, dh,
                                              float(scale), float(drop_p), int(seed) & (2 ** 64 - 1), _stream()),
              "mt_attention_fwd_ex")
    return out, lse2


def attention_dropout_mask(B: int, N: int, num_heads: int, drop_p: float, seed: int, device) -> torch.Tensor:
    """The keep mask [B,H,N,N] (uint8) the attention kernels apply for (drop_p, seed): test / debugging aid."""
    mask = torch.empty((B, num_heads, N, N), dtype=torch.uint8, device=device)
    with torch.cuda.device(mask.device):
        check(_lib.load().mt_attention_dropout_mask(mask.data_ptr(), B, N, num_heads, float(drop_p),
                                                    int(seed) & (2 ** 64 - 1), _stream()), "mt_attention_dropout_mask")
    return mask


def attention_bwd(qkv: torch.Tensor, out: torch.Tensor, dout: torch.Tensor, lse2: torch.Tensor, num_heads: int,
                  scale: Optional[float] = None, drop_p: float = 0.0, seed: int = 0) -> torch.Tensor:
    """d(qkv) [B,N,3*H*dh] bf16 of attention(); `out`/`lse2` are the forward's results, `dout` the gradient of `out`."""
    import ctypes
    _require_cuda(qkv, out, dout, lse2)
    B, N, C3 = qkv.shape
    dh = C3 // (3 * num_heads)
    if scale is None:
        scale = dh ** -0.5
    dqkv = torch.empty_like(qkv)
    need = ctypes.c_size_t()
    lib = _lib.load()
    check(lib.mt_attention_bwd_workspace_bytes(B, N, num_heads, ctypes.byref(need)), "mt_attention_bwd_workspace_bytes")
    ws = torch.empty(need.value, dtype=torch.uint8, device=qkv.device)
    with torch.cuda.device(qkv.device):
        check(lib.mt_attention_bwd_ex(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse2.data_ptr(), dqkv.data_ptr(),
                                      ws.data_ptr(), ws.numel(), B, N, num_heads, dh, float(scale), float(drop_p),
                                      int(seed) & (2 ** 64 - 1), _stream()), "mt_attention_bwd_ex")
    return dqkv


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, stats: torch.Tensor, gamma: Optional[torch.Tensor] = None,
                  add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx of LayerNorm: add + rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*gamma (gamma None: already applied)."""
    _require_cuda(dy, x, stats, gamma, add)
    M, d = x.shape
    dx = torch.empty((M, d), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_layernorm_bwd(dy.data_ptr(), x.data_ptr(), stats.data_ptr(), _ptr(gamma), _ptr(add),
                                           dx.data_ptr(), M, d, _stream()), "mt_layernorm_bwd")
    return dx


def gelu(u: torch.Tensor) -> torch.Tensor:
    """h = gelu(u) (exact erf), bf16."""
    _require_cuda(u)
    h = torch.empty_like(u)
    with torch.cuda.device(u.device):
        check(_lib.load().mt_gelu(u.data_ptr(), h.data_ptr(), u.numel(), _stream()), "mt_gelu")
    return h


def linear_gelu_grad(dy: torch.Tensor, weight_t: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """(dy @ weight_t.T) * gelu'(u): backward through fc2 and the GELU in one GEMM (weight_t = fc2.weight.T, [hidden, d])."""
    _require_cuda(dy, weight_t, u)
    M, K = dy.shape
    N = weight_t.shape[0]
    out = torch.empty((M, N), dtype=torch.bfloat16, device=dy.device)
    with torch.cuda.device(dy.device):
        check(_lib.load().mt_gemm_bf16(dy.data_ptr(), weight_t.data_ptr(), None, u.data_ptr(), out.data_ptr(), M, N, K, 2,
                                       0, _stream()), "mt_gemm_bf16(act=2)")
    return out


def transpose(x: torch.Tensor, pad_to: int = 8) -> torch.Tensor:
    """x [R,C] bf16 -> [C, ceil(R/pad_to)*pad_to] (zero padded): operand layout for the weight-gradient GEMM."""
    _require_cuda(x)
    R, C = x.shape
    ldd = (R + pad_to - 1) // pad_to * pad_to
    out = torch.empty((C, ldd), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_transpose_bf16(x.data_ptr(), out.data_ptr(), R, C, ldd, _stream()), "mt_transpose_bf16")
    return out


def colsum(x: torch.Tensor) -> torch.Tensor:
    """fp32 column sums of x [M,N] bf16 (bias gradient)."""
    _require_cuda(x)
    M, N = x.shape
    out = torch.empty(N, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_colsum(x.data_ptr(), out.data_ptr(), M, N, _stream()), "mt_colsum")
    return out


def ln_fold_bwd(dw_folded: torch.Tensor, weight: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, dyb: torch.Tensor):
    """(dW fp32 [N,K], dgamma fp32 [K], dbeta fp32 [K]) of a LayerNorm-folded Linear from dWf = dY^T xhat (fp32) and dyb = colsum(dY)."""
    _require_cuda(dw_folded, weight, gamma, beta, dyb)
    if dw_folded.dtype != torch.float32:
        raise ValueError("ln_fold_bwd: dw_folded must be fp32")
    if weight.dtype not in (torch.float32, torch.bfloat16):
        weight = weight.float()
    N, K = weight.shape
    dw = torch.empty((N, K), dtype=torch.float32, device=weight.device)
    dg = torch.empty(K, dtype=torch.float32, device=weight.device)
    db = torch.empty(K, dtype=torch.float32, device=weight.device)
    with torch.cuda.device(weight.device):
        check(_lib.load().mt_ln_fold_bwd(dw_folded.data_ptr(), weight.data_ptr(), _DT[weight.dtype], gamma.data_ptr(),
                                         beta.data_ptr(), dyb.data_ptr(), dw.data_ptr(), dg.data_ptr(), db.data_ptr(), N, K,
                                         _stream()), "mt_ln_fold_bwd")
    return dw, dg, db


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, splits: int = 0, via_transpose: bool = False) -> torch.Tensor:
    """dW [N,K] fp32 = dy[M,N]^T @ x[M,K] (fp32 accumulation over the M tokens, split-K tcgen05 GEMM).  Default: both
    operands are read MN-major straight from the row-major activations (mt_gemm_bf16_wgrad); via_transpose=True is the
    older form on transposed bf16 copies (kept as a cross-check)."""
    _require_cuda(dy, x)
    if dy.dtype != torch.bfloat16 or x.dtype != torch.bfloat16 or dy.shape[0] != x.shape[0]:
        raise ValueError("linear_wgrad: dy [M,N] and x [M,K] must be bf16 with the same M")
    M, N = dy.shape
    K = x.shape[1]
    dw = torch.empty((N, K), dtype=torch.float32, device=dy.device)
    with torch.cuda.device(dy.device):
        if via_transpose:
            dyt, xt = transpose(dy), transpose(x)          # [N, Mp], [K, Mp]
            check(_lib.load().mt_gemm_bf16_splitk(dyt.data_ptr(), xt.data_ptr(), dw.data_ptr(), N, K, dyt.shape[1], splits,
                                                  _stream()), "mt_gemm_bf16_splitk")
        else:
            check(_lib.load().mt_gemm_bf16_wgrad(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), M, N, K, splits, _stream()),
                  "mt_gemm_bf16_wgrad")
    return dw


def im2row1d_circular(x: torch.Tensor) -> torch.Tensor:
    """x [B,T,C] (fp32/bf16) -> rows [B*T, Kp] bf16 for the circular k=3 Conv1d token embedding (Kp = 3C rounded up to 8)."""
    _require_cuda(x)
    B, T, C = x.shape
    Kp = (3 * C + 7) // 8 * 8
    rows = torch.empty((B * T, Kp), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_im2row1d_circular(x.data_ptr(), _DT[x.dtype], rows.data_ptr(), B, T, C, Kp, _stream()),
              "mt_im2row1d_circular")
    return rows


def prepend_row(tok: torch.Tensor, first: torch.Tensor) -> torch.Tensor:
    """tok [B,n,d] bf16, first [d] bf16 -> [B,n+1,d] with `first` as token 0 of every sequence."""
    _require_cuda(tok, first)
    B, n, d = tok.shape
    out = torch.empty((B, n + 1, d), dtype=torch.bfloat16, device=tok.device)
    with torch.cuda.device(tok.device):
        check(_lib.load().mt_prepend_row(tok.data_ptr(), first.data_ptr(), out.data_ptr(), B, n, d, _stream()),
              "mt_prepend_row")
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


POOL_FIRST, POOL_MEAN, POOL_FIRST_TWO = 0, 1, 2


def assemble_tokens(tok: torch.Tensor, prefix: Optional[torch.Tensor] = None,
                    pos: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b] = [prefix ; tok[b]] + pos: tok [B,n,d], prefix [k,d], pos [n+k,d], all bf16 -> [B,n+k,d]."""
    _require_cuda(tok, prefix, pos)
    B, n, d = tok.shape
    k = 0 if prefix is None else prefix.shape[0]
    if any(t is not None and t.dtype != torch.bfloat16 for t in (tok, prefix, pos)):
        raise ValueError("assemble_tokens: bf16 tensors only")
    if pos is not None and tuple(pos.shape) != (n + k, d):
        raise ValueError(f"assemble_tokens: pos must be [{n + k}, {d}], got {tuple(pos.shape)}")
    out = torch.empty((B, n + k, d), dtype=torch.bfloat16, device=tok.device)
    with torch.cuda.device(tok.device):
        check(_lib.load().mt_assemble_tokens(tok.data_ptr(), _ptr(prefix), _ptr(pos), out.data_ptr(), B, n, k, d,
                                             _stream()), "mt_assemble_tokens")
    return out


def pool_layernorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, mode: int, eps: float = 1e-5) -> torch.Tensor:
    """Final LayerNorm fused with the pooling: x [B,N,d] bf16 -> [B,d] bf16 (POOL_FIRST / POOL_MEAN / POOL_FIRST_TWO)."""
    _require_cuda(x, weight, bias)
    if x.dtype != torch.bfloat16 or weight.dtype != torch.float32 or bias.dtype != torch.float32:
        raise ValueError("pool_layernorm: x bf16, weight/bias fp32")
    B, N, d = x.shape
    out = torch.empty((B, d), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_pool_layernorm(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(), B, N, d,
                                            int(mode), float(eps), _stream()), "mt_pool_layernorm")
    return out


def pool_layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, weight: torch.Tensor, mode: int, eps: float = 1e-5,
                       param_grads: bool = True):
    """Autograd of pool_layernorm: (dx [B,N,d] bf16, dweight [d] fp32 | None, dbias [d] fp32 | None)."""
    _require_cuda(dy, x, weight)
    B, N, d = x.shape
    if dy.dtype != torch.bfloat16 or tuple(dy.shape) != (B, d):
        raise ValueError("pool_layernorm_bwd: dy must be bf16 [B,d]")
    dx = torch.empty_like(x)
    dg = torch.empty(d, dtype=torch.float32, device=x.device) if param_grads else None
    db = torch.empty(d, dtype=torch.float32, device=x.device) if param_grads else None
    with torch.cuda.device(x.device):
        check(_lib.load().mt_pool_layernorm_bwd(dy.data_ptr(), x.data_ptr(), weight.data_ptr(), dx.data_ptr(), _ptr(dg),
                                                _ptr(db), B, N, d, int(mode), float(eps), _stream()),
              "mt_pool_layernorm_bwd")
    return dx, dg, db


def dropout(x: torch.Tensor, p: float, seed: int, offset: int, residual: Optional[torch.Tensor] = None,
            row_scale: Optional[torch.Tensor] = None, rows_per_scale: int = 1,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = residual + row_scale[row // rows_per_scale] * keep / (1 - p) * x on x [rows, d] bf16; the mask is a pure
    function of (seed, offset, element index), so calling it again on a gradient replays the forward mask."""
    _require_cuda(x, residual, row_scale, out)
    if x.dtype != torch.bfloat16 or (residual is not None and residual.dtype != torch.bfloat16):
        raise ValueError("dropout: bf16 tensors only")
    if row_scale is not None and row_scale.dtype != torch.float32:
        raise ValueError("dropout: row_scale must be fp32")
    rows, d = x.shape
    if out is None:
        out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_dropout(x.data_ptr(), _ptr(residual), _ptr(row_scale), int(rows_per_scale), out.data_ptr(),
                                     rows, d, float(p), int(seed) & (2 ** 64 - 1), int(offset), _stream()), "mt_dropout")
    return out


def pad_cols(x: torch.Tensor, width: int) -> torch.Tensor:
    """Zero-pad the last dim of x [rows, c] (fp32/bf16) to `width` columns (Data2Seq/Text.py:15-19)."""
    _require_cuda(x)
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("pad_cols: fp32 or bf16 only")
    rows, c = x.shape
    out = torch.empty((rows, width), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_pad_cols(x.data_ptr(), out.data_ptr(), _DT[x.dtype], rows, c, width, _stream()), "mt_pad_cols")
    return out


def embed_sum(table: torch.Tensor, idx: torch.Tensor, offsets: Optional[torch.Tensor] = None,
              pos: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """Embedding-bag sum: out[r] = pos[r % len(pos)] + sum_f table[idx[r, f] + offsets[f]].  table fp32 [V,d], idx int32
    [rows,F], offsets int32 [F] | None, pos fp32 [P,d] | None -> [rows, d] in out_dtype (bf16 / fp32)."""
    _require_cuda(table, idx, offsets, pos)
    if table.dtype != torch.float32 or idx.dtype != torch.int32 or (pos is not None and pos.dtype != torch.float32) \
            or (offsets is not None and offsets.dtype != torch.int32):
        raise ValueError("embed_sum: table / pos fp32, idx / offsets int32")
    rows, F = idx.shape
    d = table.shape[1]
    out = torch.empty((rows, d), dtype=out_dtype, device=table.device)
    with torch.cuda.device(table.device):
        check(_lib.load().mt_embed_sum(table.data_ptr(), idx.data_ptr(), _ptr(offsets), F, _ptr(pos),
                                       0 if pos is None else pos.shape[0], out.data_ptr(), _DT[out_dtype], rows, d, _stream()),
              "mt_embed_sum")
    return out


# ------------------------------------------------------------------------------------------------------------------
# fp32-accuracy path (3xTF32 tensor-core GEMMs on (big, small) fp32 pairs; see csrc/mt_fp32.cu)
# ------------------------------------------------------------------------------------------------------------------
def _f32(*tensors):
    _require_cuda(*tensors)
    if any(t is not None and t.dtype != torch.float32 for t in tensors):
        raise ValueError("fp32 path: fp32 tensors only")


def split_tf32(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 -> (big, small): big = x with the low 13 mantissa bits cleared (exact in TF32), small = x - big."""
    _f32(x)
    big, small = torch.empty_like(x), torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_split_tf32(x.data_ptr(), big.data_ptr(), small.data_ptr(), x.numel(), _stream()), "mt_split_tf32")
    return big, small


def layernorm_f32_split(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5):
    """LayerNorm of fp32 x [M,d] in fp32 -> (big, small) pair for gemm_f32x3."""
    _f32(x, weight, bias)
    M, d = x.shape
    big, small = torch.empty_like(x), torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_layernorm_f32_split(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), big.data_ptr(),
                                                 small.data_ptr(), M, d, float(eps), _stream()), "mt_layernorm_f32_split")
    return big, small


def gemm_f32x3(x_big: torch.Tensor, x_small: torch.Tensor, w_big: torch.Tensor, w_small: torch.Tensor,
               bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, gelu: bool = False,
               split_out: bool = False):
    """fp32-accurate y = act(x @ w.T + bias) (+ residual) from (big, small) operand pairs; returns fp32 y, or its
    (big, small) pair when split_out."""
    _f32(x_big, x_small, w_big, w_small, bias, residual)
    M, K = x_big.shape
    N = w_big.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x_big.device)
    ys = torch.empty_like(y) if split_out else None
    with torch.cuda.device(x_big.device):
        check(_lib.load().mt_gemm_f32x3(x_big.data_ptr(), x_small.data_ptr(), w_big.data_ptr(), w_small.data_ptr(), _ptr(bias),
                                        _ptr(residual), y.data_ptr(), _ptr(ys), M, N, K, 1 if gelu else 0, _stream()),
              "mt_gemm_f32x3")
    return (y, ys) if split_out else y


def attention_f32(qkv: torch.Tensor, num_heads: int, scale: Optional[float] = None):
    """fp32 attention core on qkv [B,N,3*H*dh] -> (big, small) pair of the [B,N,H*dh] output."""
    _f32(qkv)
    B, N, C3 = qkv.shape
    dh = C3 // (3 * num_heads)
    if scale is None:
        scale = dh ** -0.5
    ob = torch.empty((B, N, num_heads * dh), dtype=torch.float32, device=qkv.device)
    os_ = torch.empty_like(ob)
    with torch.cuda.device(qkv.device):
        check(_lib.load().mt_attention_f32(qkv.data_ptr(), ob.data_ptr(), os_.data_ptr(), B, N, num_heads, dh, float(scale),
                                           _stream()), "mt_attention_f32")
    return ob, os_


# ------------------------------------------------------------------------------------------------------------------
# fp8 (e4m3) opt-in Linear
# ------------------------------------------------------------------------------------------------------------------
def quantize_fp8(x: torch.Tensor, scale: float) -> torch.Tensor:
    """bf16 -> e4m3 bytes (uint8 tensor of the same shape): q = sat(x / scale), so that x ~ scale * q."""
    _require_cuda(x)
    if x.dtype != torch.bfloat16 or x.numel() % 8:
        raise ValueError("quantize_fp8: bf16 tensor with numel % 8 == 0")
    q = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().mt_quantize_fp8(x.data_ptr(), q.data_ptr(), x.numel(), float(scale), _stream()), "mt_quantize_fp8")
    return q


def linear_fp8(xq: torch.Tensor, wq: torch.Tensor, scale: float, bias: Optional[torch.Tensor] = None,
               residual: Optional[torch.Tensor] = None, gelu: bool = False) -> torch.Tensor:
    """y (bf16) = act(scale * (xq @ wq.T) + bias) (+ residual) on e4m3 operands (uint8 tensors [M,K], [N,K])."""
    _require_cuda(xq, wq, bias, residual)
    if xq.dtype != torch.uint8 or wq.dtype != torch.uint8:
        raise ValueError("linear_fp8: operands are uint8 tensors holding e4m3 bytes (see quantize_fp8)")
    M, K = xq.shape
    N = wq.shape[0]
    y = torch.empty((M, N), dtype=torch.bfloat16, device=xq.device)
    with torch.cuda.device(xq.device):
        check(_lib.load().mt_gemm_fp8(xq.data_ptr(), wq.data_ptr(), _ptr(bias), _ptr(residual), y.data_ptr(), M, N, K,
                                      1 if gelu else 0, float(scale), _stream()), "mt_gemm_fp8")
    return y
