"""ctypes binding of libmt_b200.so (the C ABI declared in include/mt_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception is raised.
The product path never routes through PyTorch eager math or through oracle/.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_ulonglong, c_void_p, POINTER, Structure
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libmt_b200.so"

MT_DT_BF16, MT_DT_F32, MT_DT_F16 = 0, 1, 2

EXPORTED_SYMBOLS = (
    "mt_version", "mt_last_error", "mt_device_ok", "mt_gemm_bf16", "mt_gemm_bf16_ln", "mt_gemm_bf16_ex", "mt_gemm_bf16_splitk", "mt_gemm_bf16_wgrad", "mt_row_stats",
    "mt_fold_layernorm", "mt_layernorm", "mt_layernorm_bwd", "mt_gelu", "mt_transpose_bf16", "mt_colsum", "mt_ln_fold_bwd", "mt_attention_fwd",
    "mt_attention_fwd_lse", "mt_attention_bwd_workspace_bytes", "mt_attention_bwd",
    "mt_im2row", "mt_im2row1d_circular", "mt_prepend_row", "mt_assemble_tokens", "mt_pool_layernorm", "mt_pool_layernorm_bwd",
    "mt_dropout", "mt_pad_cols", "mt_cast", "mt_encoder_create", "mt_encoder_destroy", "mt_encoder_workspace_bytes",
    "mt_encoder_forward", "mt_encoder_forward_ex", "mt_encoder_launches_per_forward",
    "mt_row_stats_split", "mt_ln_stats_finalize", "mt_center_split", "mt_embed_sum", "mt_split_f32", "mt_merge_f32", "mt_timeline_begin", "mt_timeline_end", "mt_allreduce", "mt_attention_fwd_ex", "mt_attention_bwd_ex",
    "mt_attention_dropout_mask", "mt_attention_set_trace", "mt_gemm_set_trace", "mt_quantize_fp8", "mt_gemm_fp8", "mt_split_tf32", "mt_layernorm_f32_split", "mt_gemm_f32x3", "mt_attention_f32",
)


class MtError(RuntimeError):
    pass


class GemmOpts(Structure):
    """mirror of `mt_gemm_opts`"""
    _fields_ = [
        ("bias", c_void_p), ("residual", c_void_p), ("act", c_int),
        ("ln_colsum", c_void_p), ("ln_row_stats", c_void_p), ("ln_stats_are_sums", c_int), ("ln_eps", c_float),
        ("out_row_sums", c_void_p), ("cta_group", c_int), ("residual_period", c_int),
        ("residual_lo", c_void_p), ("y_lo", c_void_p), ("y_row_offset", c_void_p),
    ]


class BlockWeights(Structure):
    """mirror of `mt_block_weights`"""
    _fields_ = [
        ("ln1_w", c_void_p), ("ln1_b", c_void_p),
        ("qkv_w", c_void_p), ("qkv_b", c_void_p),
        ("proj_w", c_void_p), ("proj_b", c_void_p),
        ("ln2_w", c_void_p), ("ln2_b", c_void_p),
        ("fc1_w", c_void_p), ("fc1_b", c_void_p),
        ("fc2_w", c_void_p), ("fc2_b", c_void_p),
        ("qkv_colsum", c_void_p), ("fc1_colsum", c_void_p),
    ]


_lib = None


def _declare(lib: ctypes.CDLL) -> None:
    lib.mt_version.restype = c_int
    lib.mt_version.argtypes = []
    lib.mt_last_error.restype = c_char_p
    lib.mt_last_error.argtypes = []
    lib.mt_device_ok.restype = c_int
    lib.mt_device_ok.argtypes = [c_int]
    lib.mt_gemm_bf16.restype = c_int
    lib.mt_gemm_bf16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p]
    lib.mt_gemm_bf16_ln.restype = c_int
    lib.mt_gemm_bf16_ln.argtypes = [c_void_p] * 6 + [c_int] * 5 + [c_void_p]
    lib.mt_gemm_bf16_ex.restype = c_int
    lib.mt_gemm_bf16_ex.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, POINTER(GemmOpts), c_void_p]
    lib.mt_gemm_bf16_splitk.restype = c_int
    lib.mt_gemm_bf16_splitk.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    lib.mt_gemm_bf16_wgrad.restype = c_int
    lib.mt_gemm_bf16_wgrad.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    lib.mt_row_stats.restype = c_int
    lib.mt_row_stats.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_float, c_void_p]
    lib.mt_fold_layernorm.restype = c_int
    lib.mt_fold_layernorm.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int, c_int, c_void_p]
    lib.mt_layernorm.restype = c_int
    lib.mt_layernorm.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]
    lib.mt_layernorm_bwd.restype = c_int
    lib.mt_layernorm_bwd.argtypes = [c_void_p] * 6 + [c_int, c_int, c_void_p]
    lib.mt_gelu.restype = c_int
    lib.mt_gelu.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
    lib.mt_transpose_bf16.restype = c_int
    lib.mt_transpose_bf16.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    lib.mt_colsum.restype = c_int
    lib.mt_colsum.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p]
    lib.mt_ln_fold_bwd.restype = c_int
    lib.mt_ln_fold_bwd.argtypes = [c_void_p, c_void_p, c_int] + [c_void_p] * 6 + [c_int, c_int, c_void_p]
    lib.mt_attention_fwd.restype = c_int
    lib.mt_attention_fwd.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]
    lib.mt_attention_fwd_lse.restype = c_int
    lib.mt_attention_fwd_lse.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]
    lib.mt_attention_bwd_workspace_bytes.restype = c_int
    lib.mt_attention_bwd_workspace_bytes.argtypes = [c_int, c_int, c_int, POINTER(c_size_t)]
    lib.mt_attention_bwd.restype = c_int
    lib.mt_attention_bwd.argtypes = [c_void_p] * 6 + [c_size_t, c_int, c_int, c_int, c_int, c_float, c_void_p]
    lib.mt_im2row.restype = c_int
    lib.mt_im2row.argtypes = [c_void_p, c_int, c_void_p] + [c_int] * 11 + [c_void_p]
    lib.mt_im2row1d_circular.restype = c_int
    lib.mt_im2row1d_circular.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    lib.mt_prepend_row.restype = c_int
    lib.mt_prepend_row.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    lib.mt_assemble_tokens.restype = c_int
    lib.mt_assemble_tokens.argtypes = [c_void_p] * 4 + [c_int] * 4 + [c_void_p]
    lib.mt_pool_layernorm.restype = c_int
    lib.mt_pool_layernorm.argtypes = [c_void_p] * 4 + [c_int] * 4 + [c_float, c_void_p]
    lib.mt_pool_layernorm_bwd.restype = c_int
    lib.mt_pool_layernorm_bwd.argtypes = [c_void_p] * 6 + [c_int] * 4 + [c_float, c_void_p]
    lib.mt_dropout.restype = c_int
    lib.mt_dropout.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_longlong, c_int, c_float, c_ulonglong,
                               c_ulonglong, c_void_p]
    lib.mt_pad_cols.restype = c_int
    lib.mt_pad_cols.argtypes = [c_void_p, c_void_p, c_int, c_longlong, c_int, c_int, c_void_p]
    lib.mt_cast.restype = c_int
    lib.mt_cast.argtypes = [c_void_p, c_int, c_void_p, c_int, c_size_t, c_void_p]
    lib.mt_encoder_create.restype = c_int
    lib.mt_encoder_create.argtypes = [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_float, POINTER(BlockWeights)]
    lib.mt_encoder_destroy.restype = c_int
    lib.mt_encoder_destroy.argtypes = [c_void_p]
    lib.mt_encoder_workspace_bytes.restype = c_int
    lib.mt_encoder_workspace_bytes.argtypes = [c_void_p, c_int, c_int, POINTER(c_size_t)]
    lib.mt_encoder_forward.restype = c_int
    lib.mt_encoder_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]
    lib.mt_encoder_forward_ex.restype = c_int
    lib.mt_encoder_forward_ex.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t,
                                          c_void_p]
    lib.mt_row_stats_split.restype = c_int
    lib.mt_row_stats_split.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]
    lib.mt_ln_stats_finalize.restype = c_int
    lib.mt_ln_stats_finalize.argtypes = [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]
    lib.mt_center_split.restype = c_int
    lib.mt_center_split.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]
    lib.mt_embed_sum.restype = c_int
    lib.mt_embed_sum.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_longlong, c_int,
                                 c_void_p]
    lib.mt_split_f32.restype = c_int
    lib.mt_split_f32.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.mt_merge_f32.restype = c_int
    lib.mt_merge_f32.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.mt_attention_fwd_ex.restype = c_int
    lib.mt_attention_fwd_ex.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float,
                                        c_ulonglong, c_void_p]
    lib.mt_attention_bwd_ex.restype = c_int
    lib.mt_attention_bwd_ex.argtypes = [c_void_p] * 6 + [c_size_t, c_int, c_int, c_int, c_int, c_float, c_float, c_ulonglong,
                                                         c_void_p]
    lib.mt_gemm_set_trace.restype = c_int
    lib.mt_gemm_set_trace.argtypes = [c_void_p]
    lib.mt_attention_set_trace.restype = c_int
    lib.mt_attention_set_trace.argtypes = [c_void_p]
    lib.mt_attention_dropout_mask.restype = c_int
    lib.mt_attention_dropout_mask.argtypes = [c_void_p, c_int, c_int, c_int, c_float, c_ulonglong, c_void_p]
    lib.mt_quantize_fp8.restype = c_int
    lib.mt_quantize_fp8.argtypes = [c_void_p, c_void_p, c_size_t, c_float, c_void_p]
    lib.mt_gemm_fp8.restype = c_int
    lib.mt_gemm_fp8.argtypes = [c_void_p] * 5 + [c_int, c_int, c_int, c_int, c_float, c_void_p]
    lib.mt_split_tf32.restype = c_int
    lib.mt_split_tf32.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.mt_layernorm_f32_split.restype = c_int
    lib.mt_layernorm_f32_split.argtypes = [c_void_p] * 5 + [c_int, c_int, c_float, c_void_p]
    lib.mt_gemm_f32x3.restype = c_int
    lib.mt_gemm_f32x3.argtypes = [c_void_p] * 8 + [c_int, c_int, c_int, c_int, c_void_p]
    lib.mt_attention_f32.restype = c_int
    lib.mt_attention_f32.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]
    lib.mt_allreduce.restype = c_int
    lib.mt_allreduce.argtypes = [c_void_p, c_void_p, c_size_t, c_int, c_float, c_void_p]
    lib.mt_timeline_begin.restype = c_int
    lib.mt_timeline_begin.argtypes = [c_void_p, c_int]
    lib.mt_timeline_end.restype = c_int
    lib.mt_timeline_end.argtypes = []
    lib.mt_encoder_launches_per_forward.restype = c_int
    lib.mt_encoder_launches_per_forward.argtypes = [c_void_p]


def load() -> ctypes.CDLL:
    """Load (once) and return the library; raise if it has not been built (`python -m metatransformer_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("MT_B200_LIB", LIB_PATH))
    if not path.exists():
        raise MtError(
            f"{path} not found: the sm_100a extension has not been built. Run `python -m metatransformer_b200.build` "
            "(needs nvcc). There is no CPU / eager fallback for this path.")
    import torch  # noqa: F401  (loads torch's libcudart.so.12 first so both share one CUDA runtime instance)
    lib = ctypes.CDLL(str(path), mode=ctypes.RTLD_GLOBAL)
    _declare(lib)
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().mt_last_error().decode("utf-8", "replace")
        raise MtError(f"{what} failed with status {rc}: {msg}")
