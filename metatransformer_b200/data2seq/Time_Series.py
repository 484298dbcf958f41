"""Data2Seq/Time_Series.py drop-in: time-series / IMU -> tokens.

`DataEmbedding(c_in, d_model)` = TokenEmbedding (Conv1d(c_in, d, k=3, padding=1, circular, bias=False) over time,
Data2Seq/Time_Series.py:29-42) + fixed sinusoid PositionalEmbedding (:8-26) [+ temporal embedding when x_mark is given]
-> Dropout (:109-126).  On the B200 path the conv is a circular im2row kernel + the tcgen05 GEMM whose residual input
adds the positional table; x_mark (calendar features) and dropout p > 0 in training mode are not built.
Note: the shipped dispatcher calls `DataEmbedding(cin=1, ...)`, a TypeError (Data2Seq.py:36 vs Time_Series.py:110);
this package passes `c_in=1` as intended.
"""
import math

import torch
import torch.nn as nn

from .. import ops
from ..vision_transformer import _packed


class PositionalEmbedding(nn.Module):
    def __init__(self, d_model, max_len=5000):
        super().__init__()
        pe = torch.zeros(max_len, d_model).float()
        position = torch.arange(0, max_len).float().unsqueeze(1)
        div_term = (torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model)).exp()
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer('pe', pe.unsqueeze(0))

    def forward(self, x):
        return self.pe[:, :x.size(1)]


class TokenEmbedding(nn.Module):
    def __init__(self, c_in, d_model):
        super().__init__()
        self.tokenConv = nn.Conv1d(in_channels=c_in, out_channels=d_model, kernel_size=3, padding=1,
                                   padding_mode='circular', bias=False)
        nn.init.kaiming_normal_(self.tokenConv.weight, mode='fan_in', nonlinearity='leaky_relu')


class DataEmbedding(nn.Module):
    def __init__(self, c_in, d_model, embed_type='fixed', freq='h', dropout=0.1, token_dtype=None):
        super().__init__()
        self.value_embedding = TokenEmbedding(c_in=c_in, d_model=d_model)
        self.position_embedding = PositionalEmbedding(d_model=d_model)
        self.dropout = nn.Dropout(p=dropout)
        self.token_dtype = token_dtype
        self._pe_cache = {}

    def _pe_rows(self, T, device):
        key = (T, str(device))
        pe = self._pe_cache.get(key)
        if pe is None:
            pe = ops.cast(self.position_embedding.pe[0, :T].to(device).contiguous(), torch.bfloat16)   # [T, d] table
            if len(self._pe_cache) > 8:
                self._pe_cache.clear()
            self._pe_cache[key] = pe
        return pe

    def forward(self, x, x_mark=None):
        if x_mark is not None:
            raise NotImplementedError("DataEmbedding: temporal (x_mark) embeddings are not built on the B200 path")
        if self.training and self.dropout.p > 0:
            raise NotImplementedError("DataEmbedding: dropout in training mode is not built on the B200 path; use .eval()")
        if not x.is_cuda:
            raise RuntimeError("Data2Seq tokenizers run on B200 GPUs only; there is no CPU fallback")
        if torch.is_grad_enabled() and (x.requires_grad or self.value_embedding.tokenConv.weight.requires_grad):
            raise NotImplementedError("time-series tokenizer backward is not built yet; use torch.no_grad()")
        B, T, C = x.shape
        xin = x.contiguous()
        if xin.dtype not in (torch.float32, torch.bfloat16):
            xin = xin.float()
        rows = ops.im2row1d_circular(xin)                               # [B*T, Kp]
        w = self.value_embedding.tokenConv.weight                      # [d, C, 3]
        d, Kp = w.shape[0], rows.shape[1]
        cache = self.__dict__.setdefault("_mt_pack_cache", {})
        key = (w.data_ptr(), w._version, w.device, Kp)
        hit = cache.get("w2d")
        if hit is None or hit[0] != key:
            w2 = torch.zeros((d, Kp), dtype=torch.float32, device=w.device)     # one-time weight packing
            w2[:, :3 * C] = w.detach().reshape(d, 3 * C).float()
            hit = (key, ops.cast(w2, torch.bfloat16))
            cache["w2d"] = hit
        tok = ops.linear(rows, hit[1], None, residual=self._pe_rows(T, x.device), residual_period=T)   # + PE[t]
        out_dtype = self.token_dtype or xin.dtype
        if out_dtype != torch.bfloat16:
            tok = ops.cast(tok, out_dtype)
        return tok.view(B, T, d)
