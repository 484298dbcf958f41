"""Data2Seq/Time_Series.py drop-in: time-series / IMU -> tokens.

`DataEmbedding(c_in, d_model)` = TokenEmbedding (Conv1d(c_in, d, k=3, padding=1, circular, bias=False) over time,
Data2Seq/Time_Series.py:29-42) + fixed sinusoid PositionalEmbedding (:8-26) [+ temporal embedding when x_mark is given]
-> Dropout (:109-126).  On the B200 path the conv is a circular im2row kernel + the tcgen05 GEMM whose residual input
adds the positional table (plus, with x_mark, the calendar embeddings summed by one embedding-bag kernel, or a second
small GEMM for the `timeF` Linear form); the training dropout is the Philox kernel, the weight gradients are the
MN-major split-K GEMM.
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


class FixedEmbedding(nn.Module):
    """Sinusoid lookup table held as a frozen nn.Embedding (Data2Seq/Time_Series.py:45-62): same state-dict key
    (`emb.weight`) as the reference."""

    def __init__(self, c_in, d_model):
        super().__init__()
        w = torch.zeros(c_in, d_model).float()
        position = torch.arange(0, c_in).float().unsqueeze(1)
        div_term = (torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model)).exp()
        w[:, 0::2] = torch.sin(position * div_term)
        w[:, 1::2] = torch.cos(position * div_term)
        self.emb = nn.Embedding(c_in, d_model)
        self.emb.weight = nn.Parameter(w, requires_grad=False)


class TemporalEmbedding(nn.Module):
    """month / day / weekday / hour (/ minute) calendar embeddings (Data2Seq/Time_Series.py:64-92), summed by one
    embedding-bag kernel over the concatenated tables.  x_mark columns: 0 month, 1 day, 2 weekday, 3 hour, 4 minute."""

    def __init__(self, d_model, embed_type='fixed', freq='h'):
        super().__init__()
        Embed = FixedEmbedding if embed_type == 'fixed' else nn.Embedding
        if freq == 't':
            self.minute_embed = Embed(4, d_model)
        self.hour_embed = Embed(24, d_model)
        self.weekday_embed = Embed(7, d_model)
        self.day_embed = Embed(32, d_model)
        self.month_embed = Embed(13, d_model)

    def tables(self):
        """[(x_mark column, weight)] in the order they are concatenated"""
        w = lambda m: m.emb.weight if isinstance(m, FixedEmbedding) else m.weight  # noqa: E731
        t = [(0, w(self.month_embed)), (1, w(self.day_embed)), (2, w(self.weekday_embed)), (3, w(self.hour_embed))]
        if hasattr(self, 'minute_embed'):
            t.append((4, w(self.minute_embed)))
        return t


class TimeFeatureEmbedding(nn.Module):
    """Linear(d_inp, d_model, bias=False) on real-valued time features (Data2Seq/Time_Series.py:95-106)."""

    def __init__(self, d_model, embed_type='timeF', freq='h'):
        super().__init__()
        freq_map = {'h': 4, 't': 5, 's': 6, 'm': 1, 'a': 1, 'w': 2, 'd': 3, 'b': 3}
        self.embed = nn.Linear(freq_map[freq], d_model, bias=False)


def _packed_conv_weight(module, w, Kp):
    """Conv1d weight [d, C, 3] -> bf16 [d, Kp] (zero-padded to the im2row row width), cached on the parameter."""
    cache = module.__dict__.setdefault("_mt_pack_cache", {})
    key = (w.data_ptr(), w._version, w.device, Kp)
    hit = cache.get("w2d")
    if hit is None or hit[0] != key:
        d = w.shape[0]
        w2 = torch.zeros((d, Kp), dtype=torch.float32, device=w.device)     # one-time weight packing
        w2[:, :w.shape[1] * 3] = w.detach().reshape(d, -1).float()
        hit = (key, ops.cast(w2, torch.bfloat16))
        cache["w2d"] = hit
    return hit[1]


class DataEmbedding(nn.Module):
    """value_embedding(x) + position_embedding(x) [+ temporal_embedding(x_mark)] -> dropout  (Time_Series.py:109-126).
    Same sub-module / parameter names as the reference, so its checkpoints load strict."""

    def __init__(self, c_in, d_model, embed_type='fixed', freq='h', dropout=0.1, token_dtype=None):
        super().__init__()
        self.value_embedding = TokenEmbedding(c_in=c_in, d_model=d_model)
        self.position_embedding = PositionalEmbedding(d_model=d_model)
        self.temporal_embedding = (TemporalEmbedding(d_model=d_model, embed_type=embed_type, freq=freq)
                                   if embed_type != 'timeF' else
                                   TimeFeatureEmbedding(d_model=d_model, embed_type=embed_type, freq=freq))
        self.dropout = nn.Dropout(p=dropout)
        self.token_dtype = token_dtype
        self._pe_cache = {}

    def _pe_rows(self, T, device, dtype=torch.bfloat16):
        key = (T, str(device), dtype)
        pe = self._pe_cache.get(key)
        if pe is None:
            pe = self.position_embedding.pe[0, :T].to(device).contiguous()                    # [T, d] table
            if dtype != torch.float32:
                pe = ops.cast(pe, dtype)
            if len(self._pe_cache) > 8:
                self._pe_cache.clear()
            self._pe_cache[key] = pe
        return pe

    def _temporal_plus_pe(self, x_mark, B, T, device):
        """[B*T, d] bf16 = PE[t] + temporal_embedding(x_mark)[b, t] (lookup tables) — or None for the Linear form."""
        te = self.temporal_embedding
        if not isinstance(te, TemporalEmbedding):
            return None
        tabs = te.tables()
        cache = self.__dict__.setdefault("_mt_pack_cache", {})
        key = tuple((w.data_ptr(), w._version, w.device) for _, w in tabs)
        hit = cache.get("temporal")
        if hit is None or hit[0] != key:
            table = torch.cat([w.detach().float() for _, w in tabs], dim=0).contiguous()
            sizes = [w.shape[0] for _, w in tabs]
            offs = torch.tensor([sum(sizes[:i]) for i in range(len(sizes))], dtype=torch.int32, device=device)
            hit = (key, table, offs)
            cache["temporal"] = hit
        cols = [c for c, _ in tabs]
        idx = x_mark.reshape(B * T, -1)[:, cols].to(torch.int32).contiguous()      # .long() in the reference (:86)
        return ops.embed_sum(hit[1], idx, hit[2], pos=self._pe_rows(T, device, torch.float32))

    def _forward_impl(self, w, w_time, x, x_mark):
        B, T, C = x.shape
        xin = x.contiguous()
        if xin.dtype not in (torch.float32, torch.bfloat16):
            xin = xin.float()
        rows = ops.im2row1d_circular(xin)                               # [B*T, Kp]
        w2 = _packed_conv_weight(self, w, rows.shape[1])
        d = w.shape[0]
        mark_rows = None
        if x_mark is None:
            tok = ops.linear(rows, w2, None, residual=self._pe_rows(T, x.device), residual_period=T)   # + PE[t]
        else:
            res = self._temporal_plus_pe(x_mark, B, T, x.device)
            if res is not None:
                tok = ops.linear(rows, w2, None, residual=res)
            else:
                # Linear time features: a second small GEMM whose residual input is the conv result
                tok = ops.linear(rows, w2, None, residual=self._pe_rows(T, x.device), residual_period=T)
                xm = x_mark.reshape(B * T, -1).float().contiguous()
                Km = (xm.shape[1] + 7) // 8 * 8
                mark_rows = ops.cast(ops.pad_cols(xm, Km), torch.bfloat16)
                wt = _packed(self.temporal_embedding, "embed.weight", w_time, torch.float32)
                wt2 = ops.cast(ops.pad_cols(wt, Km), torch.bfloat16)
                tok = ops.linear(mark_rows, wt2, None, residual=tok)
        drop_key = None
        if self.training and self.dropout.p > 0:
            from ..vision_transformer import _philox_reserve
            drop_key = _philox_reserve(x.device, tok.numel())
            tok = ops.dropout(tok, self.dropout.p, *drop_key, out=tok)
        out_dtype = self.token_dtype or xin.dtype
        out = tok if out_dtype == torch.bfloat16 else ops.cast(tok, out_dtype)
        return out.view(B, T, d), rows, mark_rows, drop_key

    def forward(self, x, x_mark=None):
        if not x.is_cuda:
            raise RuntimeError("Data2Seq tokenizers run on B200 GPUs only; there is no CPU fallback")
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("gradient with respect to the tokenizer INPUT is not implemented (the reference "
                                      "pipelines never need it); the weight gradients are")
        w = self.value_embedding.tokenConv.weight
        w_time = self.temporal_embedding.embed.weight if isinstance(self.temporal_embedding, TimeFeatureEmbedding) else None
        if x_mark is not None and isinstance(self.temporal_embedding, TemporalEmbedding) and torch.is_grad_enabled() and any(
                t.requires_grad for _, t in self.temporal_embedding.tables()):
            raise NotImplementedError("DataEmbedding: gradients of LEARNED calendar embeddings (embed_type != 'fixed') are "
                                      "not built; the default fixed tables and the timeF Linear are")
        if torch.is_grad_enabled() and (w.requires_grad or (x_mark is not None and w_time is not None and w_time.requires_grad)):
            return _DataEmbeddingFn.apply(w, w_time if x_mark is not None else None, self, x, x_mark)
        return self._forward_impl(w, w_time, x, x_mark)[0]


class _DataEmbeddingFn(torch.autograd.Function):
    """Trainable time-series tokenizer: dW of the circular Conv1d (and of the timeF Linear) = dTokens^T @ rows with the
    MN-major split-K tcgen05 GEMM, the training dropout mask replayed from its Philox key."""

    @staticmethod
    def forward(ctx, w, w_time, module, x, x_mark):
        tok, rows, mark_rows, drop_key = module._forward_impl(w, w_time, x, x_mark)
        ctx.module, ctx.rows, ctx.mark_rows, ctx.drop_key = module, rows, mark_rows, drop_key
        ctx.wshape, ctx.wdtype = w.shape, w.dtype
        ctx.tshape = None if w_time is None else (w_time.shape, w_time.dtype)
        return tok

    @staticmethod
    def backward(ctx, g):
        if ctx.rows is None:
            raise RuntimeError("DataEmbedding: backward through the same graph twice is not supported (retain_graph)")
        d = ctx.wshape[0]
        g2 = g.contiguous().view(-1, d)
        g2 = g2 if g2.dtype == torch.bfloat16 else ops.cast(g2, torch.bfloat16)
        if ctx.drop_key is not None:
            g2 = ops.dropout(g2, ctx.module.dropout.p, *ctx.drop_key)
        dw = dwt = None
        if ctx.needs_input_grad[0]:
            dw = ops.linear_wgrad(g2, ctx.rows)[:, :ctx.wshape[1] * 3].contiguous()          # drop the K padding
            dw = (dw if ctx.wdtype == torch.float32 else ops.cast(dw, ctx.wdtype)).view(ctx.wshape)
        if ctx.tshape is not None and ctx.needs_input_grad[1] and ctx.mark_rows is not None:
            dwt = ops.linear_wgrad(g2, ctx.mark_rows)[:, :ctx.tshape[0][1]].contiguous()
            dwt = (dwt if ctx.tshape[1] == torch.float32 else ops.cast(dwt, ctx.tshape[1])).view(ctx.tshape[0])
        ctx.rows = ctx.mark_rows = None
        return dw, dwt, None, None, None
