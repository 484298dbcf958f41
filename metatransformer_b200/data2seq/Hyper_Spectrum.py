"""Data2Seq/Hyper_Spectrum.py drop-in: hyperspectral band patches -> tokens.

The reference class cannot run (`self.cls_tokens` / `embed_dim=1` kwarg to cat / `x = self.dropout`,
Data2Seq/Hyper_Spectrum.py:21-24); the working form is SpectralFormer's front end
(Hyper-spectrum/metatransformer.py:115-118, 152-159): Linear(patch_dim, d) -> [cls ; tokens] + pos[:, :n+1] -> dropout.
Constructor arguments / parameter names follow Data2Seq/Hyper_Spectrum.py:7-14.
"""
import torch
from torch import nn

from .. import ops
from ..vision_transformer import _packed


class PatchEmbed(nn.Module):
    def __init__(self, near_band=1, img_size=224, patch_size=16, embed_dim=768, dropout=0, token_dtype=None):
        super().__init__()
        patch_dim = img_size ** 2 * near_band
        self.pos_embedding = nn.Parameter(torch.randn(1, patch_size + 1, embed_dim))
        self.patch_to_embedding = nn.Linear(patch_dim, embed_dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.dropout = nn.Dropout(dropout)
        self.token_dtype = token_dtype

    def forward(self, x, mask=None):
        # x: [b, n, patch_dim]  ("Assure that x is like [b, c, (h w)]", Hyper_Spectrum.py:18)
        if not x.is_cuda:
            raise RuntimeError("Data2Seq tokenizers run on B200 GPUs only; there is no CPU fallback")
        if self.training and self.dropout.p > 0:
            raise NotImplementedError("Hyper_Spectrum.PatchEmbed: dropout in training mode is not built; use .eval()")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError("hyperspectral tokenizer backward is not built yet; use torch.no_grad()")
        b, n, k = x.shape
        if n + 1 > self.pos_embedding.shape[1]:
            raise ValueError(f"{n} patches but the positional table holds {self.pos_embedding.shape[1] - 1}")
        if k % 8 != 0:
            raise NotImplementedError("patch_dim must be a multiple of 8 on the B200 path")
        x2 = x.reshape(b * n, k).contiguous()
        x2 = x2 if x2.dtype == torch.bfloat16 else ops.cast(x2 if x2.dtype == torch.float32 else x2.float(), torch.bfloat16)
        w = _packed(self, "w", self.patch_to_embedding.weight, torch.bfloat16)
        bias = _packed(self, "b", self.patch_to_embedding.bias, torch.float32)
        pos = _packed(self, "pos", self.pos_embedding, torch.bfloat16)[0, :n + 1].contiguous()      # [n+1, d]
        cls = _packed(self, "cls", self.cls_token, torch.bfloat16).view(1, -1)
        tok = ops.linear(x2, w, bias)                                  # [b*n, d] = patches W^T + b
        out = ops.assemble_tokens(tok.view(b, n, -1), cls, pos)        # [b, n+1, d] = [cls ; tok] + pos[:n+1]
        out_dtype = self.token_dtype or (x.dtype if x.dtype in (torch.float32, torch.bfloat16, torch.float16) else torch.float32)
        return out if out_dtype == torch.bfloat16 else ops.cast(out, out_dtype)
