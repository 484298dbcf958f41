"""Data2Seq/Acoustic.py drop-in: spectrogram -> patch tokens (AST tokenizer).

The reference constructor crashes (nested-tuple kernel_size, Data2Seq/Acoustic.py:10,16); the working form is AST's
conv (Audio/src/models/ast_models.py:18-33, 86): Conv2d(1, d, (16,16), stride=(fstride, tstride)) on a
(B, 1, F, T) spectrogram -> (B, f*t, d), f = (F-16)//fstride + 1, t = (T-16)//tstride + 1.
"""
from torch import nn

from ._conv_tokens import conv_tokens


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=1, embed_dim=768, fstride=10, tstride=10, token_dtype=None):
        super().__init__()
        img_size = _to_2tuple(img_size)
        patch_size = _to_2tuple(patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0])
        self.fstride, self.tstride = fstride, tstride
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=(fstride, tstride))
        self.token_dtype = token_dtype

    def forward(self, x, pos_embed=None, prefix_tokens=None):
        # x: (B, in_chans, F, T)  (Audio/src/models/ast_models.py:152-155 builds it from (B, T, F))
        k = (1,) + tuple(self.patch_size)
        s = (1, self.fstride, self.tstride)
        return conv_tokens(self, self.proj, x.unsqueeze(2), k, s, self.token_dtype, pos_embed, prefix_tokens)
