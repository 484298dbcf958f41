"""Drop-in for the reference's image tokenizer (`Data2Seq/Image.py`; image / infrared / x-ray modalities).

Public surface kept from the reference class (`Data2Seq/Image.py:4-28`): constructor arguments, the attributes
`img_size`, `patch_size`, `grid_size`, `num_patches`, `proj` (a `Conv2d(in_c, embed_dim, patch, stride=patch)`, so its
state-dict keys match) and `norm` — which, as in the reference (`:27` is commented out), exists but is never applied.
The forward itself is different: patches are gathered by the im2row kernel and projected by the tcgen05 GEMM, the result
is a contiguous (B, N, embed_dim) tensor, and the positional table / prefix tokens can be fused in (see _conv_tokens).
"""
from torch import nn

from ._conv_tokens import conv_tokens


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_c=3, embed_dim=768, norm_layer=None, token_dtype=None):
        super().__init__()
        side, patch = int(img_size), int(patch_size)
        per_side = side // patch
        self.img_size, self.patch_size = (side, side), (patch, patch)
        self.grid_size, self.num_patches = (per_side, per_side), per_side * per_side
        self.proj = nn.Conv2d(in_c, embed_dim, self.patch_size, self.patch_size)
        self.norm = nn.Identity() if norm_layer is None else norm_layer(embed_dim)
        self.token_dtype = token_dtype

    def forward(self, x, pos_embed=None, prefix_tokens=None):
        height, width = int(x.shape[-2]), int(x.shape[-1])
        if (height, width) != self.img_size:
            raise AssertionError(f"Input image size ({height}*{width}) doesn't match model "
                                 f"({self.img_size[0]}*{self.img_size[1]}).")
        window = (1,) + self.patch_size                      # (kt, kh, kw); stride == window for non-overlapping patches
        return conv_tokens(self, self.proj, x.unsqueeze(2), window, window, self.token_dtype, pos_embed, prefix_tokens)
