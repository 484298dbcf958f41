"""Data2Seq/Image.py drop-in: 2-D image -> patch tokens (image / infrared / x-ray modalities)."""
from torch import nn

from ._conv_tokens import conv_tokens


class PatchEmbed(nn.Module):
    """2D Image to Patch Embedding (Data2Seq/Image.py:4-28): Conv2d(in_c, embed_dim, k=patch, s=patch) ->
    (B, N, embed_dim).  Like the reference, `norm` is constructed but NOT applied (Image.py:27 is commented out)."""

    def __init__(self, img_size=224, patch_size=16, in_c=3, embed_dim=768, norm_layer=None, token_dtype=None):
        super().__init__()
        img_size = (img_size, img_size)
        patch_size = (patch_size, patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_c, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()
        self.token_dtype = token_dtype

    def forward(self, x, pos_embed=None, prefix_tokens=None):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        k = (1,) + tuple(self.patch_size)
        return conv_tokens(self, self.proj, x.unsqueeze(2), k, k, self.token_dtype, pos_embed, prefix_tokens)
