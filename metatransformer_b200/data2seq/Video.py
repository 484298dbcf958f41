"""Data2Seq/Video.py drop-in: video clip -> tubelet tokens.

The reference file cannot run (Conv3d built with a 2-tuple kernel, 4-way unpack of a 5-D input, Data2Seq/Video.py:16,20);
the working form it was adapted from is VideoMAE's PatchEmbed (Video/models/modeling_finetune.py:263-297), which is
what this class implements: Conv3d(in_chans, d, (tubelet, p, p), stride = kernel) -> (B, T/tubelet * (H/p)*(W/p), d).
"""
from torch import nn

from ._conv_tokens import conv_tokens


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, num_frames=16, tubelet_size=2,
                 in_c=None, norm_layer=None, token_dtype=None):
        super().__init__()
        if in_c is not None:          # Data2Seq/Video.py:8 spells it `in_c`
            in_chans = in_c
        img_size = (img_size, img_size)
        patch_size = (patch_size, patch_size)
        num_spatial_patches = (img_size[0] // patch_size[0]) * (img_size[1] // patch_size[1])
        self.img_size = img_size
        self.tubelet_size = tubelet_size
        self.patch_size = patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = num_spatial_patches * (num_frames // tubelet_size)
        self.proj = nn.Conv3d(in_channels=in_chans, out_channels=embed_dim,
                              kernel_size=(tubelet_size, patch_size[0], patch_size[1]),
                              stride=(tubelet_size, patch_size[0], patch_size[1]))
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()
        self.token_dtype = token_dtype

    def forward(self, x, pos_embed=None, prefix_tokens=None, **kwargs):
        B, C, T, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        k = (self.tubelet_size,) + tuple(self.patch_size)
        return conv_tokens(self, self.proj, x, k, k, self.token_dtype, pos_embed, prefix_tokens)
