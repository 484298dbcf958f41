"""Data2Seq/Text.py drop-in: text features -> one token per sentence, zero-padded to the encoder width.

The reference runs the whole CLIP ViT-B/32 text tower (an external pip package with downloaded weights, Text.py:7-11)
and then pads its 512-wide sentence embedding with zeros up to `tar_dim` (Text.py:15-19).  The CLIP tower is outside the
hot path (and `clip` is not installable here), so the B200 side takes the sentence embeddings as INPUT and does the one
thing the reference's own code does to them; `get_text_embeddings` keeps its name and accepts a caller-supplied encoder.
"""
import torch
from torch import nn

from .. import ops


def zero_padding(text_tensor, tar_dim, device=None):
    """(B, c) fp32 / bf16 CUDA features -> (B, tar_dim) with zeros appended (Data2Seq/Text.py:15-19)."""
    if not text_tensor.is_cuda:
        raise RuntimeError("Data2Seq tokenizers run on B200 GPUs only; there is no CPU fallback")
    if text_tensor.dim() != 2 or text_tensor.shape[1] > tar_dim:
        raise ValueError(f"zero_padding: expected (B, c <= {tar_dim}) features, got {tuple(text_tensor.shape)}")
    x = text_tensor.contiguous()
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    return ops.pad_cols(x, tar_dim)


def get_text_embeddings(text, tar_dim=768, encode_text=None):
    """`encode_text(text) -> (B, c)` stands in for clip.load('ViT-B/32').encode_text(clip.tokenize(text)) (Text.py:6-11)."""
    if encode_text is None:
        raise NotImplementedError("the CLIP text tower is an external dependency outside the hot path; pass "
                                  "encode_text=<callable returning (B, c) CUDA features>")
    return zero_padding(encode_text(text), tar_dim)


class TextFeaturePad(nn.Module):
    """`embed` of Data2Seq(modality='text'): (B, c) sentence features -> (B, 1, dim) tokens."""

    def __init__(self, embed_dim=768):
        super().__init__()
        self.embed_dim = embed_dim

    def forward(self, feats):
        return zero_padding(feats, self.embed_dim).unsqueeze(1)
