"""metatransformer_b200 — B200-native (sm_100a) implementation of Meta-Transformer's shared encoder hot path.

Public surface (mirrors the reference's):
    Block, Attention, Mlp          timm.models.vision_transformer-compatible modules
    Data2Seq                       Data2Seq(modality, dim)(data) -> (B, N, dim)
    build_encoder, FusedEncoder    README encoder builder and the one-call whole-encoder fast path
    ClassifierHead, pool_norm      final LayerNorm + pooling (+ linear head) after the encoder
    ops                            one-kernel functional wrappers over the C ABI (include/mt_b200.h)
"""
from .vision_transformer import Attention, Block, DropPath, Mlp  # noqa: F401
from .data2seq import Data2Seq  # noqa: F401
from .encoder import ENCODER_CONFIGS, FusedEncoder, build_encoder  # noqa: F401
from .heads import ClassifierHead, pool_norm  # noqa: F401

__version__ = "0.1.0"
