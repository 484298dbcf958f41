"""The Python drop-in surface: constructor signatures, attribute / state-dict names, strict loading of
a reference-produced state dict, loud failure off-GPU.  CPU only."""
import inspect

import numpy as np
import pytest
import torch
import torch.nn as nn

import metatransformer_b200 as mt
from metatransformer_b200 import data2seq


def test_block_signature_matches_timm_0_4_12():
    sig = inspect.signature(mt.Block.__init__)
    names = list(sig.parameters)[1:]
    assert names[:10] == ["dim", "num_heads", "mlp_ratio", "qkv_bias", "qk_scale", "drop", "attn_drop", "drop_path",
                          "act_layer", "norm_layer"]
    assert sig.parameters["mlp_ratio"].default == 4.0 and sig.parameters["qkv_bias"].default is False
    a = inspect.signature(mt.Attention.__init__)
    assert list(a.parameters)[1:] == ["dim", "num_heads", "qkv_bias", "qk_scale", "attn_drop", "proj_drop"]
    m = inspect.signature(mt.Mlp.__init__)
    assert list(m.parameters)[1:] == ["in_features", "hidden_features", "out_features", "act_layer", "drop"]


def test_readme_encoder_builds_and_loads_reference_state_dict(golden_dir):
    z = np.load(golden_dir / "enc_d128_h2_L2.npz")
    dim, heads, depth, B, N = [int(v) for v in z["meta"]]
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    enc = nn.Sequential(*[mt.Block(dim=dim, num_heads=heads, mlp_ratio=4., qkv_bias=True, norm_layer=nn.LayerNorm,
                                   act_layer=nn.GELU) for _ in range(depth)])
    missing = enc.load_state_dict(sd, strict=True)     # README.md:135
    assert not missing.missing_keys and not missing.unexpected_keys
    blk = enc[0]
    for attr in ("norm1", "attn", "drop_path", "norm2", "mlp"):
        assert hasattr(blk, attr)
    assert blk.attn.qkv.weight.shape == (3 * dim, dim) and blk.mlp.fc1.weight.shape == (4 * dim, dim)
    assert blk.attn.scale == (dim // heads) ** -0.5 and blk.norm1.eps == 1e-5


def test_parameter_count_matches_published_sizes():
    base = mt.build_encoder("base")
    assert sum(p.numel() for p in base.parameters()) == 85_054_464        # README.md:103 "85M"
    assert mt.ENCODER_CONFIGS["large"] == dict(dim=1024, num_heads=16, depth=24)


def test_no_cpu_fallback():
    blk = mt.Block(64, 1, qkv_bias=True)
    with pytest.raises(RuntimeError, match="no CPU"):
        blk(torch.zeros(1, 4, 64))
    with pytest.raises(RuntimeError, match="no CPU"):
        mt.Data2Seq("image", 64)(torch.zeros(1, 3, 224, 224))
    with pytest.raises(RuntimeError, match="no CPU"):
        mt.FusedEncoder(nn.Sequential(blk))(torch.zeros(1, 4, 64))


def test_unsupported_configurations_raise():
    with pytest.raises(NotImplementedError):
        mt.Block(64, 1, act_layer=nn.ReLU)
    with pytest.raises(NotImplementedError):
        mt.Block(64, 1, norm_layer=nn.BatchNorm1d)
    gt = mt.Data2Seq("graph", 768).embed
    assert isinstance(gt, data2seq.Graph.GraphFeatureTokenizer) and gt.encoder_embed_dim == 768
    assert {"atom_encoder.weight", "edge_encoder.weight", "graph_token.weight", "null_token.weight", "orf_encoder.weight",
            "order_encoder.weight"} <= set(gt.state_dict().keys())
    assert isinstance(mt.Data2Seq("text", 768).embed, data2seq.Text.TextFeaturePad)
    with pytest.raises(NotImplementedError):
        data2seq.Text.get_text_embeddings(["a sentence"])        # the CLIP tower itself is not part of the path
    with pytest.raises(ValueError):
        mt.Data2Seq("smell", 768)


def test_data2seq_surface():
    t = mt.Data2Seq("image", 768)
    assert t.modality == "image" and t.embed_dim == 768
    pe = t.embed
    assert isinstance(pe, data2seq.Image.PatchEmbed)
    assert pe.img_size == (224, 224) and pe.patch_size == (16, 16) and pe.grid_size == (14, 14) and pe.num_patches == 196
    assert pe.proj.weight.shape == (768, 3, 16, 16) and pe.proj.bias.shape == (768,)
    v = mt.Data2Seq("video", 1024).embed
    assert v.proj.weight.shape == (1024, 3, 2, 16, 16) and v.num_patches == 1568
    a = mt.Data2Seq("audio", 768).embed
    assert a.proj.weight.shape == (768, 1, 16, 16) and a.proj.stride == (10, 10)
    hy = mt.Data2Seq("hyper", 768).embed
    assert hy.patch_to_embedding.weight.shape == (768, 224 * 224) and hy.pos_embedding.shape == (1, 17, 768)
    ts = mt.Data2Seq("imu", 768).embed
    assert ts.value_embedding.tokenConv.weight.shape == (768, 1, 3) and ts.position_embedding.pe.shape == (1, 5000, 768)


def test_compat_shim_resolves_reference_imports():
    import metatransformer_b200.compat as compat
    compat.install()
    from Data2Seq import Data2Seq          # README.md:117
    assert Data2Seq is mt.Data2Seq
    import sys
    if getattr(sys.modules.get("timm"), "_mt_b200_shim", False):   # real timm absent -> shim installed
        from timm.models.vision_transformer import Block    # README.md:116
        assert Block is mt.Block
    compat.install()                                        # idempotent


def test_classifier_head_surface_and_no_cpu_fallback():
    head = mt.ClassifierHead(64, 10, pool="cls_dist")
    assert sorted(k for k, _ in head.named_parameters()) == ["head.bias", "head.weight", "norm.bias", "norm.weight"]
    assert head.head.weight.shape == (10, 64) and head.norm.eps == 1e-5
    feat = mt.ClassifierHead(64, 0, pool="mean")
    assert isinstance(feat.head, nn.Identity)
    with pytest.raises(ValueError):
        mt.ClassifierHead(64, 10, pool="max")
    with pytest.raises(NotImplementedError):
        mt.ClassifierHead(64, 10, norm_layer=nn.BatchNorm1d)
    with pytest.raises(RuntimeError, match="no CPU"):
        head(torch.zeros(2, 5, 64))


def test_tokenizer_forward_signatures_accept_pos_embed_and_prefix_tokens():
    import inspect
    for mod in (data2seq.Image, data2seq.Acoustic, data2seq.Video):
        params = inspect.signature(mod.PatchEmbed.forward).parameters
        assert "pos_embed" in params and "prefix_tokens" in params
        assert params["pos_embed"].default is None and params["prefix_tokens"].default is None   # reference call unchanged
    with pytest.raises(RuntimeError, match="no CPU"):
        data2seq.Text.zero_padding(torch.zeros(2, 512), 768)


def test_block_dropout_configuration_is_accepted_and_attn_drop_documented():
    blk = mt.Block(64, 2, qkv_bias=True, drop=0.1, attn_drop=0.0, drop_path=0.2)
    assert blk.attn.proj_drop.p == 0.1 and blk.mlp.drop.p == 0.1 and blk.drop_path.drop_prob == 0.2
    assert blk._is_stochastic() and not blk.eval()._is_stochastic()
    assert mt.Block(64, 2)._is_stochastic() is False


def test_package_synthetic_generators_match_the_oracle_copy():
    """bench.py / tools seed their weights from metatransformer_b200.synthetic (nothing measured imports oracle/); the
    oracle's own generators must stay bit-identical so that parity checks compare like with like."""
    from metatransformer_b200 import synthetic
    from oracle import encoder_oracle as orc
    a, b = synthetic.make_encoder_state_dict(64, 2, seed=3), orc.make_encoder_state_dict(64, 2, seed=3)
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(synthetic.make_tokens(2, 5, 64, seed=7), orc.make_tokens(2, 5, 64, seed=7))
    assert synthetic.encoder_flops(256, 196, 768, 12) == orc.encoder_flops(256, 196, 768, 12) == 256 * 196 * 177094656.0
