"""Pin the CPU oracle (oracle/encoder_oracle.py) against vectors produced by the reference's own code
(tests/golden/make_golden.py, which executed /root/reference's in-tree Block / Image.PatchEmbed / DataEmbedding)."""
import numpy as np
import pytest
import torch

from oracle import encoder_oracle as orc


def _load(golden_dir, name):
    z = np.load(golden_dir / name)
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.mark.parametrize("name", ["enc_d128_h2_L2.npz", "enc_d96_h3_L2.npz"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.float64, 2e-5)])
def test_encoder_matches_reference_blocks(golden_dir, name, dtype, tol):
    g = _load(golden_dir, name)
    dim, heads, depth, B, N = [int(v) for v in g["meta"]]
    sd = {k[3:]: v.to(dtype) for k, v in g.items() if k.startswith("sd.")}
    x = g["x"].to(dtype)
    y = orc.encoder(x, sd, depth, heads)
    assert y.shape == (B, N, dim)
    # golden is fp32 output of the reference modules; fp64 oracle differs only by fp32 rounding of the reference
    assert (y.float() - g["y"]).abs().max().item() < tol * max(1.0, g["y"].abs().max().item())


def test_block_stages_match_reference(golden_dir):
    g = _load(golden_dir, "enc_d128_h2_L2.npz")
    dim, heads, depth, B, N = [int(v) for v in g["meta"]]
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    x = g["x"]
    h1 = orc.layer_norm(x, sd["0.norm1.weight"], sd["0.norm1.bias"], 1e-5)
    assert (h1 - g["b0_norm1"]).abs().max() < 1e-5
    a1 = orc.attention(h1, sd["0.attn.qkv.weight"], sd["0.attn.qkv.bias"], sd["0.attn.proj.weight"],
                       sd["0.attn.proj.bias"], heads)
    assert (a1 - g["b0_attn"]).abs().max() < 1e-5
    x1 = x + a1
    m1 = orc.mlp(orc.layer_norm(x1, sd["0.norm2.weight"], sd["0.norm2.bias"], 1e-5), sd["0.mlp.fc1.weight"],
                 sd["0.mlp.fc1.bias"], sd["0.mlp.fc2.weight"], sd["0.mlp.fc2.bias"])
    assert (m1 - g["b0_mlp"]).abs().max() < 1e-5


def test_image_patch_embed_matches_reference(golden_dir):
    g = _load(golden_dir, "image_patch_embed.npz")
    y = orc.patch_embed_2d(g["x"].float(), g["w"], g["b"], (16, 16))
    assert y.shape == g["y"].shape == (2, 196, 64)
    assert (y - g["y"]).abs().max() < 1e-4


def test_audio_patch_embed_matches_reference(golden_dir):
    """AST tokenizer: the reference's PatchEmbed.forward with the stride-10 conv ASTModel installs (ast_models.py:18-33,86)"""
    g = _load(golden_dir, "audio_patch_embed.npz")
    y = orc.patch_embed_2d(g["x"].float(), g["w"], g["b"], tuple(int(v) for v in g["stride"]))
    assert y.shape == g["y"].shape == (2, 45, 64)          # f = (64-16)//10+1 = 5, t = (100-16)//10+1 = 9
    assert (y - g["y"]).abs().max() < 1e-4


def test_video_patch_embed_matches_reference(golden_dir):
    """VideoMAE tubelet tokenizer (Video/models/modeling_finetune.py:263-297)"""
    g = _load(golden_dir, "video_patch_embed.npz")
    y = orc.patch_embed_3d(g["x"].float(), g["w"], g["b"], (2, 16, 16))
    assert y.shape == g["y"].shape == (2, 8, 64)
    assert (y - g["y"]).abs().max() < 1e-4


def test_time_series_embed_matches_reference(golden_dir):
    g = _load(golden_dir, "time_series_embed.npz")
    y = orc.time_series_embed(g["x"], g["w"])
    assert (y - g["y"]).abs().max() < 1e-5


def test_elementary_ops_match_torch_functional():
    """Independent cross-check of the written-out formulas against torch.nn.functional (fp64)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 7, 48, generator=g, dtype=torch.float64)
    w = torch.randn(48, generator=g, dtype=torch.float64)
    b = torch.randn(48, generator=g, dtype=torch.float64)
    assert torch.allclose(orc.layer_norm(x, w, b), F.layer_norm(x, (48,), w, b, 1e-5), atol=1e-12)
    assert torch.allclose(orc.gelu_erf(x), F.gelu(x), atol=1e-12)
    W = torch.randn(20, 48, generator=g, dtype=torch.float64)
    assert torch.allclose(orc.linear(x, W, b[:20]), F.linear(x, W, b[:20]), atol=1e-12)


def test_conv_tokenizers_match_torch_conv():
    """Audio / video tokenizers: no importable reference class (timm at file top) -> pinned against F.conv2d/3d."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(1)
    # AST: (B,T,F) -> (B,1,F,T) -> Conv2d(1,d,16,stride=10)  (Audio/src/models/ast_models.py:86,152-153)
    fb = torch.randn(2, 100, 64, generator=g)
    w = torch.randn(24, 1, 16, 16, generator=g) * 0.05
    b = torch.randn(24, generator=g)
    x = orc.audio_input_prep(fb)
    assert x.shape == (2, 1, 64, 100)
    ref = F.conv2d(x, w, b, stride=(10, 10)).flatten(2).transpose(1, 2)
    assert torch.allclose(orc.patch_embed_2d(x, w, b, (10, 10)), ref, atol=1e-4)
    # VideoMAE tubelets: Conv3d(3,d,(2,16,16),(2,16,16))  (Video/models/modeling_finetune.py:285-297)
    v = torch.randn(1, 3, 4, 32, 48, generator=g)
    w3 = torch.randn(16, 3, 2, 16, 16, generator=g) * 0.05
    b3 = torch.randn(16, generator=g)
    ref3 = F.conv3d(v, w3, b3, stride=(2, 16, 16)).flatten(2).transpose(1, 2)
    assert torch.allclose(orc.patch_embed_3d(v, w3, b3, (2, 16, 16)), ref3, atol=1e-4)


def test_oracle_vs_torchvision_encoderblock():
    """Second independent implementation of the same math."""
    from functools import partial
    tv = pytest.importorskip("torchvision.models.vision_transformer")
    dim, heads = 64, 2
    sd = orc.make_encoder_state_dict(dim, 1, seed=5)
    blk = tv.EncoderBlock(heads, dim, 4 * dim, 0.0, 0.0, norm_layer=partial(torch.nn.LayerNorm, eps=1e-5)).eval()
    with torch.no_grad():
        blk.ln_1.weight.copy_(sd["0.norm1.weight"]); blk.ln_1.bias.copy_(sd["0.norm1.bias"])
        blk.self_attention.in_proj_weight.copy_(sd["0.attn.qkv.weight"])
        blk.self_attention.in_proj_bias.copy_(sd["0.attn.qkv.bias"])
        blk.self_attention.out_proj.weight.copy_(sd["0.attn.proj.weight"])
        blk.self_attention.out_proj.bias.copy_(sd["0.attn.proj.bias"])
        blk.ln_2.weight.copy_(sd["0.norm2.weight"]); blk.ln_2.bias.copy_(sd["0.norm2.bias"])
        blk.mlp[0].weight.copy_(sd["0.mlp.fc1.weight"]); blk.mlp[0].bias.copy_(sd["0.mlp.fc1.bias"])
        blk.mlp[3].weight.copy_(sd["0.mlp.fc2.weight"]); blk.mlp[3].bias.copy_(sd["0.mlp.fc2.bias"])
        x = orc.make_tokens(2, 19, dim)
        assert torch.allclose(orc.encoder(x, sd, 1, heads), blk(x), atol=2e-5)


def test_graph_tokenizer_oracle_matches_reference_golden(golden_dir):
    """The oracle's restatement of Data2Seq/Graph.py:234-305 reproduces the reference's own output, given the two random
    node-identifier draws the generator script replayed (tests/golden/make_golden.py)."""
    import numpy as np
    z = np.load(golden_dir / "graph_tokenizer.npz")
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    y = orc.graph_tokens(sd, torch.from_numpy(z["node_data"]), torch.from_numpy(z["edge_data"]), torch.from_numpy(z["edge_index"]),
                         z["node_num"].tolist(), z["edge_num"].tolist(), torch.from_numpy(z["lap_eigvec"]),
                         torch.from_numpy(z["rand_node_id"]), torch.from_numpy(z["orf_node_id"]))
    ref = torch.from_numpy(z["y"])
    assert y.shape == ref.shape and (y - ref).abs().max().item() <= 2e-6
