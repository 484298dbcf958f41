"""GPU parity of the rows SURVEY.md §8f marks "next": positional-embedding / prefix-token fusion at the tokenizer,
final LayerNorm + pooling + head, training-time dropout / stochastic depth, text zero-padding.  Same bar as the
hot path: every call goes through the C ABI and is compared with an fp64 torch restatement on the same inputs."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import encoder_oracle as orc  # noqa: E402
from tests.test_gpu_parity import _assert_close, _bf16_round, mt  # noqa: E402,F401


def _rel(got, ref):
    return float((got.double().cpu() - ref.double().cpu()).norm() / ref.double().norm().clamp_min(1e-12))


# ---------------------------------------------------------------------------------------------------------------
# tokenizer side
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("M,N,K,period", [(300, 96, 40, 100), (1568, 768, 256, 196), (37, 32, 8, 37), (1024, 768, 768, 7)])
def test_gemm_residual_table_broadcast(mt, cg, M, N, K, period):
    g = torch.Generator().manual_seed(M + period)
    x = _bf16_round(torch.randn(M, K, generator=g))
    w = _bf16_round(torch.randn(N, K, generator=g) * 0.1)
    b = torch.randn(N, generator=g)
    table = _bf16_round(torch.randn(period, N, generator=g))
    ref = x.double() @ w.double().t() + b.double() + table.double()[torch.arange(M) % period]
    got = mt.ops.linear(x.cuda().bfloat16(), w.cuda().bfloat16(), b.cuda(), residual=table.cuda().bfloat16(),
                        residual_period=period, cta_group=cg)
    _assert_close(got, ref, what=f"gemm + table[m % {period}] cg={cg}")


@pytest.mark.parametrize("B,n,k,d,use_pos", [(3, 50, 2, 768, True), (2, 7, 1, 64, True), (4, 12, 0, 128, True), (2, 9, 3, 256, False)])
def test_assemble_tokens(mt, B, n, k, d, use_pos):
    g = torch.Generator().manual_seed(n * 7 + k)
    tok = _bf16_round(torch.randn(B, n, d, generator=g))
    pre = _bf16_round(torch.randn(k, d, generator=g)) if k else None
    pos = _bf16_round(torch.randn(n + k, d, generator=g)) if use_pos else None
    ref = tok if k == 0 else torch.cat([pre.unsqueeze(0).expand(B, -1, -1), tok], 1)
    if use_pos:
        ref = ref + pos
    got = mt.ops.assemble_tokens(tok.cuda().bfloat16(), None if pre is None else pre.cuda().bfloat16(),
                                 None if pos is None else pos.cuda().bfloat16())
    assert got.shape == (B, n + k, d)
    assert torch.equal(got.float().cpu(), ref.bfloat16().float())            # one rounding of an exact fp32 sum


def test_audio_tokenizer_with_cls_dist_and_pos_embed_trains(mt):
    """AST front end (Audio/src/models/ast_models.py:152-159): conv tokens, [cls, dist] prefix, + pos_embed, as one
    fused call; gradients of conv weight / bias, pos_embed and the prefix tokens vs torch autograd in fp64."""
    from metatransformer_b200.data2seq import Acoustic
    torch.manual_seed(4)
    d, B = 128, 3
    pe = Acoustic.PatchEmbed(embed_dim=d, fstride=16, tstride=16).cuda()
    with torch.no_grad():
        pe.proj.weight.copy_(_bf16_round(pe.proj.weight))
    x = _bf16_round(torch.randn(B, 1, 64, 96))                            # (B, 1, F, T) -> 4 x 6 = 24 patches
    n, k = 24, 2
    pos = nn.Parameter(_bf16_round(torch.randn(1, n + k, d) * 0.5).cuda())
    cls = nn.Parameter(_bf16_round(torch.randn(1, 1, d)).cuda())
    dist = nn.Parameter(_bf16_round(torch.randn(1, 1, d)).cuda())
    gy = _bf16_round(torch.randn(B, n + k, d) * 0.1)
    tok = pe(x.cuda(), pos_embed=pos, prefix_tokens=torch.cat([cls, dist], 1))
    assert tok.shape == (B, n + k, d)
    tok.backward(gy.cuda().to(tok.dtype))
    # fp64 restatement
    w = pe.proj.weight.detach().double().cpu().requires_grad_(True)
    bb = pe.proj.bias.detach().double().cpu().requires_grad_(True)
    posr = pos.detach().double().cpu().requires_grad_(True)
    clsr = cls.detach().double().cpu().requires_grad_(True)
    distr = dist.detach().double().cpu().requires_grad_(True)
    t = F.conv2d(x.double(), w, bb, stride=16).flatten(2).transpose(1, 2)
    ref = torch.cat([clsr.expand(B, -1, -1), distr.expand(B, -1, -1), t], 1) + posr
    ref.backward(gy.double())
    _assert_close(tok, ref.detach(), what="AST tokens")
    for name, got, want in [("conv.weight", pe.proj.weight.grad, w.grad), ("conv.bias", pe.proj.bias.grad, bb.grad),
                            ("pos_embed", pos.grad, posr.grad), ("cls", cls.grad, clsr.grad), ("dist", dist.grad, distr.grad)]:
        assert got is not None, name
        assert _rel(got, want) <= 2e-2, f"{name}: rel_l2 {_rel(got, want):.3e}"


def test_video_tokenizer_pos_embed_in_gemm_epilogue(mt):
    from metatransformer_b200.data2seq import Video
    torch.manual_seed(9)
    d = 64
    pe = Video.PatchEmbed(img_size=32, patch_size=16, embed_dim=d, num_frames=4, tubelet_size=2).cuda()
    x = _bf16_round(torch.randn(2, 3, 4, 32, 32))
    n = pe.num_patches
    pos = _bf16_round(torch.randn(1, n, d)).cuda()
    with torch.no_grad():
        got = pe(x.cuda(), pos_embed=pos)
    ref = F.conv3d(x.double(), _bf16_round(pe.proj.weight.detach().cpu()).double(), pe.proj.bias.detach().cpu().double(),
                   stride=(2, 16, 16)).flatten(2).transpose(1, 2) + pos.double().cpu()
    _assert_close(got, ref, what="video tokens + pos")


def test_text_zero_padding(mt):
    from metatransformer_b200.data2seq import Text
    f = torch.randn(5, 512)
    out = Text.zero_padding(f.cuda(), 768)
    assert out.shape == (5, 768) and torch.equal(out[:, :512].cpu(), f) and float(out[:, 512:].abs().max()) == 0.0
    tok = mt.Data2Seq("text", 768)(f.cuda().bfloat16())
    assert tok.shape == (5, 1, 768) and torch.equal(tok[:, 0, :512].float().cpu(), f.bfloat16().float())


# ---------------------------------------------------------------------------------------------------------------
# head side
# ---------------------------------------------------------------------------------------------------------------
def _pool_ref(x, gamma, beta, pool, eps=1e-5):
    if pool == "cls":
        return orc.layer_norm(x[:, 0], gamma, beta, eps)
    if pool == "mean":
        return orc.layer_norm(x.mean(1), gamma, beta, eps)
    y = orc.layer_norm(x[:, :2], gamma, beta, eps)
    return (y[:, 0] + y[:, 1]) / 2


@pytest.mark.parametrize("pool", ["cls", "mean", "cls_dist"])
@pytest.mark.parametrize("B,N,d", [(3, 197, 768), (2, 50, 1024), (5, 9, 200), (1, 1568, 1024)])
def test_pool_layernorm_forward_backward(mt, pool, B, N, d):
    g = torch.Generator().manual_seed(N + d)
    x = _bf16_round(torch.randn(B, N, d, generator=g) * 1.5 + 0.2)
    gamma = 1 + torch.randn(d, generator=g) * 0.2
    beta = torch.randn(d, generator=g) * 0.2
    gy = _bf16_round(torch.randn(B, d, generator=g))
    mode = {"cls": mt.ops.POOL_FIRST, "mean": mt.ops.POOL_MEAN, "cls_dist": mt.ops.POOL_FIRST_TWO}[pool]
    xr, gr, br = x.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = _pool_ref(xr, gr, br, pool)
    ref.backward(gy.double())
    got = mt.ops.pool_layernorm(x.cuda().bfloat16(), gamma.cuda(), beta.cuda(), mode)
    _assert_close(got, ref.detach(), what=f"pool_layernorm {pool}")
    dx, dg, db = mt.ops.pool_layernorm_bwd(gy.cuda().bfloat16(), x.cuda().bfloat16(), gamma.cuda(), mode)
    # mean pooling spreads the gradient over N tokens: compare in relative L2 (values are O(1/N))
    assert _rel(dx, xr.grad) <= 1e-2, f"dx rel_l2 {_rel(dx, xr.grad):.3e}"
    assert _rel(dg, gr.grad) <= 1e-2 and _rel(db, br.grad) <= 1e-2


@pytest.mark.parametrize("pool,classes", [("cls", 10), ("mean", 400), ("cls_dist", 35)])
def test_classifier_head_trains(mt, pool, classes):
    torch.manual_seed(classes)
    B, N, d = 6, 20, 256
    head = mt.ClassifierHead(d, classes, pool=pool).cuda()
    with torch.no_grad():
        head.norm.weight.add_(torch.randn(d, device="cuda") * 0.1)
        head.norm.bias.add_(torch.randn(d, device="cuda") * 0.1)
        head.head.weight.copy_(_bf16_round(head.head.weight))
    x = _bf16_round(torch.randn(B, N, d))
    gy = _bf16_round(torch.randn(B, classes) * 0.1)
    xg = x.cuda().requires_grad_(True)                                     # fp32 tokens in, fp32 logits out
    out = head(xg)
    assert out.shape == (B, classes) and out.dtype == torch.float32
    out.backward(gy.cuda())
    ref_p = {k: v.detach().double().cpu().requires_grad_(True) for k, v in head.named_parameters()}
    xr = x.double().requires_grad_(True)
    ref = _pool_ref(xr, ref_p["norm.weight"], ref_p["norm.bias"], pool) @ ref_p["head.weight"].t() + ref_p["head.bias"]
    ref.backward(gy.double())
    _assert_close(out, ref.detach(), what=f"head logits ({pool})")
    assert _rel(xg.grad, xr.grad) <= 2e-2, f"dx {_rel(xg.grad, xr.grad):.3e}"
    for name, prm in head.named_parameters():
        assert prm.grad is not None, name
        assert _rel(prm.grad, ref_p[name].grad) <= 2e-2, f"{name}: {_rel(prm.grad, ref_p[name].grad):.3e}"
    with torch.no_grad():
        feat = mt.pool_norm(x.cuda(), head.norm, pool)
    _assert_close(feat, _pool_ref(x.double(), ref_p["norm.weight"].detach(), ref_p["norm.bias"].detach(), pool), what="pool_norm")


# ---------------------------------------------------------------------------------------------------------------
# dropout / stochastic depth
# ---------------------------------------------------------------------------------------------------------------
def test_dropout_kernel_statistics_and_replay(mt):
    rows, d, p = 4096, 768, 0.1
    ones = torch.ones(rows, d, device="cuda", dtype=torch.bfloat16)
    a = mt.ops.dropout(ones, p, seed=1234, offset=0)
    b = mt.ops.dropout(ones, p, seed=1234, offset=0)
    c = mt.ops.dropout(ones, p, seed=1234, offset=rows * d // 8)
    e = mt.ops.dropout(ones, p, seed=99, offset=0)
    assert torch.equal(a, b), "same (seed, offset) must replay the same mask"
    assert not torch.equal(a, c) and not torch.equal(a, e)
    keep = (a != 0).float().mean().item()
    sigma = (p * (1 - p) / (rows * d)) ** 0.5
    assert abs(keep - (1 - p)) < 6 * sigma + 2e-5, f"keep fraction {keep}"
    vals = a[a != 0].float()
    assert torch.allclose(vals, torch.full_like(vals, 1 / (1 - p)), rtol=1e-2)        # bf16(1/0.9)
    assert abs(a.float().mean().item() - 1.0) < 5e-3                                  # E[out] = x
    # masks of neighbouring elements / rows are uncorrelated (lane extraction from one Philox word)
    m = (a != 0).float()
    assert abs(((m[:, :-1] - m.mean()) * (m[:, 1:] - m.mean())).mean().item()) < 5e-4
    # p = 0 with residual and per-sample scale = drop-path + residual add
    x = torch.randn(6 * 10, 64, device="cuda").bfloat16()
    r = torch.randn(6 * 10, 64, device="cuda").bfloat16()
    rs = torch.tensor([0.0, 1.25, 1.25, 0.0, 1.25, 1.25], device="cuda")
    out = mt.ops.dropout(x, 0.0, 1, 0, residual=r, row_scale=rs, rows_per_scale=10)
    ref = r.float() + rs.repeat_interleave(10)[:, None] * x.float()
    assert torch.equal(out.float(), ref.bfloat16().float())


def _masks_from_saved(mt, drop, M, d, hidden, N):
    """Replay the three dropout sites of a Block on ones -> multiplicative masks (incl. 1/(1-p) and drop-path scale)."""
    one_d = torch.ones(M, d, device="cuda", dtype=torch.bfloat16)
    one_h = torch.ones(M, hidden, device="cuda", dtype=torch.bfloat16)
    m1 = mt.ops.dropout(one_d, drop["p_proj"], *drop["k1"], row_scale=drop["rs1"], rows_per_scale=N)
    mh = mt.ops.dropout(one_h, drop["p_mlp"], *drop["kh"]) if drop["p_mlp"] > 0 else one_h
    m2 = mt.ops.dropout(one_d, drop["p_mlp"], *drop["k2"], row_scale=drop["rs2"], rows_per_scale=N)
    return m1.double().cpu(), mh.double().cpu(), m2.double().cpu()


@pytest.mark.parametrize("drop,drop_path", [(0.1, 0.0), (0.0, 0.3), (0.1, 0.2)])
def test_block_training_with_dropout_and_drop_path(mt, drop, drop_path):
    """Block(drop, drop_path) in train mode: forward, input gradient and all parameter gradients vs an fp64 restatement of
    attention.py:55-58 / mlp.py:30-35 that uses the very masks the kernels drew (replayed from the saved Philox keys)."""
    dim, heads, B, N = 256, 4, 4, 36
    torch.manual_seed(11)
    torch.cuda.manual_seed(11)
    blk = mt.Block(dim, heads, qkv_bias=True, drop=drop, drop_path=drop_path).cuda().train()
    with torch.no_grad():
        for name, prm in blk.named_parameters():
            if prm.dim() == 2:
                prm.copy_(_bf16_round(prm))
            else:
                prm.add_(torch.randn_like(prm) * 0.05)
    x = _bf16_round(torch.randn(B, N, dim))
    gy = _bf16_round(torch.randn(B, N, dim) * 0.1)
    x2 = x.cuda().bfloat16().view(B * N, dim)
    out2, saved = blk._forward_train(x2, B, N)
    cfg = saved[-1]
    assert cfg is not None
    if drop_path > 0:
        assert all(v == 0.0 or abs(v - 1 / (1 - drop_path)) < 1e-6 for v in cfg["rs1"].tolist() + cfg["rs2"].tolist())
    m1, mh, m2 = _masks_from_saved(mt, cfg, B * N, dim, 4 * dim, N)
    p = {k: v.detach().double().cpu().requires_grad_(True) for k, v in blk.named_parameters()}
    xr = x.double().view(B * N, dim).requires_grad_(True)
    h = orc.layer_norm(xr, p["norm1.weight"], p["norm1.bias"], 1e-5)
    qkv = (h @ p["attn.qkv.weight"].t() + p["attn.qkv.bias"]).view(B, N, 3, heads, dim // heads).permute(2, 0, 3, 1, 4)
    a = ((qkv[0] @ qkv[1].transpose(-2, -1)) * (dim // heads) ** -0.5).softmax(-1) @ qkv[2]
    o = a.transpose(1, 2).reshape(B * N, dim)
    xm = xr + m1 * (o @ p["attn.proj.weight"].t() + p["attn.proj.bias"])
    h2 = orc.layer_norm(xm, p["norm2.weight"], p["norm2.bias"], 1e-5)
    hid = orc.gelu_erf(h2 @ p["mlp.fc1.weight"].t() + p["mlp.fc1.bias"]) * mh
    ref = xm + m2 * (hid @ p["mlp.fc2.weight"].t() + p["mlp.fc2.bias"])
    ref.backward(gy.double().view(B * N, dim))
    _assert_close(out2, ref.detach(), what="train-mode Block forward with masks")
    dx, inter = blk._backward_2d(gy.cuda().bfloat16().view(B * N, dim), saved, B, N)
    assert _rel(dx, xr.grad) <= 2e-2, f"dx {_rel(dx, xr.grad):.3e}"
    grads = blk._weight_grads(saved, inter)
    for name, got in zip(blk._PARAM_ORDER, grads):
        want = p[name].grad
        assert _rel(got.reshape(want.shape), want) <= 3e-2, f"{name}: {_rel(got.reshape(want.shape), want):.3e}"


def test_dropout_follows_torch_cuda_rng_and_checkpointing(mt):
    """Masks come from torch's CUDA generator: reseeding replays them, eval mode ignores them, and
    torch.utils.checkpoint (Video/models/modeling_finetune.py:440-441) recomputes the same forward."""
    from torch.utils.checkpoint import checkpoint
    dim, heads, B, N = 128, 2, 2, 40
    torch.manual_seed(0)
    blk = mt.Block(dim, heads, qkv_bias=True, drop=0.2, drop_path=0.1).cuda().train()
    x = torch.randn(B, N, dim, device="cuda").bfloat16()

    def run(seed, use_cp=False):
        torch.cuda.manual_seed(seed)
        xg = x.clone().requires_grad_(True)
        for prm in blk.parameters():
            prm.grad = None
        y = checkpoint(blk, xg, use_reentrant=False) if use_cp else blk(xg)
        y.float().square().mean().backward()
        return y.detach(), xg.grad.detach(), blk.mlp.fc2.weight.grad.detach().clone()

    y1, g1, w1 = run(5)
    y2, g2, w2 = run(5)
    y3, _, _ = run(6)
    assert torch.equal(y1, y2) and torch.allclose(g1.float(), g2.float(), rtol=1e-2, atol=1e-6)
    assert not torch.equal(y1, y3)
    yc, gc, wc = run(5, use_cp=True)
    assert torch.equal(y1, yc) and torch.allclose(g1.float(), gc.float(), rtol=1e-2, atol=1e-6)
    assert torch.allclose(w1, wc, rtol=1e-2, atol=1e-5)
    with torch.no_grad():
        torch.cuda.manual_seed(5)
        yt = blk(x)                        # train mode, no autograd: masks still apply and follow the generator
        assert torch.equal(yt, y1)
        blk.eval()
        ye1, ye2 = blk(x), blk(x)
        assert torch.equal(ye1, ye2) and not torch.equal(ye1, y1)


@pytest.mark.parametrize("B,N,H,p", [(2, 196, 3, 0.1), (1, 300, 2, 0.25), (2, 64, 1, 0.5)])
def test_attention_dropout_forward_backward_match_oracle_with_the_same_mask(mt, B, N, H, p):
    """attn = softmax(q k^T scale); attn = keep * attn / (1 - p); out = attn @ v (attention.py:31-35 with attn_drop > 0).
    The kernels generate keep(b, h, q, k) from a hash; mt_attention_dropout_mask dumps the same mask for the fp64 oracle."""
    g = torch.Generator().manual_seed(N + H)
    C = H * 64
    seed = 0x1234_5678_9ABC_DEF0 + N
    qkv = _bf16_round(torch.randn(B, N, 3 * C, generator=g))
    do = _bf16_round(torch.randn(B, N, C, generator=g))
    keep = mt.ops.attention_dropout_mask(B, N, H, p, seed, "cuda").cpu().double()
    frac = keep.mean().item()
    assert abs(frac - (1 - p)) < 0.01, f"keep fraction {frac:.4f} vs {1 - p}"
    assert abs(keep[:, :, :, ::2].mean().item() - (1 - p)) < 0.015 and abs(keep[:, :, ::2].mean().item() - (1 - p)) < 0.015
    qr = qkv.double().requires_grad_(True)
    q, k, v = qr.reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
    a = ((q @ k.transpose(-2, -1)) * 0.125).softmax(-1) * keep / (1 - p)
    ref = (a @ v).transpose(1, 2).reshape(B, N, C)
    ref.backward(do.double())
    qd = qkv.cuda().bfloat16()
    out, lse2 = mt.ops.attention_lse(qd, H, drop_p=p, seed=seed)
    _assert_close(out, ref.detach(), what=f"attention + dropout({p}) forward")
    dqkv = mt.ops.attention_bwd(qd, out, do.cuda().bfloat16(), lse2, H, drop_p=p, seed=seed)
    _assert_close(dqkv, qr.grad, rel=2e-2, mx=2e-2, what=f"attention + dropout({p}) backward")
    out2, _ = mt.ops.attention_lse(qd, H, drop_p=p, seed=seed + 1)
    assert not torch.equal(out2, out)


def test_graph_fork_block_trains_with_attn_drop(mt):
    """Graph fork Block (dim 768, 32 heads, attn_drop 0.1; tokengt_graph_encoder.py:193-203) in training mode: runs through
    the tcgen05 kernels, masks follow torch's CUDA generator, eval mode is deterministic and mask-free."""
    torch.manual_seed(0)
    blk = mt.Block(dim=768, num_heads=32, mlp_ratio=4., qkv_bias=True, drop=0.0, attn_drop=0.1).cuda().train()
    x = torch.randn(4, 50, 768, device="cuda").requires_grad_(True)
    torch.cuda.manual_seed(11)
    y1 = blk(x)
    y1.float().square().mean().backward()
    assert torch.isfinite(x.grad).all() and all(torch.isfinite(q.grad).all() for q in blk.parameters())
    torch.cuda.manual_seed(11)
    y2 = blk(x)
    torch.cuda.manual_seed(12)
    y3 = blk(x)
    assert torch.equal(y1, y2) and not torch.equal(y1, y3)
    blk.eval()
    with torch.no_grad():
        e1, e2 = blk(x), blk(x)
    assert torch.equal(e1, e2)
    rel = float((y1.detach() - e1).norm() / e1.norm())
    assert 1e-4 < rel < 0.2, f"train-vs-eval relative difference {rel:.3e}"


# ---------------------------------------------------------------------------------------------------------------
# round-2 additions: calendar embeddings / training of the time-series tokenizer, stand-alone DropPath, Block(x, pos)
# ---------------------------------------------------------------------------------------------------------------
def test_time_series_calendar_embeddings_match_reference_golden(mt, golden_dir):
    """DataEmbedding(x, x_mark) for the fixed sinusoid tables (freq 't': month/day/weekday/hour/minute) and the timeF
    Linear, against vectors produced by the reference's own Data2Seq/Time_Series.py (tests/golden/make_golden.py)."""
    import numpy as np
    z = np.load(golden_dir / "time_series_embed_mark.npz")
    x = torch.from_numpy(z["x"]).cuda()
    emb = mt.data2seq.Time_Series.DataEmbedding(c_in=7, d_model=64, embed_type="fixed", freq="t")
    emb.value_embedding.tokenConv.weight.data.copy_(torch.from_numpy(z["w_fixed"]))
    emb = emb.cuda().eval()
    with torch.no_grad():
        y = emb(x, torch.from_numpy(z["x_mark_fixed"]).cuda())
    _assert_close(y, torch.from_numpy(z["y_fixed"]), what="DataEmbedding + fixed calendar tables")
    emb = mt.data2seq.Time_Series.DataEmbedding(c_in=7, d_model=64, embed_type="timeF", freq="h")
    emb.value_embedding.tokenConv.weight.data.copy_(torch.from_numpy(z["w_timef"]))
    emb.temporal_embedding.embed.weight.data.copy_(torch.from_numpy(z["w_timef_embed"]))
    emb = emb.cuda().eval()
    with torch.no_grad():
        y = emb(x, torch.from_numpy(z["x_mark_timef"]).cuda())
    _assert_close(y, torch.from_numpy(z["y_timef"]), what="DataEmbedding + timeF Linear")
    # a reference state-dict loads strict (same sub-module names)
    assert set(emb.state_dict().keys()) == {"value_embedding.tokenConv.weight", "position_embedding.pe",
                                            "temporal_embedding.embed.weight"}


def test_time_series_tokenizer_trains(mt):
    """default construction (train mode, dropout 0.1) works out of the box; weight gradients vs fp64 autograd at p = 0"""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 48, 5, generator=g)
    xt = torch.rand(2, 48, 4, generator=g) - 0.5
    emb = mt.data2seq.Time_Series.DataEmbedding(c_in=5, d_model=96, embed_type="timeF", freq="h", dropout=0.0).cuda().train()
    y = emb(x.cuda(), xt.cuda())
    gy = torch.randn(2, 48, 96, generator=g)
    y.backward(gy.cuda())
    w = emb.value_embedding.tokenConv.weight.detach().cpu().double().requires_grad_(True)
    wt = emb.temporal_embedding.embed.weight.detach().cpu().double().requires_grad_(True)
    xp = x.double().permute(0, 2, 1)
    ref = F.conv1d(F.pad(xp, (1, 1), mode="circular"), w).transpose(1, 2) + xt.double() @ wt.t() + \
        orc.sinusoid_pe(48, 96).double()
    ref.backward(gy.double())
    _assert_close(y, ref.detach(), what="time-series tokens (train mode)")
    _assert_close(emb.value_embedding.tokenConv.weight.grad, w.grad, rel=2e-2, mx=2e-2, what="dW conv1d")
    _assert_close(emb.temporal_embedding.embed.weight.grad, wt.grad, rel=2e-2, mx=2e-2, what="dW timeF")
    d2s = mt.Data2Seq("time-series", 64).cuda()            # default: train mode, dropout 0.1, trainable conv
    out = d2s(torch.randn(2, 32, 1, device="cuda"))
    out.float().sum().backward()
    assert out.shape == (2, 32, 64) and d2s.embed.value_embedding.tokenConv.weight.grad is not None
    zeros = (out == 0).float().mean().item()
    assert 0.03 < zeros < 0.2, f"dropout(0.1) zeroed {zeros:.3f} of the tokens"


def test_standalone_droppath_and_block_pos_prologue(mt):
    g = torch.Generator().manual_seed(5)
    dp = mt.vision_transformer.DropPath(0.4).cuda().train()
    x = torch.randn(64, 10, 32, generator=g).cuda().requires_grad_(True)
    y = dp(x)
    scale = (y.detach() / x.detach()).mean(dim=(1, 2))
    kept = (scale - 1 / 0.6).abs() < 2e-2
    dropped = scale.abs() < 1e-6
    assert bool((kept | dropped).all()) and 0.3 < kept.float().mean().item() < 0.9
    y.sum().backward()
    assert torch.allclose(x.grad[:, 0, 0], torch.where(kept, torch.tensor(1 / 0.6, device="cuda"), torch.tensor(0.0, device="cuda")),
                          atol=2e-2)
    assert dp.eval()(x) is x
    # Block(x, pos) == Block(x + pos) (PointCloud fork's per-layer positional add)
    sd = orc.make_encoder_state_dict(128, 1, seed=7)
    blk = mt.Block(dim=128, num_heads=2, mlp_ratio=4., qkv_bias=True).cuda().eval()
    blk.load_state_dict({k[2:]: v for k, v in sd.items()}, strict=True)
    xb = _bf16_round(torch.randn(3, 40, 128, generator=g))
    pos = _bf16_round(torch.randn(1, 40, 128, generator=g) * 0.5)
    ref = orc.block((xb + pos).double(), {k: v.double() for k, v in sd.items()}, "0.", 2)
    with torch.no_grad():
        got = blk(xb.cuda(), pos=pos.cuda())
    _assert_close(got, ref, what="Block(x, pos)")
    posp = pos.clone().cuda().requires_grad_(True)
    out = blk(xb.cuda().requires_grad_(True), pos=posp)
    out.float().sum().backward()
    assert posp.grad is not None and posp.grad.shape == posp.shape and torch.isfinite(posp.grad).all()


def test_graph_tokenizer_matches_reference_golden(mt, golden_dir):
    """Data2Seq/Graph.py GraphFeatureTokenizer: three ragged graphs, every identifier encoder on; vectors produced by the
    reference's own class on CPU, its two random node-identifier draws replayed and injected (make_golden.py)."""
    import numpy as np
    z = np.load(golden_dir / "graph_tokenizer.npz")
    tok = mt.data2seq.Graph.GraphFeatureTokenizer(num_atoms=40, num_edges=12, rand_node_id=True, rand_node_id_dim=16,
                                                  orf_node_id=True, orf_node_id_dim=16, lap_node_id=True, lap_node_id_k=8,
                                                  lap_node_id_sign_flip=False, lap_node_id_eig_dropout=0.0, type_id=True,
                                                  hidden_dim=64, n_layers=2)
    tok.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}, strict=True)
    tok = tok.cuda().eval()
    batch = dict(node_data=torch.from_numpy(z["node_data"]).cuda(), edge_data=torch.from_numpy(z["edge_data"]).cuda(),
                 edge_index=torch.from_numpy(z["edge_index"]).cuda(), lap_eigvec=torch.from_numpy(z["lap_eigvec"]).cuda(),
                 node_num=z["node_num"].tolist(), edge_num=z["edge_num"].tolist())
    with torch.no_grad():
        feat, pmask, pidx = tok(batch, rand_node_id=torch.from_numpy(z["rand_node_id"]).cuda(),
                                orf_node_id=torch.from_numpy(z["orf_node_id"]).cuda())
    assert torch.equal(pmask.cpu(), torch.from_numpy(z["padding_mask"])) and torch.equal(pidx.cpu(), torch.from_numpy(z["padded_index"]))
    _assert_close(feat, torch.from_numpy(z["y"]), what="GraphFeatureTokenizer golden")
    assert float(feat[pmask].abs().max()) == 0.0                     # padding tokens are exactly zero
    # the dispatcher + the Graph fork's Block geometry (768 wide, 32 heads) on the tokens
    d2s = mt.Data2Seq("graph", 768).cuda().eval()
    g = torch.Generator().manual_seed(3)
    nn_, en_ = [6, 4], [9, 5]
    gb = dict(node_data=torch.randint(1, 4000, (10, 9), generator=g).cuda(), edge_data=torch.randint(1, 1000, (14, 3), generator=g).cuda(),
              edge_index=torch.cat([torch.randint(0, n, (2, e), generator=g) for n, e in zip(nn_, en_)], dim=1).cuda(),
              node_num=nn_, edge_num=en_)
    with torch.no_grad():
        t = d2s(gb)
        y = mt.Block(dim=768, num_heads=32, qkv_bias=True).cuda().eval()(t)
    assert t.shape == (2, 2 + 15, 768) and y.shape == t.shape and torch.isfinite(y).all()


@pytest.mark.parametrize("M,N,K,gelu,resid", [(300, 768, 3072, False, True), (197, 3072, 768, True, False), (64, 96, 48, False, False)])
def test_fp8_e4m3_linear_opt_in(mt, M, N, K, gelu, resid):
    """mt_quantize_fp8 + mt_gemm_fp8 (tcgen05.mma.kind::f8f6f4): the quantiser agrees with torch's e4m3 cast, and the GEMM
    is exact on the quantised operands (products of e4m3 values accumulate in fp32) up to the bf16 output rounding."""
    g = torch.Generator().manual_seed(M + K)
    x = _bf16_round(torch.randn(M, K, generator=g) * 2)
    w = _bf16_round(torch.randn(N, K, generator=g) * 0.05)
    b = torch.randn(N, generator=g) * 0.3
    r = _bf16_round(torch.randn(M, N, generator=g)) if resid else None
    sx, sw = float(x.abs().max()) / 448.0, float(w.abs().max()) / 448.0
    xq = mt.ops.quantize_fp8(x.cuda().bfloat16(), sx)
    wq = mt.ops.quantize_fp8(w.cuda().bfloat16(), sw)
    xq_ref = (x / sx).to(torch.float8_e4m3fn)
    assert torch.equal(xq.cpu().view(torch.float8_e4m3fn).float(), xq_ref.float())
    xd, wd = xq.cpu().view(torch.float8_e4m3fn).double() * sx, wq.cpu().view(torch.float8_e4m3fn).double() * sw
    ref = xd @ wd.t() + b.double()
    if gelu:
        ref = orc.gelu_erf(ref)
    if resid:
        ref = ref + r.double()
    y = mt.ops.linear_fp8(xq, wq, sx * sw, bias=b.cuda(), residual=None if r is None else r.cuda().bfloat16(), gelu=gelu)
    err = (y.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2 ** -8, f"fp8 GEMM on quantised operands: {err:.2e}"
    # against the unquantised product the error is the e4m3 quantisation error (3 mantissa bits per operand)
    full = x.double() @ w.double().t() + b.double()
    if not gelu and not resid:
        assert float((y.double().cpu() - full).norm() / full.norm()) < 0.08
