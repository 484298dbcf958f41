"""GPU parity of the backward kernels against torch autograd of the fp64 CPU oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import encoder_oracle as orc  # noqa: E402
from tests.test_gpu_parity import _assert_close, _bf16_round, mt  # noqa: E402,F401


@pytest.mark.parametrize("B,N,H", [(1, 128, 1), (2, 100, 2), (2, 196, 3), (1, 256, 2), (1, 300, 1), (1, 520, 2)])
def test_attention_backward_matches_autograd(mt, B, N, H):
    g = torch.Generator().manual_seed(N * 3 + H)
    C = H * 64
    qkv = _bf16_round(torch.randn(B, N, 3 * C, generator=g))
    dout = _bf16_round(torch.randn(B, N, C, generator=g))
    x = qkv.double().requires_grad_(True)
    q, k, v = x.reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
    ref = (((q @ k.transpose(-2, -1)) * 0.125).softmax(-1) @ v).transpose(1, 2).reshape(B, N, C)
    ref.backward(dout.double())
    out, lse2 = mt.ops.attention_lse(qkv.cuda().bfloat16(), H)
    _assert_close(out, ref.detach(), what="attention fwd (lse variant)")
    s = (q @ k.transpose(-2, -1)) * 0.125
    lse_ref = torch.logsumexp(s.detach(), dim=-1) * 1.4426950408889634          # [B,H,N], base 2
    assert (lse2.double().cpu() - lse_ref).abs().max() < 2e-2
    dqkv = mt.ops.attention_bwd(qkv.cuda().bfloat16(), out, dout.cuda().bfloat16(), lse2, H)
    _assert_close(dqkv, x.grad, rel=2e-2, mx=2e-2, what=f"attention bwd B={B} N={N} H={H}")


@pytest.mark.parametrize("M,d,use_gamma,use_add", [(37, 768, False, True), (200, 1024, True, False), (5, 200, True, True)])
def test_layernorm_backward_matches_autograd(mt, M, d, use_gamma, use_add):
    g = torch.Generator().manual_seed(d + M)
    x = _bf16_round(torch.randn(M, d, generator=g) * 2 + 0.3)
    dy = _bf16_round(torch.randn(M, d, generator=g))
    gamma = 1 + torch.randn(d, generator=g) * 0.2
    add = _bf16_round(torch.randn(M, d, generator=g))
    xx = x.double().requires_grad_(True)
    w = gamma.double() if use_gamma else torch.ones(d, dtype=torch.float64)
    orc.layer_norm(xx, w, torch.zeros(d, dtype=torch.float64), 1e-5).backward(dy.double())
    ref = xx.grad + (add.double() if use_add else 0)
    stats = mt.ops.row_stats(x.cuda().bfloat16(), 1e-5)
    dx = mt.ops.layernorm_bwd(dy.cuda().bfloat16(), x.cuda().bfloat16(), stats, gamma.cuda() if use_gamma else None,
                              add.cuda().bfloat16() if use_add else None)
    _assert_close(dx, ref, what="layernorm bwd")


def test_gelu_and_gelu_grad_gemm(mt):
    g = torch.Generator().manual_seed(5)
    M, d, hidden = 300, 256, 1024
    u = _bf16_round(torch.randn(M, hidden, generator=g) * 1.5)
    h = mt.ops.gelu(u.cuda().bfloat16())
    _assert_close(h, orc.gelu_erf(u.double()), what="gelu")
    dy = _bf16_round(torch.randn(M, d, generator=g))
    w2 = _bf16_round(torch.randn(d, hidden, generator=g) * 0.05)          # fc2.weight [d, hidden]
    uu = u.double().requires_grad_(True)
    (orc.gelu_erf(uu) @ w2.double().t()).backward(dy.double())
    got = mt.ops.linear_gelu_grad(dy.cuda().bfloat16(), w2.t().contiguous().cuda().bfloat16(), u.cuda().bfloat16())
    _assert_close(got, uu.grad, what="fc2 dgrad * gelu'")


def test_block_input_gradient_matches_oracle_autograd(mt):
    """Frozen Base-16 Blocks, trainable input (the tokenizer's output): dL/dx through 2 Blocks vs fp64 autograd."""
    depth = 2
    sd = orc.make_encoder_state_dict(768, depth, seed=3)
    sd = {k: (_bf16_round(v) if v.dim() == 2 else v) for k, v in sd.items()}
    enc = mt.build_encoder("base", depth=depth)
    enc.load_state_dict(sd, strict=True)
    enc = enc.cuda().eval().requires_grad_(False)
    g = torch.Generator().manual_seed(21)
    x = _bf16_round(torch.randn(2, 196, 768, generator=g))
    gy = _bf16_round(torch.randn(2, 196, 768, generator=g))
    xr = x.double().requires_grad_(True)
    yr = orc.encoder(xr, {k: v.double() for k, v in sd.items()}, depth, 12)
    yr.backward(gy.double())
    xg = x.cuda().bfloat16().requires_grad_(True)
    y = enc(xg)
    _assert_close(y, yr.detach(), what="training-mode forward")
    y.backward(gy.cuda().bfloat16())
    assert xg.grad is not None and xg.grad.dtype == torch.bfloat16
    _assert_close(xg.grad, xr.grad, rel=2e-2, mx=2e-2, what="dL/dx through 2 frozen Blocks")


def test_finetune_weight_gradients_match_oracle_autograd(mt):
    """Fine-tune mode: gradients of all 12 parameters of a Block vs fp64 autograd of the oracle."""
    dim, heads = 768, 12
    sd = orc.make_encoder_state_dict(dim, 1, seed=8)
    sd = {k: (_bf16_round(v) if v.dim() == 2 else v) for k, v in sd.items()}
    enc = mt.build_encoder("base", depth=1)
    enc.load_state_dict(sd, strict=True)
    enc = enc.cuda().train()
    g = torch.Generator().manual_seed(77)
    x = _bf16_round(torch.randn(4, 196, dim, generator=g))
    gy = _bf16_round(torch.randn(4, 196, dim, generator=g) * 0.05)
    ref_p = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    orc.encoder(x.double(), ref_p, 1, heads).backward(gy.double())
    y = enc(x.cuda().bfloat16())             # input does not require grad; parameters do
    y.backward(gy.cuda().bfloat16())
    for name, prm in enc.named_parameters():
        assert prm.grad is not None, name
        ref = ref_p[name].grad
        tol = 3e-2
        got = prm.grad.reshape(ref.shape)
        rel = float((got.double().cpu() - ref).norm() / ref.norm().clamp_min(1e-12))
        assert rel <= tol, f"{name}: rel_l2 {rel:.3e}"


def test_transpose_and_colsum(mt):
    g = torch.Generator().manual_seed(9)
    x = _bf16_round(torch.randn(301, 96, generator=g))
    xt = mt.ops.transpose(x.cuda().bfloat16())
    assert xt.shape == (96, 304)
    assert torch.equal(xt[:, :301].float().cpu(), x.t()) and float(xt[:, 301:].abs().max()) == 0.0
    cs = mt.ops.colsum(x.cuda().bfloat16())
    assert (cs.cpu() - x.sum(0)).abs().max() < 1e-3


@pytest.mark.parametrize("M,N", [(50176, 3072), (50176, 768), (1000, 2304), (37, 8), (5, 264), (4096, 100), (3, 250000)])
def test_colsum_bias_gradient(mt, M, N):
    """db = column sums of dY in fp32 (vectorised kernel for N % 8 == 0, scalar kernel otherwise)"""
    g = torch.Generator().manual_seed(M + N)
    x = _bf16_round(torch.randn(M, N, generator=g))
    cs = mt.ops.colsum(x.cuda().bfloat16())
    ref = x.double().sum(0)
    assert cs.dtype == torch.float32 and cs.shape == (N,)
    assert (cs.double().cpu() - ref).abs().max() <= 1e-5 * M ** 0.5 * 8 + 1e-4


def test_trainable_tokenizer_under_frozen_encoder(mt):
    """The AST training pattern (Audio/src/models/ast_models.py:66-70): encoder frozen, patch-embed conv trainable.
    Gradients of the conv weight / bias through 2 frozen Blocks vs fp64 autograd of the oracle."""
    depth = 2
    sd = orc.make_encoder_state_dict(768, depth, seed=6)
    sd = {k: (_bf16_round(v) if v.dim() == 2 else v) for k, v in sd.items()}
    enc = mt.build_encoder("base", depth=depth)
    enc.load_state_dict(sd, strict=True)
    enc = enc.cuda().eval().requires_grad_(False)
    tok = mt.Data2Seq("image", 768).cuda()
    with torch.no_grad():
        tok.embed.proj.weight.copy_(_bf16_round(tok.embed.proj.weight.cpu()).cuda())
    g = torch.Generator().manual_seed(31)
    img = _bf16_round(torch.rand(2, 3, 224, 224, generator=g))
    gy = _bf16_round(torch.randn(2, 196, 768, generator=g) * 0.1)
    # oracle
    w = tok.embed.proj.weight.detach().cpu().double().requires_grad_(True)
    b = tok.embed.proj.bias.detach().cpu().double().requires_grad_(True)
    t_ref = orc.patch_embed_2d(img.double(), w, b, (16, 16))
    y_ref = orc.encoder(t_ref, {k: v.double() for k, v in sd.items()}, depth, 12)
    y_ref.backward(gy.double())
    # kernels
    y = enc(tok(img.cuda()))
    y.backward(gy.cuda())
    gw, gb = tok.embed.proj.weight.grad, tok.embed.proj.bias.grad
    assert gw is not None and gw.shape == tok.embed.proj.weight.shape and gb is not None
    _assert_close(gw.reshape(768, -1), w.grad.reshape(768, -1), rel=3e-2, mx=3e-2, what="conv weight grad")
    _assert_close(gb, b.grad, rel=3e-2, mx=3e-2, what="conv bias grad")
    assert all(p.grad is None for p in enc.parameters())


@pytest.mark.parametrize("via_transpose", [False, True])
@pytest.mark.parametrize("M,N,K,splits", [(1000, 96, 200, 0), (50176, 768, 768, 0), (4096, 3072, 768, 7), (333, 64, 64, 1),
                                          (197, 8, 3072, 0), (2, 520, 72, 3)])
def test_splitk_weight_gradient_gemm(mt, M, N, K, splits, via_transpose):
    """dW = dY^T X in fp32 by the split-K GEMM (fp32 atomics): MN-major operands straight from the row-major
    activations (default) and the older form on transposed copies."""
    g = torch.Generator().manual_seed(M + N)
    dy = _bf16_round(torch.randn(M, N, generator=g) * 0.1)
    x = _bf16_round(torch.randn(M, K, generator=g))
    ref = dy.double().t() @ x.double()
    dw = mt.ops.linear_wgrad(dy.cuda().bfloat16(), x.cuda().bfloat16(), splits, via_transpose=via_transpose)
    assert dw.dtype == torch.float32 and dw.shape == (N, K)
    _assert_close(dw, ref, rel=1e-4, mx=1e-4, what=f"split-K wgrad {M}x{N}x{K} via_transpose={via_transpose}")


@pytest.mark.parametrize("dim,heads", [(768, 32), (96, 3)])
def test_graph_fork_head_dim_below_64_trains_on_tensor_cores(mt, dim, heads):
    """Graph fork geometry (Graph/metatransformer/modules/tokengt_graph_encoder.py:193-203: dim 768, 32 heads -> head_dim
    24): the packed weights are zero-padded to 64 lanes per head, forward and backward run on the tcgen05 attention
    kernels; output, input gradient and all 12 parameter gradients vs fp64 autograd of the oracle."""
    sd = orc.make_encoder_state_dict(dim, 1, seed=12)
    sd = {k: (_bf16_round(v) if v.dim() == 2 else v) for k, v in sd.items()}
    blk = mt.Block(dim=dim, num_heads=heads, mlp_ratio=4., qkv_bias=True, norm_layer=torch.nn.LayerNorm,
                   act_layer=torch.nn.GELU)
    blk.load_state_dict({k[2:]: v for k, v in sd.items()}, strict=True)
    blk = blk.cuda().train()
    g = torch.Generator().manual_seed(5)
    B, N = 3, 70
    x = _bf16_round(torch.randn(B, N, dim, generator=g))
    gy = _bf16_round(torch.randn(B, N, dim, generator=g) * 0.05)
    ref_p = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    xr = x.double().requires_grad_(True)
    yr = orc.encoder(xr, ref_p, 1, heads)
    yr.backward(gy.double())
    xg = x.cuda().bfloat16().requires_grad_(True)
    y = blk(xg)
    _assert_close(y, yr.detach(), what=f"Block dim={dim} heads={heads} forward")
    y.backward(gy.cuda().bfloat16())
    _assert_close(xg.grad, xr.grad, rel=2e-2, mx=2e-2, what="dL/dx")
    for name, prm in blk.named_parameters():
        ref = ref_p["0." + name].grad
        assert prm.grad is not None and prm.grad.shape == ref.shape, name
        rel = float((prm.grad.double().cpu() - ref).norm() / ref.norm().clamp_min(1e-12))
        assert rel <= 3e-2, f"{name}: rel_l2 {rel:.3e}"
    with torch.no_grad():
        _assert_close(blk.eval()(x.cuda()), yr.detach(), what="eval forward, fp32 tokens")
