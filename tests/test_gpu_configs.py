"""GPU parity at every BASELINE.json configuration's own depth and shape (run with `-m gpu` on a B200).

C2  Base-16  x12, N = 196           C3  Large-14 x24, N = 256            (synthetic tokens -> encoder)
C4  fbank (B, 1024, 128) -> stride-16 spectrogram conv -> 512 tokens -> Base-16 x12
C5  clip (B, 3, 16, 224, 224) -> tubelet conv -> 1568 tokens -> Large-14 x24
Batch sizes are cut to 1-2 sequences (samples are independent, tests/test_gpu_parity.py checks that at B = 256); depth,
sequence length and widths are the configurations' own.  Gate (BASELINE north_star, "1e-2 bf16"), BOTH required:
    rel_l2 = ||got - ref|| / ||ref|| <= 1e-2      and      max|got - ref| / max(1, max|ref|) <= 1e-2
against the fp32 CPU oracle evaluated (a) on the unrounded fp32 weights the modules were loaded with and (b) on the
same weights rounded to bf16 (what the tensor cores multiply) — both are printed, both must pass.
"""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import encoder_oracle as orc  # noqa: E402


@pytest.fixture(scope="module")
def mt():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (B200)")
    import metatransformer_b200 as m
    from metatransformer_b200 import _lib
    _lib.load()
    return m


def _r16(t):
    return t.bfloat16().float()


def _errs(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape and torch.isfinite(got).all()
    return (((got - ref).norm() / ref.norm().clamp_min(1e-12)).item(),
            (got - ref).abs().max().item() / max(1.0, ref.abs().max().item()))


def _gate(got, refs, what, rel=1e-2, mx=1e-2):
    msgs, ok = [], True
    for name, ref in refs.items():
        r, m = _errs(got, ref)
        msgs.append(f"{name}: rel_l2={r:.3e} max_abs_norm={m:.3e}")
        ok = ok and r <= rel and m <= mx
    print(f"[parity] {what}: " + " | ".join(msgs))
    assert ok, f"{what}: " + " | ".join(msgs) + f" (gates {rel}, {mx})"


def _encoder(mt, scale, seed):
    dim, heads, depth = (768, 12, 12) if scale == "base" else (1024, 16, 24)
    sd = orc.make_encoder_state_dict(dim, depth, seed=seed)
    enc = mt.build_encoder(scale)
    enc.load_state_dict(sd, strict=True)
    sd16 = {k: (_r16(v) if v.dim() == 2 else v) for k, v in sd.items()}
    return enc.cuda().eval(), sd, sd16, dim, heads, depth


def _both(x, sd, sd16, depth, heads):
    with torch.no_grad():
        return {"oracle(fp32 weights)": orc.encoder(x, sd, depth, heads),
                "oracle(bf16-rounded weights)": orc.encoder(x, sd16, depth, heads)}


@pytest.mark.parametrize("scale,B,N", [("base", 2, 196), ("large", 2, 256), ("large", 1, 1568), ("base", 2, 512)])
def test_full_depth_encoder_every_baseline_shape(mt, scale, B, N):
    """C2 / C3 / C5 / C4 encoder shapes at their own depth: bf16 tokens through the fused driver and through the
    per-Block modules (nn.Sequential), fp32 tokens through both as well."""
    enc, sd, sd16, dim, heads, depth = _encoder(mt, scale, seed=3)
    x = _r16(orc.make_tokens(B, N, dim, seed=77))
    refs = _both(x, sd, sd16, depth, heads)
    fused = mt.FusedEncoder(enc)
    with torch.no_grad():
        y = fused(x.cuda().bfloat16())
        assert y.dtype == torch.bfloat16
        _gate(y, refs, f"{scale} x{depth} N={N} fused, bf16 tokens")
        # fp32 tokens under autocast — how the reference's audio / video forks call the encoder (ast_models.py:145,
        # engine_for_finetuning.py:97): fp32 residual stream in and out, bf16 tensor-core operands in between
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y32 = fused(x.cuda())
            assert y32.dtype == torch.float32
            _gate(y32, refs, f"{scale} x{depth} N={N} fused, fp32 tokens under autocast")
            if N <= 512:
                _gate(enc(x.cuda()), refs, f"{scale} x{depth} N={N} per-Block modules, fp32 tokens under autocast")


def test_c4_audio_tokenizer_into_base_encoder(mt):
    """BASELINE config 4 end to end: (B, T=1024, F=128) fbank -> (B,1,128,1024) -> Conv2d(1,768,16,stride 16) -> 512 tokens
    -> Base-16 x12  (Audio/src/models/ast_models.py:86,152-162 with fstride = tstride = 16)."""
    enc, sd, sd16, dim, heads, depth = _encoder(mt, "base", seed=5)
    tok = mt.data2seq.Acoustic.PatchEmbed(embed_dim=768, fstride=16, tstride=16).cuda().eval()
    g = torch.Generator().manual_seed(11)
    fbank = torch.randn(2, 1024, 128, generator=g)
    spec = orc.audio_input_prep(fbank)
    w, b = tok.proj.weight.detach().cpu(), tok.proj.bias.detach().cpu()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        t = tok(spec.cuda())
        assert t.shape == (2, 512, 768)
        y = mt.FusedEncoder(enc)(t)
        refs = {}
        for name, ww, ssd in (("oracle(fp32 weights)", w, sd), ("oracle(bf16-rounded weights)", _r16(w), sd16)):
            t_ref = orc.patch_embed_2d(spec, ww, b, (16, 16))
            refs[name] = orc.encoder(t_ref, ssd, depth, heads)
        _gate(t, {"oracle tokens": orc.patch_embed_2d(spec, w, b, (16, 16))}, "C4 tokens")
    _gate(y, refs, "C4 fbank -> tokens -> Base-16 x12")


def test_c5_video_tokenizer_into_large_encoder(mt):
    """BASELINE config 5 end to end: (B,3,16,224,224) -> Conv3d tubelets (2,16,16) -> 1568 tokens -> Large-14 x24
    (Video/models/modeling_finetune.py:285-297, 434-443)."""
    enc, sd, sd16, dim, heads, depth = _encoder(mt, "large", seed=6)
    tok = mt.data2seq.Video.PatchEmbed(embed_dim=1024).cuda().eval()
    g = torch.Generator().manual_seed(12)
    clip = torch.randn(1, 3, 16, 224, 224, generator=g)
    w, b = tok.proj.weight.detach().cpu(), tok.proj.bias.detach().cpu()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        t = tok(clip.cuda())
        assert t.shape == (1, 1568, 1024)
        y = mt.FusedEncoder(enc)(t)
        refs = {}
        for name, ww, ssd in (("oracle(fp32 weights)", w, sd), ("oracle(bf16-rounded weights)", _r16(w), sd16)):
            t_ref = orc.patch_embed_3d(clip, ww, b, (2, 16, 16))
            refs[name] = orc.encoder(t_ref, ssd, depth, heads)
    _gate(y, refs, "C5 clip -> tokens -> Large-14 x24")


@pytest.mark.parametrize("lo,hi", [(50.0, 500.0), (5.0, 20.0)])
def test_fused_driver_rows_with_mean_much_larger_than_std(mt, lo, hi):
    """LayerNorm stress through the FUSED driver: every token row = offset (|offset| in [lo, hi]) + N(0,1) noise.
    The branch outputs (y - x) are compared, so the offset cannot hide an error behind max|ref|."""
    enc, sd, sd16, dim, heads, _ = _encoder(mt, "base", seed=8)
    depth = 3
    enc = torch.nn.Sequential(*list(enc)[:depth])
    g = torch.Generator().manual_seed(21)
    B, N = 2, 196
    off = (torch.rand(B, N, 1, generator=g) * (hi - lo) + lo) * torch.where(torch.rand(B, N, 1, generator=g) < 0.5, -1.0, 1.0)
    x = torch.randn(B, N, dim, generator=g) + off
    with torch.no_grad():
        ref = orc.encoder(x.double(), {k: v.double() for k, v in sd.items()}, depth, heads)
        y = mt.FusedEncoder(enc, precision="bf16")(x.cuda())
    assert y.dtype == torch.float32
    _gate(y.cpu().double() - x.double(), {"oracle fp64 (branch sum y - x)": ref - x.double()},
          f"mean>>std rows (|offset| in [{lo},{hi}]), fused x{depth}")


def test_split_stream_residual_gemm_and_row_statistics(mt):
    """mt_gemm_bf16_ex with the split (hi + lo) residual stream: y + y_lo reproduces the fp64 sum to fp32-grade accuracy,
    a per-row offset is added before the split, and the emitted row-statistics records give the row mean / variance."""
    from metatransformer_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(31)
    M, N, K = 333, 768, 256
    x = _r16(torch.randn(M, K, generator=g))
    w = _r16(torch.randn(N, K, generator=g) * 0.05)
    b = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g) * 3 + torch.randn(M, 1, generator=g) * 40      # rows with a large common offset
    rh = r.bfloat16()
    rl = (r - rh.float()).bfloat16()
    coff = torch.randn(M, generator=g) * 10
    ref = x.double() @ w.double().t() + b.double() + rh.double() + rl.double() + coff.double()[:, None]
    dev = "cuda"
    xd, wd, bd, rhd, rld, cd = x.bfloat16().to(dev), w.bfloat16().to(dev), b.to(dev), rh.to(dev), rl.to(dev), coff.to(dev)
    y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    ylo = torch.empty_like(y)
    S = (N + 63) // 64
    rec = torch.zeros(M, S, 4, dtype=torch.float32, device=dev)
    for cg in (1, 2):
        o = _lib.GemmOpts(bias=bd.data_ptr(), residual=rhd.data_ptr(), residual_lo=rld.data_ptr(), y_lo=ylo.data_ptr(),
                          y_row_offset=cd.data_ptr(), out_row_sums=rec.data_ptr(), cta_group=cg)
        _lib.check(lib.mt_gemm_bf16_ex(xd.data_ptr(), wd.data_ptr(), y.data_ptr(), M, N, K, ctypes.byref(o),
                                       torch.cuda.current_stream().cuda_stream), "mt_gemm_bf16_ex")
        torch.cuda.synchronize()
        got = y.double().cpu() + ylo.double().cpu()
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        assert err < 3e-5, f"split stream cg={cg}: max rel err {err:.2e}"
        assert (y.double().cpu() - ref).abs().max().item() / ref.abs().max().item() < 2 ** -8   # hi alone = bf16(v)
        rc = rec.double().cpu()
        n_i = 64.0
        mean_i = rc[..., 0] + rc[..., 1] / n_i
        m2_i = rc[..., 2] - rc[..., 1] ** 2 / n_i
        mean = mean_i.mean(-1)
        var = (m2_i.sum(-1) + (n_i * (mean_i - mean[:, None]) ** 2).sum(-1)) / N
        assert (mean - ref.mean(-1)).abs().max() < 1e-3
        assert ((var - ref.var(-1, unbiased=False)).abs() / ref.var(-1, unbiased=False)).max() < 1e-4


# ---------------------------------------------------------------------------------------------------------------
# fp32-accuracy path (plain-fp32 callers, X-Ray/train.py:52-60): gate 1e-4 on BOTH metrics
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,gelu,resid,split", [(300, 768, 768, False, True, False), (197, 3072, 768, True, False, True),
                                                    (128, 96, 40, False, False, False), (513, 768, 3072, False, True, False)])
def test_gemm_f32x3_matches_fp64(mt, M, N, K, gelu, resid, split):
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g) * 2
    w = torch.randn(N, K, generator=g) * 0.05
    b = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g) if resid else None
    ref = x.double() @ w.double().t() + b.double()
    if gelu:
        ref = orc.gelu_erf(ref)
    if resid:
        ref = ref + r.double()
    xb, xs = mt.ops.split_tf32(x.cuda())
    wb, ws = mt.ops.split_tf32(w.cuda())
    assert torch.equal((xb + xs).cpu(), x)                        # the split is exact
    y = mt.ops.gemm_f32x3(xb, xs, wb, ws, bias=b.cuda(), residual=None if r is None else r.cuda(), gelu=gelu, split_out=split)
    if split:
        assert torch.equal((y[0].view(torch.int32) & 0x1FFF), torch.zeros_like(y[0], dtype=torch.int32))
        y = y[0] + y[1]
    err = (y.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 6e-5, f"3xTF32 GEMM {M}x{N}x{K}: max rel err {err:.2e}"      # ~2^-21 per product, fp32 accumulation over K


@pytest.mark.parametrize("B,N,H,dh", [(2, 196, 3, 64), (1, 300, 2, 64), (2, 50, 4, 24), (1, 17, 2, 32)])
def test_attention_and_layernorm_f32(mt, B, N, H, dh):
    g = torch.Generator().manual_seed(N + dh)
    C = H * dh
    qkv = torch.randn(B, N, 3 * C, generator=g)
    q, k, v = qkv.double().reshape(B, N, 3, H, dh).permute(2, 0, 3, 1, 4).unbind(0)
    ref = (((q @ k.transpose(-2, -1)) * dh ** -0.5).softmax(-1) @ v).transpose(1, 2).reshape(B, N, C)
    ob, os_ = mt.ops.attention_f32(qkv.cuda(), H)
    err = ((ob + os_).double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, f"fp32 attention: {err:.2e}"
    x = torch.randn(B * N, C, generator=g) * 3 + 1.5
    w, b = 1 + torch.randn(C, generator=g) * 0.2, torch.randn(C, generator=g) * 0.2
    hb, hs = mt.ops.layernorm_f32_split(x.cuda(), w.cuda(), b.cuda(), 1e-5)
    ref = orc.layer_norm(x.double(), w.double(), b.double(), 1e-5)
    assert ((hb + hs).double().cpu() - ref).abs().max().item() < 5e-6 * ref.abs().max().item()


@pytest.mark.parametrize("scale,B,N", [("base", 2, 196), ("large", 1, 256)])
def test_fp32_path_full_depth_gate_1e_4(mt, scale, B, N):
    """fp32 tokens outside autocast select the fp32-accurate path automatically (per-Block modules and FusedEncoder);
    full depth, fp64 oracle on the unrounded weights, gate 1e-4 on rel-L2 and on max-abs / max|ref|."""
    enc, sd, sd16, dim, heads, depth = _encoder(mt, scale, seed=13)
    x = orc.make_tokens(B, N, dim, seed=5)
    with torch.no_grad():
        ref = orc.encoder(x.double(), {k: v.double() for k, v in sd.items()}, depth, heads)
        y = enc(x.cuda())
        yf = mt.FusedEncoder(enc)(x.cuda())
        yb = mt.FusedEncoder(enc, precision="bf16")(x.cuda())
    assert y.dtype == torch.float32 and torch.equal(y, yf)
    _gate(y, {"oracle fp64": ref}, f"{scale} x{depth} N={N} fp32 path", rel=1e-4, mx=1e-4)
    r_bf16, _ = _errs(yb, ref)
    assert r_bf16 > 5e-4                      # precision="bf16" really is the other path
