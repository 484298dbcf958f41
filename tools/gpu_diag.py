"""GPU bring-up diagnostics: each group runs in its own subprocess (a trapped kernel cannot poison the rest).

    python tools/gpu_diag.py            # run every group, logs under gpurun_out/diag/
    python tools/gpu_diag.py gemm1      # run one group in-process

Not a test-suite (tests/ holds the parity tests); this prints error *patterns* to debug descriptor/layout bugs.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _report(name, got, ref, tol=2e-2):
    import torch
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    rel = err.max().item() / denom
    bad = err > tol * max(denom, 1e-6)
    nb = int(bad.sum())
    nan = int(torch.isnan(got).sum())
    status = "OK " if (nb == 0 and nan == 0) else "BAD"
    print(f"[{status}] {name}: max_abs_err={err.max().item():.4g} rel_to_max={rel:.3g} ref_max={denom:.4g} "
          f"bad={nb}/{bad.numel()} nan={nan}", flush=True)
    if nb and got.dim() == 2:
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        print(f"      bad rows: n={rows.numel()} first={rows[:12].tolist()} last={rows[-4:].tolist()}")
        print(f"      bad cols: n={cols.numel()} first={cols[:12].tolist()} last={cols[-4:].tolist()}")
        r, c = int(rows[0]), int(cols[0])
        print(f"      got[{r},{c}:{c+8}]={got[r, c:c+8].tolist()}")
        print(f"      ref[{r},{c}:{c+8}]={ref[r, c:c+8].tolist()}")
    return nb == 0 and nan == 0


def group_basic():
    import torch
    from metatransformer_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(1000, 768, device=dev)
    y = ops.cast(x, torch.bfloat16)
    _report("cast f32->bf16", y, x.bfloat16())
    for (M, d, dt) in [(777, 768, torch.bfloat16), (513, 1024, torch.float32), (100, 200, torch.bfloat16),
                       (64, 4096, torch.float32)]:
        x = (torch.randn(M, d, device=dev) * 2 + 0.5).to(dt)
        g = torch.randn(d, device=dev) * 0.1 + 1
        b = torch.randn(d, device=dev) * 0.1
        y = ops.layernorm(x, g, b, 1e-5)
        ref = torch.nn.functional.layer_norm(x.float(), (d,), g, b, 1e-5)
        _report(f"layernorm M={M} d={d} {dt}", y, ref, tol=1e-2)
    # im2row
    for (shape, k, s) in [((2, 3, 1, 224, 224), (1, 16, 16), (1, 16, 16)), ((2, 3, 4, 32, 32), (2, 16, 16), (2, 16, 16)),
                          ((2, 1, 1, 128, 100), (1, 16, 16), (1, 10, 10))]:
        x = torch.randn(*shape, device=dev)
        rows, (Tp, Hp, Wp) = ops.im2row(x, k, s)
        B, C = shape[0], shape[1]
        pat = x.unfold(2, k[0], s[0]).unfold(3, k[1], s[1]).unfold(4, k[2], s[2])  # B,C,Tp,Hp,Wp,kt,kh,kw
        ref = pat.permute(0, 2, 3, 4, 1, 5, 6, 7).reshape(B * Tp * Hp * Wp, -1)
        _report(f"im2row {shape} k={k} s={s}", rows, ref.bfloat16())
    torch.cuda.synchronize()


def _gemm_case(cg, M, N, K, bias, gelu, resid, ints=False, seed=0):
    import torch
    from metatransformer_b200 import ops
    torch.manual_seed(seed)
    dev = "cuda"
    if ints:
        x = torch.randint(-2, 3, (M, K), device=dev).bfloat16()
        w = torch.randint(-2, 3, (N, K), device=dev).bfloat16()
    else:
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = (torch.randn(N, device=dev) * 0.5) if bias else None
    r = torch.randn(M, N, device=dev).bfloat16() if resid else None
    y = ops.linear(x, w, b, r, gelu=gelu, cta_group=cg)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t()
    if bias:
        ref = ref + b
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    if resid:
        ref = ref + r.float()
    return _report(f"gemm cg={cg} M={M} N={N} K={K} bias={bias} gelu={gelu} resid={resid} ints={ints}", y, ref,
                   tol=1e-2)


def _group_gemm(cg):
    ok = True
    ok &= _gemm_case(cg, 128 * cg, 256, 64, False, False, False, ints=True)
    ok &= _gemm_case(cg, 128 * cg, 256, 128, False, False, False, ints=True)
    ok &= _gemm_case(cg, 128 * cg, 256, 768, False, False, False)
    ok &= _gemm_case(cg, 256 * cg, 512, 256, True, False, False)
    ok &= _gemm_case(cg, 300, 768, 768, True, False, True)
    ok &= _gemm_case(cg, 1568, 2304, 768, True, False, False)
    ok &= _gemm_case(cg, 1568, 3072, 768, True, True, False)
    ok &= _gemm_case(cg, 1568, 768, 3072, True, False, True)
    ok &= _gemm_case(cg, 50176, 768, 768, True, False, True)
    ok &= _gemm_case(cg, 1000, 320, 200, True, True, True)
    return ok


def group_gemm1():
    _group_gemm(1)


def group_gemm2():
    _group_gemm(2)


def _attn_case(B, N, H, seed=0):
    import torch
    from metatransformer_b200 import ops
    torch.manual_seed(seed)
    dev = "cuda"
    qkv = torch.randn(B, N, 3 * H * 64, device=dev).bfloat16()
    out = ops.attention(qkv, H)
    torch.cuda.synchronize()
    q, k, v = qkv.float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
    att = (q @ k.transpose(-2, -1)) * 0.125
    ref = (att.softmax(-1) @ v).transpose(1, 2).reshape(B, N, H * 64)
    return _report(f"attention B={B} N={N} H={H}", out.reshape(B * N, -1), ref.reshape(B * N, -1), tol=1e-2)


def group_attn():
    _attn_case(1, 128, 1)
    _attn_case(1, 64, 2)
    _attn_case(2, 196, 3)
    _attn_case(2, 256, 2)
    _attn_case(1, 300, 2)
    _attn_case(1, 512, 12)
    _attn_case(2, 1568, 4)
    _attn_case(64, 196, 12)


def group_perf():
    import torch
    from metatransformer_b200 import ops
    dev = "cuda"
    torch.manual_seed(0)

    def timeit(fn, iters=20, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    M = 50176
    for (N, K, gelu, resid, name) in [(2304, 768, False, False, "qkv"), (768, 768, False, True, "proj"),
                                      (3072, 768, True, False, "fc1"), (768, 3072, False, True, "fc2")]:
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        b = torch.randn(N, device=dev)
        r = torch.randn(M, N, device=dev).bfloat16() if resid else None
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for cg in (1, 2):
            ms = timeit(lambda: ops.linear(x, w, b, r, gelu=gelu, cta_group=cg, out=out))
            print(f"[perf] gemm {name} cg={cg} M={M} N={N} K={K}: {ms:.4f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s",
                  flush=True)
        ms = timeit(lambda: torch.nn.functional.linear(x, w))
        print(f"[perf] cublas {name}: {ms:.4f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    for (B, N, H) in [(256, 196, 12), (128, 256, 16), (64, 512, 12), (32, 1568, 16)]:
        qkv = torch.randn(B, N, 3 * H * 64, device=dev).bfloat16()
        out = torch.empty(B, N, H * 64, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: ops.attention(qkv, H, out=out))
        fl = 4.0 * B * H * N * N * 64
        print(f"[perf] attention B={B} N={N} H={H}: {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    for (B, N, H) in [(256, 196, 12), (32, 1568, 16)]:
        qkv = torch.randn(B, N, 3 * H * 64, device=dev).bfloat16()
        o, lse = ops.attention_lse(qkv, H)
        do = torch.randn_like(o)
        ms = timeit(lambda: ops.attention_bwd(qkv, o, do, lse, H), iters=10)
        fl = 10.0 * B * H * N * N * 64
        print(f"[perf] attention_bwd B={B} N={N} H={H}: {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    xx = torch.randn(M, 768, device=dev).bfloat16()
    st = ops.row_stats(xx)
    dyy = torch.randn_like(xx)
    ms = timeit(lambda: ops.layernorm_bwd(dyy, xx, st, None, dyy))
    print(f"[perf] layernorm_bwd M={M} d=768: {ms:.4f} ms  {M * 768 * 8 / ms / 1e6:.1f} GB/s", flush=True)
    tt = torch.randn(M, 3072, device=dev).bfloat16()
    ms = timeit(lambda: ops.transpose(tt), iters=10)
    print(f"[perf] transpose M={M} x 3072: {ms:.4f} ms  {M * 3072 * 4 / ms / 1e6:.1f} GB/s", flush=True)
    x = torch.randn(M, 768, device=dev).bfloat16()
    g = torch.ones(768, device=dev)
    bb = torch.zeros(768, device=dev)
    out = torch.empty_like(x)
    ms = timeit(lambda: ops.layernorm(x, g, bb, out=out))
    print(f"[perf] layernorm M={M} d=768: {ms:.4f} ms  {M * 768 * 4 / ms / 1e6:.1f} GB/s", flush=True)


def group_afw():
    """attention forward only, C2 shape, 50 timed iterations (ablation runs via MT_B200_AFW_SKIP)"""
    import torch
    from metatransformer_b200 import ops
    dev = torch.device("cuda")
    for (B, N, H) in [(256, 196, 12), (32, 1568, 16)]:
        qkv = torch.randn(B, N, 3 * H * 64, device=dev).bfloat16()
        out = torch.empty(B, N, H * 64, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.attention(qkv, H, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            ops.attention(qkv, H, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        print(f"[afw skip={os.environ.get('MT_B200_AFW_SKIP', '0')}] B={B} N={N} H={H}: {ms:.4f} ms  "
              f"{4.0 * B * H * N * N * 64 / ms / 1e9:.1f} TFLOP/s", flush=True)


def group_wgrad():
    """fine-tune backward pieces at C2 size: the four weight-gradient GEMMs, bias column sums, recomputations"""
    import torch
    from metatransformer_b200 import ops
    dev = torch.device("cuda")
    M = 50176

    def timeit(fn, iters=10):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    for name, n, k in [("qkv", 2304, 768), ("proj", 768, 768), ("fc1", 3072, 768), ("fc2", 768, 3072)]:
        dy = torch.randn(M, n, device=dev).bfloat16()
        x = torch.randn(M, k, device=dev).bfloat16()
        ms = timeit(lambda: ops.linear_wgrad(dy, x))
        print(f"[wgrad] dW {name} [{n}x{k}] over {M} tokens: {ms:.4f} ms  {2.0 * M * n * k / ms / 1e9:.1f} TFLOP/s", flush=True)
        ms = timeit(lambda: ops.colsum(dy))
        print(f"[wgrad] colsum [{M}x{n}]: {ms:.4f} ms  {M * n * 2 / ms / 1e6:.1f} GB/s", flush=True)
    u = torch.randn(M, 3072, device=dev).bfloat16()
    ms = timeit(lambda: ops.gelu(u))
    print(f"[wgrad] gelu [{M}x3072]: {ms:.4f} ms  {M * 3072 * 4 / ms / 1e6:.1f} GB/s", flush=True)
    w = torch.randn(3072, 768, device=dev)
    dwf = torch.randn(3072, 768, device=dev)
    g = torch.ones(768, device=dev)
    b = torch.zeros(768, device=dev)
    dyb = torch.randn(3072, device=dev)
    ms = timeit(lambda: ops.ln_fold_bwd(dwf, w, g, b, dyb))
    print(f"[wgrad] ln_fold_bwd [3072x768]: {ms:.4f} ms", flush=True)


def group_abw():
    """attention backward only: the four BASELINE shapes, 30 timed iterations each (MT_B200_ATTN_BWD_OWNER=0|1 for A/B)."""
    import torch
    dev = torch.device("cuda")
    from metatransformer_b200 import ops
    mode = os.environ.get("MT_B200_ATTN_BWD_OWNER", "auto")
    iters = int(os.environ.get("ABW_ITERS", "30"))
    shapes = [(256, 196, 12), (128, 256, 16), (64, 512, 12), (32, 1568, 16)]
    if "ABW_ONLY" in os.environ:
        shapes = [shapes[int(os.environ["ABW_ONLY"])]]
    for (B, N, H) in shapes:
        qkv = torch.randn(B, N, 3 * H * 64, device=dev).bfloat16()
        o, lse = ops.attention_lse(qkv, H)
        do = torch.randn_like(o)
        for _ in range(3):
            ops.attention_bwd(qkv, o, do, lse, H)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            ops.attention_bwd(qkv, o, do, lse, H)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"[abw owner={mode}] B={B} N={N} H={H}: {ms:.4f} ms  {10.0 * B * H * N * N * 64 / ms / 1e9:.1f} TFLOP/s", flush=True)


def group_prof():
    """tiny launch set for ncu: fc2-shaped GEMM, cg=1 then cg=2, plus attention and layernorm"""
    import torch
    from metatransformer_b200 import ops
    dev = "cuda"
    M, N, K = 50176, 768, 3072
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for cg in (1, 2):
        for _ in range(2):
            ops.linear(x, w, b, r, cta_group=cg, out=out)
    torch.cuda.synchronize()
    qkv = torch.randn(256, 196, 3 * 768, device=dev).bfloat16()
    for _ in range(2):
        ops.attention(qkv, 12)
    torch.cuda.synchronize()


def group_prof_fc1():
    """fc1-shaped GEMM (K=768, N=3072, GELU epilogue), cg=2, for an ncu capture of the epilogue-heavy case"""
    import torch
    from metatransformer_b200 import ops
    dev = "cuda"
    M, N, K = 50176, 3072, 768
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.linear(x, w, b, None, gelu=True, cta_group=2, out=out)
    torch.cuda.synchronize()


def group_prof_fc2():
    """fc2-shaped GEMM (K=3072, N=768, residual), cg=2: main-loop dominated"""
    import torch
    from metatransformer_b200 import ops
    dev = "cuda"
    M, N, K = 50176, 768, 3072
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.linear(x, w, b, r, cta_group=2, out=out)
    torch.cuda.synchronize()


def group_prof_layer():
    """one Base-16 layer's worth of kernels at C2 size through the fused driver (for ncu --set full captures)"""
    import torch
    import metatransformer_b200 as mt
    from metatransformer_b200 import synthetic as orc
    enc = mt.build_encoder("base", depth=2)
    enc.load_state_dict(orc.make_encoder_state_dict(768, 2, seed=0), strict=True)
    enc = enc.cuda().eval()
    fused = mt.FusedEncoder(enc)
    x = torch.randn(256, 196, 768).bfloat16().cuda()
    with torch.no_grad():
        for _ in range(2):
            fused(x)
    torch.cuda.synchronize()


GROUPS = {"afw": group_afw, "wgrad": group_wgrad, "abw": group_abw, "prof": group_prof, "prof_layer": group_prof_layer, "prof_fc1": group_prof_fc1, "prof_fc2": group_prof_fc2, "basic": group_basic, "gemm1": group_gemm1, "gemm2": group_gemm2, "attn": group_attn, "perf": group_perf}


def main():
    if len(sys.argv) > 1 and sys.argv[1] in GROUPS:
        GROUPS[sys.argv[1]]()
        return
    names = sys.argv[1:] if len(sys.argv) > 1 else [g for g in GROUPS if not g.startswith("prof")]
    outdir = ROOT / "gpurun_out" / "diag"
    outdir.mkdir(parents=True, exist_ok=True)
    for name in names:
        t0 = time.time()
        log = outdir / f"{name}.log"
        with open(log, "w") as f:
            try:
                r = subprocess.run([sys.executable, __file__, name], stdout=f, stderr=subprocess.STDOUT, timeout=300,
                                   cwd=str(ROOT))
                rc = r.returncode
            except subprocess.TimeoutExpired:
                rc = "TIMEOUT"
        txt = log.read_text()
        print(f"===== group {name}: rc={rc} ({time.time() - t0:.1f}s) =====")
        print(txt[-6000:])
        sys.stdout.flush()


if __name__ == "__main__":
    main()
