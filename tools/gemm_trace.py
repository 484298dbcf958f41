"""Where the warps of the GEMM kernel wait (needs a -DMT_GEMM_TRACE=1 build, see tools/build_variant.py):

    python tools/build_variant.py gtrace mt_gemm.cu -DMT_GEMM_TRACE=1
    MT_B200_LIB=metatransformer_b200/lib/variants/libmt_b200_gtrace.so python tools/gemm_trace.py --config c2

For each of the four Linear shapes of a layer: cycles (clock64) that the TMA producer, the MMA issuer and the 16 epilogue
warps of CTA 0 spend inside their barrier waits, and the epilogue's time per tile before / after the accumulator wait.
"""
import argparse
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from metatransformer_b200 import _lib, ops  # noqa: E402

CONFIGS = {"c2": (256, 196, 768), "c3": (128, 256, 1024), "c4": (64, 512, 768), "c5": (32, 1568, 1024)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    a = ap.parse_args()
    B, N, d = CONFIGS[a.config]
    M = B * N
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).bfloat16().cuda()

    x, hid = rnd(M, d), rnd(M, 4 * d)
    r, rl = rnd(M, d), rnd(M, d, scale=0.004)
    stats = torch.zeros(M, 2, device="cuda")
    stats[:, 1] = 1
    y3, y4, y1, ylo = (torch.empty(M, n, dtype=torch.bfloat16, device="cuda") for n in (3 * d, 4 * d, d, d))
    w_qkv, w_fc1, w_proj, w_fc2 = rnd(3 * d, d, scale=0.02), rnd(4 * d, d, scale=0.02), rnd(d, d, scale=0.02), rnd(d, 4 * d, scale=0.02)
    z = lambda n: torch.zeros(n, device="cuda")  # noqa: E731
    shapes = {
        "qkv (LN fold)": lambda: ops.linear_ln(x, w_qkv, z(3 * d), z(3 * d), stats, out=y3),
        "fc1 (LN fold + GELU)": lambda: ops.linear_ln(x, w_fc1, z(4 * d), z(4 * d), stats, gelu=True, out=y4),
        "proj (split stream)": lambda: ops.linear(x, w_proj, z(d), residual=r, residual_lo=rl, out=y1, out_lo=ylo),
        "fc2 (split stream)": lambda: ops.linear(hid, w_fc2, z(d), residual=r, residual_lo=rl, out=y1, out_lo=ylo),
        "proj (bf16 stream)": lambda: ops.linear(x, w_proj, z(d), residual=r, out=y1),
    }
    buf = torch.zeros(4 * 20 * 8, dtype=torch.int64, device="cuda")
    for name, fn in shapes.items():
        for _ in range(3):
            fn()
        buf.zero_()
        if lib.mt_gemm_set_trace(ctypes.c_void_p(buf.data_ptr())) == 1:
            raise SystemExit("this build has no counters: build with -DMT_GEMM_TRACE=1 and set MT_B200_LIB")
        fn()
        torch.cuda.synchronize()
        lib.mt_gemm_set_trace(ctypes.c_void_p(0))
        w = buf.cpu().reshape(4, 20, 8)
        print(f"== {name}  (M, d) = ({M}, {d})")
        for blk in (0, 1):
            tot = int(w[blk, 0, 4])
            pct = lambda v: f"{int(v):7d} ({100.0 * int(v) / max(tot, 1):4.1f}%)"  # noqa: E731
            print(f" CTA {blk}: kernel {tot} clk | producer waits for a free stage {pct(w[blk, 0, 0])}"
                  f" | MMA issuer waits: accumulator drained {pct(w[blk, 1, 0])}, operands landed {pct(w[blk, 1, 1])}")
            ep = w[blk, 4:20].double()
            print(f"        epilogue warps (mean of 16): accumulator ready wait {pct(ep[:, 0].mean())}, TMA store read wait"
                  f" {pct(ep[:, 1].mean())}, tile prologue (bias / residual requests) {pct(ep[:, 2].mean())},"
                  f" pieces incl. store waits {pct(ep[:, 3].mean())}")


if __name__ == "__main__":
    main()
