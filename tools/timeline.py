#!/usr/bin/env python
"""Device timeline of one CUDA-graph replay of the encoder forward (no nsys in this image).

    python tools/timeline.py [--config c2|c3|c4|c5] [--out gpurun_out/timeline_c2.json]

Every instrumented kernel (entry pass, GEMMs, attention) records min(%globaltimer) over its CTAs at start and
max(%globaltimer) at exit into its own slot (mt_timeline_begin in include/mt_b200.h).  The launches are captured into a
CUDA graph with the timeline armed, the graph is replayed until the clocks settle, the slots are reset and ONE replay is
recorded: start / end of each kernel inside a steady-state step, hence also the gaps between kernels that per-kernel
durations cannot show.  Prints a table and writes JSON.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

CONFIGS = {"c2": ("base", 256, 196), "c3": ("large", 128, 256), "c4": ("base", 64, 512), "c5": ("large", 32, 1568)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--out", default=None)
    ap.add_argument("--settle", type=int, default=60, help="graph replays before the recorded one")
    args = ap.parse_args()
    import torch
    import metatransformer_b200 as mt
    from metatransformer_b200 import _lib, synthetic

    scale, B, N = CONFIGS[args.config]
    dim, depth = (768, 12) if scale == "base" else (1024, 24)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    enc = mt.build_encoder(scale)
    enc.load_state_dict(synthetic.make_encoder_state_dict(dim, depth, seed=0), strict=True)
    enc = enc.to(dev).eval()
    fused = mt.FusedEncoder(enc, use_cuda_graph=False)
    x = synthetic.make_tokens(B, N, dim, seed=1234).bfloat16().to(dev)
    y = torch.empty_like(x)
    from metatransformer_b200 import timeline
    res = timeline.record(fused, x, y, settle=args.settle)
    rows, summary = res["rows"], res["summary"]
    summary.update({"config": args.config, "B": B, "N": N})
    print(f"{'kernel':16s} {'start_us':>10s} {'dur_us':>9s} {'gap_us':>8s}")
    for r in rows[: 1 + 5 * 3] + rows[-5:]:
        gp = "" if r["gap_before_us"] is None else f"{r['gap_before_us']:8.2f}"
        print(f"{r['kernel']:16s} {r['start_us']:10.2f} {r['dur_us']:9.2f} {gp:>8s}")
    print(json.dumps(summary, indent=1))
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps({"summary": summary, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
