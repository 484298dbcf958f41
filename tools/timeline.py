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
    lib = _lib.load()
    with torch.no_grad():
        fused(x, out=y)                       # warm-up outside capture (function attributes, weight packing)
        torch.cuda.synchronize()
        cap = fused.launches_per_forward + 8
        slots = torch.zeros(2 * cap, dtype=torch.int64, device=dev)
        _lib.check(lib.mt_timeline_begin(slots.data_ptr(), cap), "mt_timeline_begin")
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fused(x, out=y)
        used = lib.mt_timeline_end()
        for _ in range(args.settle):
            g.replay()
        torch.cuda.synchronize()
        slots[0::2] = -1                      # 0xFFFF... for atomicMin
        slots[1::2] = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1)
    t = slots.cpu().tolist()
    names = ["center_split"]
    for l in range(depth):
        names += [f"L{l}.qkv", f"L{l}.attention", f"L{l}.proj", f"L{l}.fc1", f"L{l}.fc2"]
    rows = []
    t0 = min(t[2 * i] for i in range(used))
    prev_end = None
    for i in range(used):
        s, e = t[2 * i] - t0, t[2 * i + 1] - t0
        rows.append({"kernel": names[i] if i < len(names) else f"k{i}", "start_us": s / 1e3, "end_us": e / 1e3,
                     "dur_us": (e - s) / 1e3, "gap_before_us": None if prev_end is None else (s - prev_end) / 1e3})
        prev_end = e
    total_us = max(r["end_us"] for r in rows)
    kinds = {}
    for r in rows:
        k = r["kernel"].split(".")[-1]
        kinds.setdefault(k, [0.0, 0])
        kinds[k][0] += r["dur_us"]
        kinds[k][1] += 1
    gaps = [r["gap_before_us"] for r in rows if r["gap_before_us"] is not None]
    summary = {"config": args.config, "B": B, "N": N, "kernels": used, "span_us": total_us, "event_step_ms": step_ms,
               "sum_kernel_us": sum(r["dur_us"] for r in rows), "sum_gap_us": sum(gaps),
               "mean_gap_us": sum(gaps) / max(len(gaps), 1), "max_gap_us": max(gaps) if gaps else 0.0,
               "by_kind_us": {k: {"total": v[0], "launches": v[1], "mean": v[0] / v[1]} for k, v in kinds.items()}}
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
