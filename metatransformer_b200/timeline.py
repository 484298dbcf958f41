"""Device timeline of one CUDA-graph replay of a FusedEncoder forward (measurement aid; no nsys in this image).

Every instrumented kernel (entry pass, GEMMs, attention) records min(%globaltimer) over its CTAs at start and
max(%globaltimer) at exit into its own slot (mt_timeline_begin, include/mt_b200.h).  The launches are captured into a
CUDA graph with the timeline armed, the graph is replayed `settle` times so that clocks and power reach the sustained
regime, the slots are reset and ONE more replay is recorded: start / end of every kernel inside a steady-state step,
hence also the gaps between kernels that per-kernel durations cannot show.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import _lib


def kernel_names(depth: int) -> List[str]:
    names = ["center_split"]
    for l in range(depth):
        if l > 0:
            names.append(f"L{l}.stats1")
        names += [f"L{l}.qkv", f"L{l}.attention", f"L{l}.proj", f"L{l}.stats2", f"L{l}.fc1", f"L{l}.fc2"]
    return names


def record(fused, x: torch.Tensor, y: torch.Tensor, settle: int = 60) -> Dict:
    """Returns {"rows": [{kernel, start_us, end_us, dur_us, gap_before_us}], "summary": {...}} for one replay."""
    lib = _lib.load()
    depth = len(fused.blocks)
    with torch.no_grad():
        fused(x, out=y)                       # warm-up outside capture (function attributes, weight packing)
        torch.cuda.synchronize()
        cap = fused.launches_per_forward + 8
        slots = torch.zeros(2 * cap, dtype=torch.int64, device=x.device)
        _lib.check(lib.mt_timeline_begin(slots.data_ptr(), cap), "mt_timeline_begin")
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fused(x, out=y)
        finally:
            used = lib.mt_timeline_end()
        for _ in range(settle):
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
    names = kernel_names(depth)
    t0 = min(t[2 * i] for i in range(used))
    rows, prev_end = [], None
    for i in range(used):
        s, e = t[2 * i] - t0, t[2 * i + 1] - t0
        rows.append({"kernel": names[i] if i < len(names) else f"k{i}", "start_us": s / 1e3, "end_us": e / 1e3,
                     "dur_us": (e - s) / 1e3, "gap_before_us": None if prev_end is None else (s - prev_end) / 1e3})
        prev_end = e
    kinds: Dict[str, List[float]] = {}
    for r in rows:
        k = r["kernel"].split(".")[-1]
        kinds.setdefault(k, [0.0, 0])
        kinds[k][0] += r["dur_us"]
        kinds[k][1] += 1
    gaps = [r["gap_before_us"] for r in rows if r["gap_before_us"] is not None]
    summary = {"kernels": used, "span_us": max(r["end_us"] for r in rows), "event_step_ms": step_ms,
               "sum_kernel_us": sum(r["dur_us"] for r in rows), "sum_gap_us": sum(gaps),
               "mean_gap_us": sum(gaps) / max(len(gaps), 1), "max_gap_us": max(gaps) if gaps else 0.0,
               "by_kind_us": {k: {"total": v[0], "launches": v[1], "mean": v[0] / v[1]} for k, v in kinds.items()}}
    return {"rows": rows, "summary": summary}
