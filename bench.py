#!/usr/bin/env python
"""bench.py — headline benchmark: encoder tokens/sec on the BASELINE.json configuration
  C2 = Base-16 encoder (12 x timm Block, d=768, 12 heads), synthetic bf16 tokens (B, N, d) = (256, 196, 768) per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 5            # our arm (C-ABI kernels on cuda:LOCAL_RANK)
    python bench.py --impl reference --gpus 1 --steps 3       # reference arm: the CPU oracle port on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                # one rank per GPU, weak scaling (replicas, no collective)

One "step" = one forward of the whole encoder over one batch.  Prints ONE JSON line (rank 0).
Keys: value = device-resident whole-job tokens/s (CUDA events, max over ranks); e2e = same metric through the public
API with pinned HOST buffers (H2D of the tokens and D2H of the features inside the timed region, software-pipelined
over 3 slots); roofline = the dominant kernel (gemm_bf16_kernel, all four Linear layers) timed on the device inside
>= 2 s of continuous graph replays (globaltimer timeline), roofline_others = attention and the image tokenizer;
cpu_baseline = the oracle port on this box's host cores (bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

B, N, D, DEPTH, HEADS = 256, 196, 768, 12, 12
METRIC = "encoder tokens/sec at (B,N,d)=(256,196,768)"
UNIT = "tokens/s"
CPU_SAMPLE_B = 8            # bounded CPU sample: 8 of the 256 sequences (BASELINE config 1 batch)


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on the host cores (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------------------------
def cpu_oracle_tokens_per_s(steps: int, warmup: int, budget_s: float = 25.0):
    import torch
    from oracle import encoder_oracle as orc
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sd = orc.make_encoder_state_dict(D, DEPTH, seed=0)
    x = orc.make_tokens(CPU_SAMPLE_B, N, D, seed=1234)
    # use as many host threads as actually help: probe powers of two up to the available cores on one Block
    # (oversubscribed intra-op pools make eager PyTorch slower, and the reference deserves its best setting)
    best, threads = None, 1
    sd1 = {k: v for k, v in sd.items() if k.startswith("0.")}
    cand = sorted({c for c in (4, 8, 16, 32, 64, 128, avail) if c <= avail} | {min(avail, 8)})
    with torch.no_grad():
        for c in cand:
            torch.set_num_threads(c)
            orc.encoder(x[:2], sd1, 1, HEADS)
            t0 = time.perf_counter()
            orc.encoder(x, sd1, 1, HEADS)
            dt1 = time.perf_counter() - t0
            if best is None or dt1 < best:
                best, threads = dt1, c
    torch.set_num_threads(threads)
    with torch.no_grad():
        for _ in range(warmup):
            orc.encoder(x, sd, DEPTH, HEADS)
        t0 = time.perf_counter()
        done = 0
        for _ in range(steps):
            orc.encoder(x, sd, DEPTH, HEADS)
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    tps = done * CPU_SAMPLE_B * N / dt
    return dict(value=tps, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"{done} forward(s) of the fp32 oracle port on {CPU_SAMPLE_B}x{N} tokens "
                       f"({CPU_SAMPLE_B}/{B} of one C2 batch), {dt:.2f} s, {threads} of {avail} host threads "
                       f"(best of {cand})"), dt / max(done, 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    base, sec_per_step = cpu_oracle_tokens_per_s(max(args.steps, 1), min(args.warmup, 1), budget_s=120.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2 Base-16 encoder forward; reference = CPU oracle port (timm Block restatement), "
                               f"bounded sample of {CPU_SAMPLE_B} sequences per step", "global_batch": CPU_SAMPLE_B,
                   "seq_len": N, "parallelism": "host threads"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# clocks sampling
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        clocks, maxc, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                clocks.append(float(f[1])); maxc = float(f[2]); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not clocks:
            return {"sm_mhz": None, "sm_max_mhz": maxc, "reasons": ["no samples"]}
        # "under load": samples drawing more than half of the max observed power
        thr = 0.5 * max(power)
        load = [c for c, p in zip(clocks, power) if p >= thr] or clocks
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": maxc, "power_w_max": max(power),
                "samples": len(clocks), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import metatransformer_b200 as mt
    from metatransformer_b200 import ops
    from metatransformer_b200.distributed import init_process_group
    from metatransformer_b200 import synthetic

    rank, local_rank, world = init_process_group("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sustained_tf, burst_tf, hbm_gbs, peak_src = _peaks()

    # ---- model: identical seeded weights on every rank, inputs seeded per rank ----
    sd = synthetic.make_encoder_state_dict(D, DEPTH, seed=0)
    enc = mt.build_encoder("base")
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(dev).eval()
    fused = mt.FusedEncoder(enc, use_cuda_graph=args.cuda_graph)
    NBUF = 4   # rotating inputs: 4 x 77 MB = 308 MB > 126 MB L2 (plus ~1.3 GB of intermediates written per step)
    g = torch.Generator().manual_seed(1234 + rank)
    xs = [torch.randn(B, N, D, generator=g).bfloat16().to(dev) for _ in range(NBUF)]
    y = torch.empty_like(xs[0])
    tokens_per_step = B * N

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity spot check (first 2 sequences): the oracle is used here as the checker only ----
    from oracle import encoder_oracle as orc
    with torch.no_grad():
        got = fused(xs[0][:2].contiguous()).float().cpu()
    x_chk = xs[0][:2].float().cpu()
    w16 = {k: (v.bfloat16().float() if v.dim() == 2 else v) for k, v in sd.items()}

    def _par(ref):
        return {"rel_l2_err": float((got - ref).norm() / ref.norm()),
                "max_abs_err_over_max_ref": float((got - ref).abs().max() / max(1.0, float(ref.abs().max()))),
                "max_abs_err": float((got - ref).abs().max()), "ref_abs_max": float(ref.abs().max())}
    parity = {"vs_oracle_fp32_weights": _par(orc.encoder(x_chk, sd, DEPTH, HEADS)),
              "vs_oracle_bf16_rounded_weights": _par(orc.encoder(x_chk, w16, DEPTH, HEADS)),
              "gate": "rel_l2 <= 1e-2 and max_abs/max(1,max|ref|) <= 1e-2 (tests/test_gpu_configs.py)",
              "checked": "first 2 of the 256 sequences, all 12 layers, bf16 output of the fused driver vs the fp32 CPU "
                         "oracle on the same bf16 tokens; oracle weights: (a) the unrounded fp32 state-dict the modules "
                         "were loaded with, (b) the same rounded to bf16"}

    # ---- set-up outside any timing: one call per rotating input buffer so that every CUDA graph FusedEncoder keeps
    #      (keyed on the input / output pointers) is captured before the warm-up steps start, whatever W is ----
    with torch.no_grad():
        for i in range(NBUF):
            fused(xs[i], out=y)
    torch.cuda.synchronize()

    # ---- device-resident throughput ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)          # let nvidia-smi come up; it keeps sampling through warm-up and the timed region
    with torch.no_grad():
        for i in range(args.warmup):
            fused(xs[i % NBUF], out=y)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            fused(xs[i % NBUF], out=y)
        e1.record()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * tokens_per_step * args.steps / (ms_total * 1e-3)
    flops_step = synthetic.encoder_flops(B, N, D, DEPTH)

    # ---- end to end through the public API with pinned host buffers (3-slot software pipeline) ----
    SLOTS = 3
    h_in = [torch.randn(B, N, D, generator=g).bfloat16().pin_memory() for _ in range(SLOTS)]
    h_out = [torch.empty(B, N, D, dtype=torch.bfloat16).pin_memory() for _ in range(SLOTS)]
    d_in = [torch.empty(B, N, D, dtype=torch.bfloat16, device=dev) for _ in range(SLOTS)]
    d_out = [torch.empty(B, N, D, dtype=torch.bfloat16, device=dev) for _ in range(SLOTS)]
    s_h2d, s_d2h, s_cmp = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(SLOTS)]
    ev_cmp = [torch.cuda.Event() for _ in range(SLOTS)]
    ev_out = [torch.cuda.Event() for _ in range(SLOTS)]

    def e2e_steps(n):
        for i in range(n):
            s = i % SLOTS
            with torch.cuda.stream(s_h2d):
                s_h2d.wait_event(ev_cmp[s])          # slot's previous compute has consumed d_in[s]
                d_in[s].copy_(h_in[s], non_blocking=True)
                ev_in[s].record(s_h2d)
            s_cmp.wait_event(ev_in[s])
            s_cmp.wait_event(ev_out[s])              # slot's previous D2H has drained d_out[s]
            fused(d_in[s], out=d_out[s])
            ev_cmp[s].record(s_cmp)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(ev_cmp[s])
                h_out[s].copy_(d_out[s], non_blocking=True)
                ev_out[s].record(s_d2h)
        s_d2h.synchronize()

    with torch.no_grad():
        e2e_steps(SLOTS)                      # set-up: one pass per slot captures the slot's CUDA graph
        e2e_steps(min(args.warmup, 3))
        barrier()
        t0 = time.perf_counter()
        e2e_steps(args.steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * tokens_per_step * args.steps / float(t.item())
    bytes_io = B * N * D * 2

    # ---- rooflines: every kernel of one forward timed on the device inside >= 2 s of continuous CUDA-graph replays
    #      (%globaltimer at the first CTA's start / last CTA's exit: metatransformer_b200/timeline.py), clocks and
    #      power sampled during the probe itself ----
    roof, roof_others, step_timeline = None, None, None
    if rank == 0:
        from metatransformer_b200 import timeline
        probe = mt.FusedEncoder(enc, use_cuda_graph=False)
        settle = max(40, int(2.2 / (ms_total / args.steps * 1e-3)))
        psampler = ClockSampler(local_rank)
        psampler.start()
        time.sleep(0.2)
        res = timeline.record(probe, xs[0], y, settle=settle)
        pclocks = psampler.stop()
        kinds = res["summary"]["by_kind_us"]
        gemm_us = sum(kinds[k]["total"] for k in ("qkv", "proj", "fc1", "fc2"))
        n_gemm = sum(kinds[k]["launches"] for k in ("qkv", "proj", "fc1", "fc2"))
        M = B * N
        shape_flops = {"qkv": 2.0 * M * 3 * D * D, "proj": 2.0 * M * D * D, "fc1": 2.0 * M * 4 * D * D, "fc2": 2.0 * M * 4 * D * D}
        gemm_flops = DEPTH * sum(shape_flops.values())
        ach = gemm_flops / (gemm_us * 1e-6) / 1e12
        traffic = None
        tfile = ROOT / "profiles" / "r2_dram_traffic.json"
        if tfile.exists():
            traffic = json.loads(tfile.read_text())
        roof = {"kernel": "gemm_bf16_kernel<2, *> (qkv, proj, fc1+GELU, fc2; %d launches/step)" % n_gemm,
                "bound": "tensor", "achieved": ach, "peak": sustained_tf, "unit": "TFLOP/s", "frac": ach / sustained_tf,
                "frac_of_burst_peak": ach / burst_tf, "frac_of_nominal_2250": ach / 2250.0,
                "peak_source": f"{peak_src} MEASURED_PEAKS.json: bf16_tflops_sustained = {sustained_tf} (frac), bf16_tflops = "
                               f"{burst_tf} (frac_of_burst_peak)",
                "flops_per_launch": gemm_flops / n_gemm, "avg_launch_ms": gemm_us / n_gemm * 1e-3,
                "per_shape": {k: {"mean_us": kinds[k]["mean"], "tflops": shape_flops[k] / (kinds[k]["mean"] * 1e-6) / 1e12}
                              for k in ("qkv", "proj", "fc1", "fc2")},
                "traffic": None if traffic is None else traffic.get("gemm_mean_bytes_per_launch"),
                "traffic_source": None if traffic is None else traffic.get("source"),
                "probe": {"how": f"device timeline of one graph replay after {settle} back-to-back replays "
                                 f"({settle * ms_total / args.steps * 1e-3:.1f} s)", "clocks": pclocks},
                # what actually caps the main loop (profiles/r2_gemm_waits.md; constants measured with
                # tools/gemm_trace.py and tools/bench_tma.cu, not re-measured by this run)
                "limiter": {"what": "per-SM TMA operand ingest", "needed_B_per_clk_per_sm": 64.0,
                            "delivered_B_per_clk_per_sm": 43.5, "tma_microbench_B_per_clk_per_sm": 37.4,
                            "mma_issuer_operand_wait_share": {"qkv": 0.34, "fc1": 0.28, "proj": 0.30, "fc2": 0.27},
                            "source": "profiles/r2_gemm_waits.md"}}
        attn_bytes = 4.0 * M * D * 2          # read q, k, v once + write the output once, bf16
        attn_us = kinds["attention"]["mean"]
        roof_others = [{"kernel": "attention_fwd_kernel (%d launches/step)" % kinds["attention"]["launches"], "bound": "hbm",
                        "achieved": attn_bytes / (attn_us * 1e-6) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                        "frac": attn_bytes / (attn_us * 1e-6) / 1e9 / hbm_gbs, "avg_launch_ms": attn_us * 1e-3,
                        "algorithmic_bytes_per_launch": attn_bytes,
                        "traffic": None if traffic is None else traffic.get("attention_bytes_per_launch"),
                        "note": "also MUFU-bound: one ex2 per score, 16 / clk / SM puts the same ~47 us floor under it; "
                                "in-kernel wait accounting (profiles/r2_attention_trace.md): latency-bound dependency cycle, "
                                "no unit saturated"}]
        step_timeline = {k: res["summary"][k] for k in ("kernels", "span_us", "sum_kernel_us", "sum_gap_us", "mean_gap_us",
                                                        "max_gap_us")}
        step_timeline["by_kind_us"] = {k: round(v["total"], 2) for k, v in kinds.items()}
        # tokenizer (HBM-bound): Data2Seq('image') at the batch of the timed step, fp32 images -> bf16 tokens
        try:
            tok = mt.Data2Seq("image", D, token_dtype=torch.bfloat16).to(dev).eval()
            img = torch.rand(B, 3, 224, 224, device=dev)
            with torch.no_grad():
                for _ in range(3):
                    tok(img)
                t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0e.record()
                for _ in range(10):
                    tok(img)
                t1e.record()
                torch.cuda.synchronize()
            tok_ms = t0e.elapsed_time(t1e) / 10
            tok_bytes = img.numel() * 4 + B * N * D * 2 + (D * 768 + D) * 2
            roof_others.append({"kernel": "Data2Seq('image') tokenizer (patch gather + tcgen05 GEMM), B=%d" % B, "bound": "hbm",
                                "achieved": tok_bytes / (tok_ms * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                                "frac": tok_bytes / (tok_ms * 1e-3) / 1e9 / hbm_gbs, "avg_call_ms": tok_ms,
                                "algorithmic_bytes_per_call": tok_bytes, "traffic": None})
        except Exception as e:   # the tokenizer is not on the timed path: report, do not fail the bench
            roof_others.append({"kernel": "Data2Seq('image') tokenizer", "error": repr(e)})

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ----
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, _ = cpu_oracle_tokens_per_s(steps=50, warmup=1, budget_s=15.0)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "C2: Base-16 encoder forward (12 x Block, d=768, 12 heads), synthetic bf16 tokens "
                                   "(B,N,d)=(256,196,768) per GPU, seeded random weights",
                       "global_batch": B * world, "seq_len": N, "parallelism": f"dp{world} (replicas, no forward collective)",
                       "l2": "4 rotating 77 MB input buffers (308 MB > 126 MB L2) + ~1.3 GB intermediates per step",
                       "launch": "CUDA-graph replay of the per-forward launch sequence" if args.cuda_graph else "stream launches"},
            "tflops": flops_step * world * args.steps / (ms_total * 1e-3) / 1e12,
            "frac_of_sustained_peak": flops_step * args.steps / (ms_total * 1e-3) / 1e12 / sustained_tf,
            "frac_of_burst_peak": flops_step * args.steps / (ms_total * 1e-3) / 1e12 / burst_tf,
            "frac_of_nominal_2250": flops_step * args.steps / (ms_total * 1e-3) / 1e12 / 2250.0,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_io, "d2h_bytes_per_step": bytes_io,
                    "how": "pinned host tokens -> H2D -> FusedEncoder -> D2H of the (B,N,d) features, 3-slot pipeline"},
            "gpu_launches": fused.launches_per_forward * args.steps,
            "roofline": roof,
            "roofline_others": roof_others,
            "step_timeline": step_timeline,
            "cpu_baseline": cpu_base,
            "parity": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false",
                    help="launch the kernels of each forward individually instead of replaying a captured CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
