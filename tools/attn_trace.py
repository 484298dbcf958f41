"""Event trace of the attention forward kernel (needs a -DMT_ATTN_TRACE=1 build, see tools/build_variant.py):

    python tools/build_variant.py trace mt_attention.cu -DMT_ATTN_TRACE=1
    MT_B200_LIB=metatransformer_b200/lib/variants/libmt_b200_trace.so python tools/attn_trace.py --config c2

prints, for CTA 0, the clock64 time line of the softmax warps (per half tile: wait for S, tcgen05.ld, wait for P.V,
exponentials, hand-over) and of the loader / issuer threads.
"""
import argparse
import ctypes
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import metatransformer_b200 as mt  # noqa: E402
from metatransformer_b200 import _lib  # noqa: E402

CONFIGS = {"c2": (256, 196, 12), "c3": (128, 256, 16), "c4": (64, 512, 12), "c5": (32, 1568, 16)}
NAMES = {1: "wait_s", 2: "got_s", 3: "ld_done", 4: "stats", 5: "pv_ok", 6: "pv_ok_grow", 7: "exp_start", 8: "exp_done",
         9: "handed", 10: "epi_wait", 11: "epi_start", 12: "epi_done", 20: "k_full", 21: "q_full", 22: "qk_issue_t0",
         23: "qk_issue_t1", 30: "pv_issue_t0", 31: "pv_issue_t1", 40: "q_req_t0", 41: "q_req_t1", 42: "k_req", 43: "v_req"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--out", default="gpurun_out/attn_trace.json")
    ap.add_argument("--items", type=int, default=3, help="print this many steady-state items of CTA 0")
    a = ap.parse_args()
    B, N, H = CONFIGS[a.config]
    lib = _lib.load()
    qkv = (torch.randn(B, N, 3 * H * 64, device="cuda") * 1.5).bfloat16()
    for _ in range(3):
        mt.ops.attention(qkv, H)
    buf = torch.zeros(2 * 12 * 1024, dtype=torch.int64, device="cuda")
    rc = lib.mt_attention_set_trace(ctypes.c_void_p(buf.data_ptr()))
    if rc == 1:
        raise SystemExit("this build has no trace points: build with -DMT_ATTN_TRACE=1 (or 2) and set MT_B200_LIB")
    mt.ops.attention(qkv, H)
    torch.cuda.synchronize()
    lib.mt_attention_set_trace(ctypes.c_void_p(0))
    if rc == -2:
        # wait-time accounting build: cycles inside barrier waits per warp, by kind
        w = buf.cpu()[:4 * 12 * 8].reshape(4, 12, 8)
        kinds = {"softmax": ["s_full", "pv_done (P slot)", "pv_done (epilogue)", "store read"],
                 8: ["q_empty", "k_empty", "v_empty", "-"], 9: ["k_full", "q_full", "s_empty h0", "s_empty h1"],
                 10: ["v_full", "p_full", "-", "-"]}
        for blk in range(4):
            for wp in range(11):
                names = kinds["softmax"] if wp < 8 else kinds[wp]
                tot = int(w[blk, wp, 4])
                print(f"CTA {blk} warp {wp:2d}: kernel {tot:7d} clk | " + " | ".join(
                    f"{n} {int(w[blk, wp, i]):6d} ({100.0 * int(w[blk, wp, i]) / max(tot, 1):4.1f}%)" for i, n in enumerate(names) if n != "-"))
        return
    ev = buf.cpu().reshape(2, 12, 1024)
    out = {}
    for blk in range(2):
        for w in range(12):
            e = ev[blk, w]
            e = e[e != 0]
            out[f"{blk}/{w}"] = [(int(x) >> 8, int(x) & 255) for x in e.tolist()]
    Path(a.out).parent.mkdir(exist_ok=True)
    Path(a.out).write_text(json.dumps(out))
    t0 = min(v[0][0] for v in out.values() if v)
    # softmax warp 0 (tile 0) and warp 4 (tile 1) of CTA 0: per-phase durations
    for w in (0, 4, 1):
        evs = out[f"0/{w}"]
        print(f"--- CTA 0 warp {w}: {len(evs)} events")
        items, cur = [], []
        for tck, eid in evs:
            cur.append((tck - t0, eid))
            if eid == 12:
                items.append(cur)
                cur = []
        for it in items[1:1 + a.items]:
            line, prev = [], it[0][0]
            print(f" item starts at {it[0][0]} clk, lasts {it[-1][0] - it[0][0]} clk")
            for tck, eid in it:
                line.append(f"{NAMES.get(eid, eid)}+{tck - prev}")
                prev = tck
                if eid in (9, 12):
                    print("   " + " ".join(line))
                    line = []
    for w, nm in ((8, "loader"), (9, "QK issuer"), (10, "PV issuer")):
        evs = out[f"0/{w}"]
        print(f"--- CTA 0 {nm}: first 60 events (clk since kernel start)")
        print("   " + " ".join(f"{NAMES.get(e, e)}@{t - t0}" for t, e in evs[:60]))


if __name__ == "__main__":
    main()
