"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) per kernel:
count, total / mean duration and share of the captured time.   python tools/summarize_launches.py X.csv [steps]"""
import csv
import re
import sys
from collections import OrderedDict


def main(path, steps=None):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= i_val:
            continue
        name = re.sub(r"\(.*", "", r[i_name]).replace("void ", "").strip()
        ns = float(r[i_val].replace(",", ""))
        if r[i_unit] in ("us", "usecond"):
            ns *= 1e3
        elif r[i_unit] in ("ms", "msecond"):
            ns *= 1e6
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"captured launches: {n}, total device time {tot / 1e6:.3f} ms" + (f" ({tot / 1e6 / steps:.3f} ms/step over {steps} steps)" if steps else ""))
    print("| kernel | launches | total ms | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {t / 1e6:.3f} | {t / c / 1e3:.1f} | {100 * t / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
