"""Print the instructions with the most warp-stall samples from `ncu -i X.ncu-rep --page source --csv [--kernel-id :::N]`.

    ncu -i prof.ncu-rep --page source --csv --kernel-id :::3 > /tmp/src.csv; python tools/ncu_top_stalls.py /tmp/src.csv
"""
import csv
import sys


def main(path, n=30):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[h]
    i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    data = []
    for r in rows[h + 1:]:
        if len(r) <= i_s:
            continue
        try:
            data.append((int(r[i_s] or 0), r))
        except ValueError:
            pass
    tot = sum(s for s, _ in data)
    print("total samples", tot)
    for s, r in sorted(data, key=lambda t: -t[0])[:n]:
        reasons = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"{s:7d} {100.0 * s / max(tot, 1):5.1f}% ex={r[i_ex]:>9} {r[0]:>6} {r[i_src][:90]:<90} {reasons}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
