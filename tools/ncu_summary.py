"""Markdown summary of one or more .ncu-rep captures (read offline with `ncu -i ... --page raw --csv`).

    python tools/ncu_summary.py gpurun_out/prof_fc2_r1.ncu-rep [more.ncu-rep ...] > profiles/xxx.md
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
]


def main(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            print(f"## {path}\n(no data)\n")
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        print(f"## {path}\n")
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            print(f"### `{name[:110]}`\n")
            print("| metric | value |")
            print("|---|---|")
            for k, label in KEYS:
                if k in idx:
                    print(f"| {label} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |")
            print()


if __name__ == "__main__":
    main(sys.argv[1:])
