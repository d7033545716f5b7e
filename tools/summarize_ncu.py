#!/usr/bin/env python
"""Summarise ncu artefacts into small text files under profiles/ (the .ncu-rep files stay in gpurun_out/).

    python tools/summarize_ncu.py launches <launches.csv> <out.md>
    python tools/summarize_ncu.py full <report.ncu-rep> <out.md>
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__cluster_dim_x", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
]


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0][-70:]
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as out:
        out.write(f"# launch list summary of `{src}` (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n")
        out.write("cold-cache, serialised launches: compare SHARES, not absolutes\n\n")
        out.write("| launches | total us | avg us | share | kernel |\n|---:|---:|---:|---:|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write(f"| {v[0]} | {v[1] / 1e3:.1f} | {v[1] / v[0] / 1e3:.1f} | {100 * v[1] / tot:.1f}% | `{k}` |\n")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as out:
        out.write(f"# ncu --set full summary of `{src}`\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            out.write(f"## {d.get('Kernel Name', '?')[:120]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d:
                    out.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            out.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
