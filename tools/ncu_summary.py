#!/usr/bin/env python
"""Turns ncu output into the small markdown summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv   > profiles/rNN_launches.md
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [-k regex]  > profiles/rNN_<kernel>_full.md

`launches` reads the CSV of `ncu --metrics gpu__time_duration.sum --csv`; `full` reads a --set full report
through `ncu -i ... --page raw --csv` (needs ncu on PATH, no GPU)."""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_wait",
]


def launches(path):
    text = open(path).read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, r.get("Grid Size", ""), r.get("Block Size", "")])
        a[0] += 1
        a[1] += us
        total += us
    n = sum(a[0] for a in agg.values())
    print(f"# ncu launch list `{path.split('/')[-1]}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n")
    print(f"{n} launches, {total / 1e3:.3f} ms total\n")
    print("| kernel | launches | total us | share | avg us | grid | block |\n|---|---|---|---|---|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {name} | {a[0]} | {a[1]:.1f} | {100 * a[1] / total:.1f}% | {a[1] / a[0]:.1f} | {a[2]} | {a[3]} |")


def full(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of `{path.split('/')[-1]}`\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pattern and not re.search(pattern, name):
            continue
        print(f"## {name[:150]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        pat = sys.argv[sys.argv.index("-k") + 1] if "-k" in sys.argv else None
        full(sys.argv[2], pat)
