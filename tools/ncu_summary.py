#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) or a launch-list CSV (--metrics gpu__time_duration.sum) into the
small text summaries kept under profiles/.

    python tools/ncu_summary.py full   gpurun_out/x.ncu-rep  > profiles/rNN_x.md
    python tools/ncu_summary.py list   gpurun_out/launches.csv > profiles/rNN_launches.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard",
]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of `{path.split('/')[-1]}`\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"## {d.get('Kernel Name', '?')}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for h in hdr:
            base = h.split(".TriageCompute.")[-1]
            if any(base == k or base.startswith(k) for k in KEYS):
                if d[h] != "":
                    print(f"| {base} | {d[h]} | {u[h]} |")
        print()


def launch_list(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            ns = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list `{path.split('/')[-1]}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n")
    print(f"{sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms total\n")
    print("| kernel | launches | total us | share | avg us | grid | block |\n|---|---|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / tot * 100:.1f}% | {v[1] / v[0] / 1e3:.1f} | {v[2]} | {v[3]} |")


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2])
