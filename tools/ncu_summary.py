#!/usr/bin/env python3
"""Condense an .ncu-rep (or the launch-list CSV) into the few numbers the roofline discussion needs.

    python tools/ncu_summary.py report gpurun_out/prof.ncu-rep  > profiles/rNN_<name>.txt
    python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
]


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("# %s : %d profiled launch(es); ncu --set full --clock-control none; values per launch" % (path.split("/")[-1], len(data)))
    print("kernel: %s" % data[0][hdr.index("Kernel Name")])
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            print("%-72s %-16s %s" % (k, units[i], "  ".join(r[i] for r in data)))
    rd, wr, t = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    for r in data:
        tot = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
        us = float(r[t]) * {"us": 1.0, "ns": 1e-3, "ms": 1e3}[units[t]]
        print("dram traffic per launch = %.1f MB (read %.1f + write %.1f) in %.2f us -> %.0f GB/s at the DRAM pins" % (
            tot / 1e6, float(r[rd]) * scale[units[rd]] / 1e6, float(r[wr]) * scale[units[wr]] / 1e6, us, tot / us / 1e3))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0, 1e30, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg[r[ki]]
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
    tot = sum(v[1] for v in agg.values())
    print("# %s : gpu__time_duration.sum per launch (ncu --clock-control none; serialised, cold-cache: compare SHARES)" % path.split("/")[-1])
    print("%8s %14s %8s %10s %10s %10s  kernel" % ("launches", "total_" + rows[1][ui], "share%", "avg", "min", "max"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%8d %14.0f %8.2f %10.1f %10.1f %10.1f  %s" % (v[0], v[1], 100 * v[1] / tot, v[1] / v[0], v[2], v[3], k))


if __name__ == "__main__":
    {"report": report, "launches": launches}[sys.argv[1]](sys.argv[2])
