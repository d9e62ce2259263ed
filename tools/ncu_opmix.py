#!/usr/bin/env python3
"""Summarise an ncu report of one kernel launch: per-opcode instruction counts (per unit of work), pipe utilisation and
warp-stall ratios.  Usage: python tools/ncu_opmix.py report.ncu-rep <units in the launch> [top]"""
import collections
import csv
import io
import re
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, units = sys.argv[1], float(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 22
    raw = page(rep, "raw")
    hdr, vals = raw[0], raw[2]
    g = {h: v for h, v in zip(hdr, vals)}
    print("kernel:", g.get("Kernel Name"), " grid", g.get("launch__grid_size"), " regs", g.get("launch__registers_per_thread"),
          " time_ms", g.get("gpu__time_duration.sum"))
    for k in ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
              "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"):
        print("  %-70s %s" % (k, g.get(k)))
    stalls = [(float(v), h.split("issue_stalled_")[1].split("_per_issue")[0]) for h, v in g.items()
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    print("  stalls per issue:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:7]))
    src = page(rep, "source")
    h = src[1]
    ci = {x: i for i, x in enumerate(h)}
    ops, samp = collections.Counter(), collections.Counter()
    fp64 = 0
    for r in src[2:]:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]].strip())
        op = m.group(2) if m else "?"
        base = op.split(".")[0]
        if base in ("MUFU", "I2F", "F2I", "F2F"):
            base = ".".join(op.split(".")[:2])
        n = int(r[ci["Instructions Executed"]])
        ops[base] += n
        samp[base] += int(r[ci["# Samples"]])
        if base in ("DFMA", "DMUL", "DADD", "DSETP"):
            fp64 += n
    tot, tots = sum(ops.values()), max(1, sum(samp.values()))
    print("  warp instructions %d = %.1f thread-instructions per unit; FP64 pipe (DFMA+DMUL+DADD+DSETP) %.1f per unit" % (tot, tot * 32 / units, fp64 * 32 / units))
    for k, v in ops.most_common(top):
        print("    %-12s %6.2f%% inst %6.2f%% samples  %7.1f per unit" % (k, 100 * v / tot, 100 * samp[k] / tots, v * 32 / units))


if __name__ == "__main__":
    main()
