#!/usr/bin/env python3
"""BASELINE.json configs[4]: end-to-end native run (10M options, NUM_RUNS=1) INCLUDING file parse, SoA staging,
pinned H2D/D2H and the prices file, through the drop-in driver, on G GPUs of this box.

Whole-process wall time, fopen to final fclose.  Beside it: the unmodified reference FastFlow binary on a bounded
sample (2M rows; NUM_RUNS=100 is compiled in, so its ROI is reported separately from its load+write time).
Prints one JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
BIN = os.path.join(ROOT, "p3arsec_b200", "bin")


def run(cmd):
    t0 = time.perf_counter()
    cp = subprocess.run(cmd, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if cp.returncode != 0:
        raise SystemExit("failed: %s\n%s" % (" ".join(cmd), cp.stdout[-2000:] + cp.stderr[-2000:]))
    return dt, cp.stdout


def kv(stdout, prefix):
    out = {}
    for line in stdout.splitlines():
        if line.startswith(prefix):
            for tok in line[len(prefix):].replace("(", " ").replace(")", " ").split():
                if "=" in tok:
                    k, v = tok.split("=", 1)
                    try:
                        out[k] = float(v)
                    except ValueError:
                        out[k] = v
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--ref-sample", type=int, default=2_000_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--dir", default="/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
    ap.add_argument("--devices", default=None, help="value for BS_GPU_DEVICES (all = <nthreads> GPUs literally; unset = chosen from the work)")
    ap.add_argument("--fast-exit", default=None, help="value for BS_GPU_FAST_EXIT (0 = ordinary exit through the CUDA runtime's teardown)")
    ap.add_argument("--ref-full", action="store_true", help="also run the reference on the FULL file (tens of seconds)")
    a = ap.parse_args()
    import numpy as np
    import oracle_lib

    d = tempfile.mkdtemp(prefix="bs_e2e_", dir=a.dir)
    inp, out = os.path.join(d, "in.txt"), os.path.join(d, "prices.txt")
    run([os.path.join(BIN, "bs_inputgen"), str(a.n), inp])
    best = None
    if a.devices is not None:
        os.environ["BS_GPU_DEVICES"] = a.devices
    if a.fast_exit is not None:
        os.environ["BS_GPU_FAST_EXIT"] = a.fast_exit
    for _ in range(a.reps):
        wall, so = run([os.path.join(BIN, "blackscholes_gpu_runs1"), str(a.gpus), inp, out])
        info = dict(kv(so, "[BS_GPU] "), wall_s=wall)
        if best is None or wall < best["wall_s"]:
            best = info
    res = {"config": "end-to-end native %d options, fp32, NUM_RUNS=1, %d GPU(s), incl. file parse/SoA staging/H2D/D2H/prices file" % (a.n, a.gpus),
           "input_bytes": os.path.getsize(inp), "output_bytes": os.path.getsize(out), "host_cores": len(os.sched_getaffinity(0)),
           "ours": best, "ours_options_per_s_whole_process": a.n / best["wall_s"],
           "env": {"BS_GPU_DEVICES": os.environ.get("BS_GPU_DEVICES"), "BS_GPU_FAST_EXIT": os.environ.get("BS_GPU_FAST_EXIT")}}
    if a.ref_full and oracle_lib.ref_binary("bs_ref_ff"):
        # the reference on the SAME file (its NUM_RUNS=100 is compiled in: the ROI is scaled to one run, load + write are what they are)
        cores = len(os.sched_getaffinity(0))
        rout = os.path.join(d, "ref_prices.txt")
        t0 = time.perf_counter()
        so, roi = oracle_lib.run_ref("bs_ref_ff", cores, inp, rout, timeout=3600)
        wall = time.perf_counter() - t0
        res["reference_full_file"] = {"rows": a.n, "wall_s_with_100_runs": wall, "roi_s_100_runs": roi, "load_plus_write_s": wall - roi,
                                      "NUM_RUNS_1_equivalent_s": wall - roi + roi / 100, "cores": cores}
        os.unlink(rout)

    # reference beside it, bounded sample, and parity of the two prices files on that sample
    sinp, sout, gout = os.path.join(d, "s_in.txt"), os.path.join(d, "s_ref.txt"), os.path.join(d, "s_gpu.txt")
    run([os.path.join(BIN, "bs_inputgen"), str(a.ref_sample), sinp])
    if oracle_lib.ref_binary("bs_ref_ff"):
        cores = len(os.sched_getaffinity(0))
        t0 = time.perf_counter()
        so, roi = oracle_lib.run_ref("bs_ref_ff", cores, sinp, sout, timeout=1800)
        wall = time.perf_counter() - t0
        run([os.path.join(BIN, "blackscholes_gpu_runs1"), str(a.gpus), sinp, gout])
        ref = np.loadtxt(sout, skiprows=1)
        got = np.loadtxt(gout, skiprows=1)
        res["reference_sample"] = {"rows": a.ref_sample, "wall_s": wall, "roi_s_100_runs": roi, "load_plus_write_s": wall - roi,
                                   "load_plus_write_us_per_row": (wall - roi) / a.ref_sample * 1e6, "cores": cores,
                                   "extrapolated_10M_NUM_RUNS_1_s": (wall - roi) / a.ref_sample * a.n + roi / 100 * a.n / a.ref_sample}
        res["parity_max_abs_delta_vs_reference_file"] = float(np.abs(ref - got).max())
    for p in (inp, out, sinp, sout, gout):
        if os.path.exists(p):
            os.unlink(p)
    os.rmdir(d)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
