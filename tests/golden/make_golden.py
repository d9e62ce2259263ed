#!/usr/bin/env python3
"""Generate tests/golden/* by running the UNMODIFIED reference binaries (oracle/_ref/*).

The reference tree has no golden vectors for blackscholes, so the fixtures are
outputs of the reference itself, compiled from /root/reference by oracle/Makefile:

  <name>.in.txt            input in the reference grammar
  <name>.ref_f32.txt       prices written by bs_ref_serial          (fptype=float, the shipped build)
  <name>.ref_f64.txt       prices written by bs_ref_serial_fp64     (fptype=double substitution)
  <name>.errchk.json       "Num Errors" and the "Error on ..." lines of the -DERR_CHK builds

The FastFlow, OpenMP and SkePU builds are run too and must be byte-identical to the serial output
(asserted here, so every golden is the answer of all four reference variants).

Inputs:
  hull4      PARSEC in_4.txt: the four Hull textbook rows
  table1k    the 1000-row base table (one full period of the inputgen cycle)
  ragged37   37 rows: not a multiple of 4, 32 or any tile size
  single1    one option
  edge2k     2048 seeded rows with full-precision fields that stress the formula: deep ITM/OTM,
             vol 0.01..1.5, t 0.003..10y, rates 0..0.25, type chars beyond P/C (anything != 'P' is
             a call, blackscholes.c:761), and DGrefval deliberately off for ~3% of rows so ERR_CHK
             fires.
Run in the build container only (needs oracle/_ref):  python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402


def table_rows():
    rows = []
    with open(os.path.join(ROOT, "p3arsec_b200", "data", "optionData.txt")) as f:
        for line in f:
            line = line.strip().strip(",").strip("{}")
            if not line:
                continue
            p = [x.strip() for x in line.split(",")]
            rows.append("%s %s %s %s %s %s %s %s %s" % (p[0], p[1], p[2], p[3], p[4], p[5], p[6].strip("'"), p[7], p[8]))
    return rows


def edge_rows(n=2048, seed=20261017):
    rng = np.random.RandomState(seed)
    lib = oracle_lib.lib()
    rows = []
    kinds = "PCPCPCXcp"
    for i in range(n):
        s = float(np.float32(rng.uniform(1.0, 500.0)))
        mode = i % 8
        if mode == 0:
            k = s * float(rng.uniform(0.2, 0.6))       # deep in/out of the money
        elif mode == 1:
            k = s * float(rng.uniform(1.6, 4.0))
        elif mode == 2:
            k = s * float(rng.uniform(0.999, 1.001))    # at the money: log(s/k) ~ 0
        else:
            k = s * float(rng.uniform(0.7, 1.3))
        r = float(rng.choice([0.0, 0.001, 0.0250, 0.05, 0.1, 0.25]))
        v = float(rng.choice([0.01, 0.05, 0.2, 0.65, 1.0, 1.5])) if mode in (3, 4) else float(rng.uniform(0.05, 0.65))
        t = float(rng.choice([0.003, 0.01, 0.0833, 1.0, 5.0, 10.0])) if mode in (5, 6) else float(rng.uniform(0.05, 1.0))
        ty = kinds[int(rng.randint(0, len(kinds)))]
        # what the loader will see after text round trip
        s6, k6, r6, v6, t6 = (float("%.6f" % x) for x in (s, k, r, v, t))
        ref = lib.bs_oracle_price_f64(s6, k6, r6, v6, t6, 1 if ty == "P" else 0)
        if rng.uniform() < 0.03:
            ref += float(rng.choice([-1.0, 1.0])) * float(rng.uniform(1.5e-4, 0.5))
        rows.append("%.6f %.6f %.6f 0.00 %.6f %.6f %s 0.00 %.18f" % (s, k, r, v, t, ty, ref))
    return rows


def write_input(path, rows):
    with open(path, "w") as f:
        f.write("%d\n" % len(rows))
        for r in rows:
            f.write(r + "\n")


def run(exe, nthreads, inp, out):
    cp = subprocess.run([os.path.join(oracle_lib.REF_DIR, exe), str(nthreads), inp, out],
                        capture_output=True, text=True, check=True)
    return cp.stdout


def main():
    oracle_lib.build()
    tab = table_rows()
    cases = {
        "hull4": tab[:4],
        "table1k": tab,
        "ragged37": tab[100:137],
        "single1": tab[7:8],
        "edge2k": edge_rows(),
    }
    for name, rows in cases.items():
        inp = os.path.join(HERE, name + ".in.txt")
        write_input(inp, rows)
        out32 = os.path.join(HERE, name + ".ref_f32.txt")
        out64 = os.path.join(HERE, name + ".ref_f64.txt")
        run("bs_ref_serial", 1, inp, out32)
        run("bs_ref_serial_fp64", 1, inp, out64)
        ref32 = open(out32).read()
        nthr = min(4, len(rows))
        for exe in ("bs_ref_ff", "bs_ref_omp", "bs_ref_skepu"):
            tmp = "/tmp/_golden_%s_%s.txt" % (name, exe)
            run(exe, nthr, inp, tmp)
            assert open(tmp).read() == ref32, "%s disagrees with serial on %s" % (exe, name)
            os.unlink(tmp)
        tmp = "/tmp/_golden_ff64.txt"
        run("bs_ref_ff_fp64", nthr, inp, tmp)
        assert open(tmp).read() == open(out64).read()
        os.unlink(tmp)
        chk = {}
        for key, exe in (("f32", "bs_ref_serial_errchk"), ("f64", "bs_ref_serial_fp64_errchk")):
            so = run(exe, 1, inp, "/tmp/_golden_chk.txt")
            lines = so.splitlines()
            errs = [l for l in lines if l.startswith("Error on ")]
            num = [l for l in lines if l.startswith("Num Errors:")][0]
            # the reference repeats every error line NUM_RUNS (=100) times; keep one run's worth
            per_run = errs[: len(errs) // 100] if errs else []
            chk[key] = {"num_errors_line": num, "num_runs": 100, "errors_one_run": per_run}
        with open(os.path.join(HERE, name + ".errchk.json"), "w") as f:
            json.dump(chk, f, indent=1)
        print(name, len(rows), "rows;", chk["f32"]["num_errors_line"], "/", chk["f64"]["num_errors_line"])


if __name__ == "__main__":
    main()
