#!/usr/bin/env python3
"""Generate p3arsec_b200/data/optionData.txt -- the 1000-row base table of the synthetic input generator.

PARSEC's inputgen writes row i of an input file as row (i % 1000) of a fixed 1000-row table
(`optionData.txt`, `#include`d into inputgen.c as an array initialiser).  That table lives in the
PARSEC 3.0 tarball, not in the P3ARSEC overlay, so it is not available offline.
This script draws a stand-in table ONCE, with a fixed seed, from the distribution SURVEY.md 8(d)
specifies, in the same initialiser syntax:

    {s, strike, r, divq, v, t, 'C'|'P', divs, DGrefval},

* rows 0-3 are the Hull textbook rows of PARSEC's in_4.txt, DGrefval as recorded there;
* every other DGrefval is the fp64 oracle price of the row (printed %.18f), i.e. what DerivaGem's
  double-precision closed form would give, so the fp32 build sits inside the reference's 1e-4 ERR_CHK
  band on all rows.

Dev-time tool (lives under tests/ because it calls the oracle); the product only reads its output.
Run:  python tests/golden/make_option_table.py
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "p3arsec_b200", "data", "optionData.txt")

HULL = [
    (42.00, 40.00, 0.1000, 0.00, 0.20, 0.50, "C", 0.00, "4.759423036851750055"),
    (42.00, 40.00, 0.1000, 0.00, 0.20, 0.50, "P", 0.00, "0.808600016880314021"),
    (100.00, 100.00, 0.0500, 0.00, 0.15, 1.00, "P", 0.00, "3.714602051381290071"),
    (100.00, 100.00, 0.0500, 0.00, 0.15, 1.00, "C", 0.00, "8.591659601309890704"),
]


def main():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "libbs_oracle.so"))
    lib.bs_oracle_price_f64.restype = ctypes.c_double
    lib.bs_oracle_price_f64.argtypes = [ctypes.c_double] * 5 + [ctypes.c_int]

    rng = np.random.RandomState(1)
    rows = list(HULL)
    rates = [0.0250, 0.0275, 0.0500, 0.0750, 0.1000]
    while len(rows) < 1000:
        s = round(float(rng.uniform(20.0, 120.0)), 2)
        k = round(s * float(rng.uniform(0.7, 1.3)), 2)
        r = rates[int(rng.randint(0, len(rates)))]
        v = round(float(rng.uniform(0.05, 0.65)), 2)
        t = round(float(rng.uniform(0.05, 1.00)), 2)
        ty = "P" if rng.uniform() < 0.5 else "C"
        # price the row exactly as a reader of the 2/4-decimal text would see it
        s, k, v, t = (float("%.2f" % x) for x in (s, k, v, t))
        ref = lib.bs_oracle_price_f64(s, k, r, v, t, 1 if ty == "P" else 0)
        rows.append((s, k, r, 0.00, v, t, ty, 0.00, "%.18f" % ref))

    with open(OUT, "w") as f:
        for (s, k, r, q, v, t, ty, d, ref) in rows:
            f.write("{%.2f, %.2f, %.4f, %.2f, %.2f, %.2f, '%s', %.2f, %s},\n" % (s, k, r, q, v, t, ty, d, ref))
    print("wrote", OUT, len(rows), "rows")


if __name__ == "__main__":
    sys.exit(main())
