#!/usr/bin/env python3
"""Generate tests/golden/sw_*.json by running the reference swaptions binaries (oracle/_ref/sw_ref_serial and
sw_ref_ff: the reference's HJM_Securities.cpp + HJM_Swaption_Blocking.cpp compiled unmodified from where they lie,
linked with the restated PARSEC leaves of oracle/sw_absent/ -- see oracle/Makefile).

The reference tree holds no golden vectors for swaptions, so the fixtures are outputs of the reference itself:

  sw_<case>.json = {"args": {ns, sm, sd}, "stdout_header": "...", "lines": [the "Swaption i: [...]" stderr lines]}

Every case is run through the serial build, the FastFlow build with 1, 3 and 8 workers and the SkePU-OpenMP build
(HJM_Securities_skepu_omp.cpp, sw_ref_skepu); all outputs must be identical (asserted here), so a golden is the answer
of every reference variant that can be built.

Cases:
  simsmall16      -ns 16 -sm 10000   (PARSEC simsmall; trials a multiple of BLOCK_SIZE)
  ragged7         -ns 7  -sm 1003 -sd 42   (trials not a multiple of 16: the last block simulates 1008; StdError is
                                            "-nan" where the extra terms make the variance estimate negative)
  single1         -ns 1  -sm 16
  two_trials      -ns 3  -sm 2  -sd 7      (fewer trials than one block)
  seeds5          -ns 5  -sm 4096 -sd 123456789
  medium32        -ns 32 -sm 20000   (PARSEC simmedium)
Run in the build container only (needs oracle/_ref):  python tests/golden/make_sw_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sw_oracle_lib as so  # noqa: E402

CASES = {
    "simsmall16": dict(ns=16, sm=10000, sd=None),
    "ragged7": dict(ns=7, sm=1003, sd=42),
    "single1": dict(ns=1, sm=16, sd=None),
    "two_trials": dict(ns=3, sm=2, sd=7),
    "seeds5": dict(ns=5, sm=4096, sd=123456789),
    "medium32": dict(ns=32, sm=20000, sd=None),
}


def main():
    for name, a in CASES.items():
        outs = []
        for binary, nt in (("sw_ref_serial", 1), ("sw_ref_ff", 1), ("sw_ref_ff", 3), ("sw_ref_ff", 8), ("sw_ref_skepu", 2)):
            if nt > a["ns"]:
                continue
            stdout, stderr, _ = so.run_ref(a["ns"], a["sm"], nt, a["sd"], binary)
            outs.append([l for l in stderr.splitlines() if l.startswith("Swaption")])
            header = [l for l in stdout.splitlines() if l.startswith("Number of Simulations")][0]
        assert all(o == outs[0] for o in outs), name
        assert len(outs[0]) == a["ns"]
        with open(os.path.join(HERE, "sw_%s.json" % name), "w") as f:
            json.dump({"args": a, "stdout_header": header, "variants_agreeing": len(outs), "lines": outs[0]}, f, indent=1)
        print(name, len(outs[0]), "lines,", len(outs), "variants agree")


if __name__ == "__main__":
    main()
