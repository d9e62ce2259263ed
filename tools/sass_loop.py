#!/usr/bin/env python3
"""Static instruction mix of one kernel of a built library (cuobjdump -sass), whole function and the longest loop.

usage: tools/sass_loop.py <lib.so> <mangled-name substring> [--dump out.txt]
The loop is the backward branch spanning the most instructions; CALL set-up blocks inside it are listed separately
(they belong to the out-of-line degenerate-input path and are not executed for in-range options)."""
import collections
import re
import subprocess
import sys


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    i = sass.index(pat, sass.index("Function : ") if "Function : " in sass else 0)
    i = sass.rfind("Function :", 0, i)
    j = sass.find("Function :", i + 10)
    body = sass[i:j if j > 0 else len(sass)]
    ins = []
    for l in body.splitlines():
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    if "--dump" in sys.argv:
        open(sys.argv[sys.argv.index("--dump") + 1], "w").write("\n".join("%04x %s" % a for a in ins))
    addr = {a: n for n, (a, _) in enumerate(ins)}
    best = (0, 0, 0)
    for n, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)* (?:P\d, )?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] < n:
            span = n - addr[int(m.group(1), 16)] + 1
            if span > best[0]:
                best = (span, addr[int(m.group(1), 16)], n)

    def mix(seq):
        c = collections.Counter(re.sub(r"^@!?U?P\d+ ", "", t).split()[0].split(".")[0] for _, t in seq)
        return dict(c.most_common())

    print(body.splitlines()[0].strip())
    print("function: %d instructions" % len(ins), mix(ins))
    span, lo, hi = best
    loop = ins[lo:hi + 1]
    calls = sum(1 for _, t in loop if t.startswith("CALL"))
    print("longest loop: %d instructions (%04x..%04x), %d CALL sites" % (span, ins[lo][0], ins[hi][0], calls), mix(loop))
    fp64 = sum(1 for _, t in loop if re.match(r"(@!?P\d+ )?D(FMA|MUL|ADD|SETP)", t))
    print("  FP64-pipe instructions in the loop: %d" % fp64)


if __name__ == "__main__":
    main()
