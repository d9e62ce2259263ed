"""bs_libm_f64.h (the exp/log of BS_MATH_REFERENCE with fptype=double) against the libm of this box, on the host.

The header restates glibc 2.39's double-precision exp/log (operation order and fma contractions of the __exp_fma /
__log_fma variants, tables read out of the image's libm by tools/gen_libm_f64_tables.py); compiled for the host it must
return the running libm's bits.  Default: 5 M draws per argument class (4e7 / 3.5e7 arguments per function, about a second);
BS_LIBM_EXHAUSTIVE=1 runs 260 M draws per class (2.1e9 / 1.8e9 arguments).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exp_log_equal_the_host_libm(tmp_path):
    exe = str(tmp_path / "libm64_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-mfma", "-I", os.path.join(ROOT, "p3arsec_b200", "csrc"),
                    os.path.join(ROOT, "tools", "libm_f64_host_check.cpp"), "-o", exe, "-lm"], check=True)
    millions = "260" if os.environ.get("BS_LIBM_EXHAUSTIVE") else "5"
    cp = subprocess.run([exe, millions], capture_output=True, text=True, timeout=1800)
    lines = dict((l.split()[0], l.split()[1:]) for l in cp.stdout.splitlines())
    assert cp.returncode == 0, cp.stdout
    assert lines["exp"][0] == "0" and lines["log"][0] == "0", cp.stdout
    assert int(lines["exp"][1]) >= 8 * int(millions) * 1_000_000


def test_tables_are_the_running_libms(tmp_path):
    # the committed table header is what tools/gen_libm_f64_tables.py extracts from this box's libm
    libm = "/lib/x86_64-linux-gnu/libm.so.6"
    if not os.path.exists(libm):
        import pytest
        pytest.skip("no libm.so.6 at the usual place")
    committed = open(os.path.join(ROOT, "p3arsec_b200", "csrc", "bs_libm_f64_tables.h")).read()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import importlib
    gen = importlib.import_module("gen_libm_f64_tables")
    saved_root, saved_argv = gen.ROOT, sys.argv
    try:
        os.makedirs(tmp_path / "p3arsec_b200" / "csrc")
        gen.ROOT, sys.argv = str(tmp_path), ["gen", libm]
        gen.main()
    finally:
        gen.ROOT, sys.argv = saved_root, saved_argv
    assert open(tmp_path / "p3arsec_b200" / "csrc" / "bs_libm_f64_tables.h").read() == committed
