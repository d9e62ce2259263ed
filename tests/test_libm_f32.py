"""bs_libm_f32.h (the expf/logf of BS_MATH_REFERENCE) against the libm of this box, on the host.

The header restates glibc 2.39's expf/logf; compiled for the host it must return the running libm's bits for every
float.  Default stride 16 (2.7e8 arguments per function, a few seconds); BS_LIBM_EXHAUSTIVE=1 checks all 2^32.
"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_expf_logf_equal_the_host_libm(tmp_path):
    exe = str(tmp_path / "libm_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-mfma", "-I", os.path.join(ROOT, "p3arsec_b200", "csrc"),
                    os.path.join(ROOT, "tools", "libm_f32_host_check.cpp"), "-o", exe, "-lm"], check=True)
    stride = "1" if os.environ.get("BS_LIBM_EXHAUSTIVE") else "16"
    cp = subprocess.run([exe, stride], capture_output=True, text=True, timeout=900)
    lines = dict((l.split()[0], l.split()[1:]) for l in cp.stdout.splitlines())
    assert cp.returncode == 0, cp.stdout
    assert lines["expf"][0] == "0" and lines["logf"][0] == "0", cp.stdout
    assert int(lines["expf"][1]) >= 2**32 // int(stride)
