"""The swaptions drop-in boundary (include/sw_gpu.h, libsw_gpu.so) on a box WITHOUT a GPU: the library loads, exports
every symbol the header declares, validates arguments, and refuses to run (no CPU fallback).  No compute call."""
import ctypes
import os
import re
import subprocess

import pytest

from p3arsec_b200 import swaptions as sw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "p3arsec_b200", "bin", "swaptions_gpu")


def _declared():
    src = open(os.path.join(ROOT, "include", "sw_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sw_gpu_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(sw.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = sw.load_library()
    for name in sw.ABI_SYMBOLS:
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", sw.LIB_PATH], capture_output=True, text=True).stdout
    assert set(sw.ABI_SYMBOLS) <= set(re.findall(r" T (sw_gpu_\w+)", out))
    assert L.sw_gpu_abi_version() == sw.ABI_VERSION


def test_library_is_sm100a_cuda_code():
    out = subprocess.run(["cuobjdump", "-lelf", sw.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layouts_match_the_header():
    assert sw.SWAPTION_DTYPE.itemsize == 6 * 8           # struct sw_gpu_swaption: six doubles
    assert ctypes.sizeof(sw.Timing) == 2 * 8 + 4 * 8     # struct sw_gpu_timing


def test_status_strings():
    L = sw.load_library()
    assert L.sw_gpu_status_string(0) == b"ok"
    for code in (-1, -2, -3, -4):
        assert L.sw_gpu_status_string(code) not in (b"ok", b"unknown status")


@pytest.mark.parametrize("args", [(0, 4, 11, 3), (1, 0, 11, 3), (1, 4, 1, 3), (1, 4, 33, 3), (1, 4, 11, 0), (1, 4, 11, 9)])
def test_init_rejects_bad_arguments_before_touching_a_device(args):
    L = sw.load_library()
    ctx = ctypes.c_void_p()
    assert L.sw_gpu_init(ctypes.byref(ctx), *args) == -1
    assert not ctx.value
    assert L.sw_gpu_init(None, 1, 4, 11, 3) == -1


def test_null_context_calls_are_safe():
    L = sw.load_library()
    assert L.sw_gpu_price(None, 0, None, None, None, 0, 0, 16, 0, None, None) == -1
    assert L.sw_gpu_set_geometry(None, 0, 0) == -1 and L.sw_gpu_get_timing(None, None) == -1
    assert L.sw_gpu_num_shards(None) == -1 and L.sw_gpu_shard(None, 0, None, None, None) == -1
    L.sw_gpu_fini(None)


def test_no_cpu_fallback_without_a_gpu():
    if sw.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(sw.SwGpuError) as ei:
        sw.SwaptionsGPU(4)
    assert ei.value.status == -2  # SW_GPU_ERR_NO_DEVICE
    cp = subprocess.run([EXE, "-ns", "2", "-sm", "32"], capture_output=True, text=True)
    assert cp.returncode == 1 and "ERROR: sw_gpu_init: no CUDA device" in cp.stderr
    assert "Swaption 0" not in cp.stderr


def test_driver_argument_handling_matches_the_reference():
    # HJM_Securities.cpp:139-147,171-190: all of this happens before any device is touched
    cp = subprocess.run([EXE], capture_output=True, text=True)
    assert cp.returncode == 1 and cp.stdout.startswith("PARSEC Benchmark Suite") and "Usage:" in cp.stderr
    assert "\t-ns [number of swaptions (should be > number of threads]\n" in cp.stderr
    cp = subprocess.run([EXE, "-zz"], capture_output=True, text=True)
    assert cp.returncode == 1 and "Error: Unknown option: -zz" in cp.stderr
    cp = subprocess.run([EXE, "-ns", "2", "-nt", "4"], capture_output=True, text=True)
    assert cp.returncode == 1 and "Error: Fewer swaptions than threads." in cp.stderr
    cp = subprocess.run([EXE, "-ns", "3", "-sm", "64", "-nt", "2"], capture_output=True, text=True)
    assert "Number of Simulations: 64,  Number of threads: 2 Number of swaptions: 3" in cp.stdout


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(ImportError):
        sw.load_library(str(tmp_path / "libsw_gpu.so"))


def test_product_does_not_link_the_checker():
    out = subprocess.run(["ldd", sw.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    text = open(os.path.join(ROOT, "p3arsec_b200", "swaptions.py")).read()
    assert "sw_oracle" not in text and "oracle_lib" not in text


C_PROBE = r"""
#include <stdio.h>
#include "sw_gpu.h"
int main(void)
{
    sw_gpu_ctx *ctx = NULL;
    sw_gpu_timing tm;
    sw_gpu_swaption s = {0};
    (void)tm; (void)s;
    printf("%d %d %s %d\n", sw_gpu_abi_version(), sw_gpu_init(&ctx, 1, 4, 99, 3), sw_gpu_status_string(SW_GPU_ERR_NO_DEVICE),
           (int)sizeof(sw_gpu_swaption));
    return ctx != NULL;
}
"""


def test_header_is_plain_c_and_links(tmp_path):
    src = tmp_path / "probe.c"
    src.write_text(C_PROBE)
    exe = tmp_path / "probe"
    libdir = os.path.dirname(sw.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lsw_gpu", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0
    assert out.stdout.split() == ["1", "-1", "no", "CUDA", "device", "48"]
