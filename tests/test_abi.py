"""The drop-in boundary on a box WITHOUT a GPU: the library loads, exports every symbol the headers
declare, validates arguments, and refuses to run (no CPU fallback).  No compute call is made."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from p3arsec_b200 import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bs_(?:gpu|io)_[a-z_0-9]+)\s*\(", src)))


def test_headers_and_binding_agree():
    assert _declared("bs_gpu.h") == sorted(host.ABI_SYMBOLS)
    assert _declared("bs_io.h") == sorted(host.IO_SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = host.load_library()
    for name in list(host.ABI_SYMBOLS) + list(host.IO_SYMBOLS):
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", host.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (bs_(?:gpu|io)_\w+)", out))
    assert set(host.ABI_SYMBOLS) | set(host.IO_SYMBOLS) <= exported
    assert L.bs_gpu_abi_version() == host.ABI_VERSION


def test_library_is_sm100a_cuda_code():
    out = subprocess.run(["cuobjdump", "-lelf", host.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_status_strings():
    L = host.load_library()
    assert L.bs_gpu_status_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert L.bs_gpu_status_string(code) not in (b"ok", b"unknown status")


def _cfg(**kw):
    cfg = host.Config()
    cfg.struct_size = ctypes.sizeof(host.Config)
    cfg.num_options, cfg.fp_bytes, cfg.num_gpus = 16, 4, 1
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


@pytest.mark.parametrize("kw", [dict(fp_bytes=2), dict(num_gpus=0), dict(math=7), dict(unroll=3), dict(threads_per_block=100),
                                dict(blocks_per_sm=99), dict(struct_size=8), dict(num_options=2**31)])
def test_init_rejects_bad_arguments_before_touching_a_device(kw):
    L = host.load_library()
    ctx = ctypes.c_void_p()
    assert L.bs_gpu_init_ex(ctypes.byref(ctx), ctypes.byref(_cfg(**kw))) == -1  # BS_GPU_ERR_INVALID
    assert not ctx.value
    assert L.bs_gpu_init_ex(None, ctypes.byref(_cfg())) == -1


def test_null_context_calls_are_safe():
    L = host.load_library()
    assert L.bs_gpu_price(None, 1, 0, None) == -1
    assert L.bs_gpu_price_aos(None, None, 0, None, 1) == -1
    assert L.bs_gpu_upload(None) == -1 and L.bs_gpu_run(None, 1, 0, None) == -1 and L.bs_gpu_download(None) == -1
    assert L.bs_gpu_host_buffer(None, 0) is None
    assert L.bs_gpu_num_shards(None) == -1
    L.bs_gpu_fini(None)


def test_no_cpu_fallback_without_a_gpu():
    if host.device_count() > 0:
        pytest.skip("a GPU is present")
    assert host.device_count() == 0
    with pytest.raises(host.BsGpuError) as ei:
        host.BlackScholesGPU(16)
    assert ei.value.status == -2  # BS_GPU_ERR_NO_DEVICE
    exe = os.path.join(os.path.dirname(host.LIB_PATH), "..", "bin", "blackscholes_gpu")
    cp = subprocess.run([exe, "1", os.path.join(ROOT, "tests", "golden", "hull4.in.txt"), "/tmp/_never.txt"],
                        capture_output=True, text=True)
    assert cp.returncode == 1 and "no usable CUDA device" in cp.stdout
    # argument handling happens before any device is touched, exactly as in the reference driver
    cp = subprocess.run([exe], capture_output=True, text=True)
    assert cp.returncode == 1 and "Usage:" in cp.stdout
    cp = subprocess.run([exe, "1", "/nonexistent/in.txt", "/tmp/_never.txt"], capture_output=True, text=True)
    assert cp.returncode == 1 and "ERROR: Unable to open file `/nonexistent/in.txt'." in cp.stdout
    cp = subprocess.run([exe, "9", os.path.join(ROOT, "tests", "golden", "hull4.in.txt"), "/tmp/_never.txt"], capture_output=True, text=True)
    assert "WARNING: Not enough work, reducing number of threads to match number of options." in cp.stdout


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(ImportError):
        host.load_library(str(tmp_path / "libbs_gpu.so"))


def test_product_never_touches_the_oracle():
    # oracle/ is test infrastructure: nothing under p3arsec_b200/ or include/ may reference it
    for base in ("p3arsec_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".c", ".h", "Makefile")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "oracle_lib" not in text and "libbs_oracle" not in text and "bs_oracle_" not in text, f
    out = subprocess.run(["ldd", host.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_bytes_per_option():
    assert host.bytes_per_option(4) == 28 and host.bytes_per_option(8) == 52
    assert host.bytes_per_option(4, err_chk=True) == 32


C_PROBE = r"""
#include <stdio.h>
#include "bs_gpu.h"
#include "bs_io.h"
int main(void)
{
    bs_gpu_config cfg = {0};
    bs_gpu_ctx *ctx = NULL;
    bs_gpu_timing tm;
    (void)tm;
    cfg.struct_size = sizeof(cfg);
    cfg.num_options = 16; cfg.fp_bytes = 3; cfg.num_gpus = 1;     /* invalid fptype: rejected before any device */
    printf("%d %d %s %d\n", bs_gpu_abi_version(), bs_gpu_init_ex(&ctx, &cfg), bs_gpu_status_string(BS_GPU_ERR_NO_DEVICE),
           (int)BS_BUF_COUNT);
    return ctx != NULL;
}
"""


def test_headers_are_plain_c_and_link(tmp_path):
    # the boundary is a C ABI: both headers must compile as C99 with -pedantic and link against the library
    src = tmp_path / "probe.c"
    src.write_text(C_PROBE)
    exe = str(tmp_path / "probe")
    lib_dir = os.path.dirname(host.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe,
                    "-L", lib_dir, "-lbs_gpu", "-Wl,-rpath," + lib_dir], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == [str(host.ABI_VERSION), "-1", "no", "usable", "CUDA", "device", "8"]


def test_integration_patches_apply_to_the_reference(tmp_path):
    # the patches a P3ARSEC maintainer would apply: both must apply cleanly to the reference as it lies
    ref = "/root/reference/parsec-ff/pkgs/apps/blackscholes/src/blackscholes.c"
    if not os.path.exists(ref):
        pytest.skip("/root/reference is not mounted on this box")
    integ = os.path.join(ROOT, "integration")
    step1 = str(tmp_path / "bs_cuda.c")
    subprocess.run(["patch", "-s", "-o", step1, ref, os.path.join(integ, "blackscholes.c.enable_cuda.patch")], check=True)
    step2 = str(tmp_path / "bs_cuda_caf.c")
    subprocess.run(["patch", "-s", "-o", step2, step1, os.path.join(integ, "blackscholes.c.caf_cuda.patch")], check=True)
    text = open(step2).read()
    assert "bs_gpu_price_aos(gpu, data_vec.data(), nv, res.data(), 1)" in text and "bs_gpu_price(gpu, NUM_RUNS" in text


def test_limit_devices_edits_cuda_visible_devices():
    # bs_gpu_limit_devices only rewrites the environment (it must run before the first CUDA call); checked in a child process
    code = ("import os, ctypes; L = ctypes.CDLL(%r); "
            "os.environ.pop('CUDA_VISIBLE_DEVICES', None); assert L.bs_gpu_limit_devices(3) == 0; a = os.environ.get('CUDA_VISIBLE_DEVICES'); "
            "import ctypes.util; libc = ctypes.CDLL(None); libc.getenv.restype = ctypes.c_char_p; a = libc.getenv(b'CUDA_VISIBLE_DEVICES'); "
            "libc.setenv(b'CUDA_VISIBLE_DEVICES', b'GPU-aa,5,2,7', 1); assert L.bs_gpu_limit_devices(2) == 0; b = libc.getenv(b'CUDA_VISIBLE_DEVICES'); "
            "assert L.bs_gpu_limit_devices(9) == 0; c = libc.getenv(b'CUDA_VISIBLE_DEVICES'); assert L.bs_gpu_limit_devices(0) == -1; "
            "print(a.decode(), b.decode(), c.decode())") % host.LIB_PATH
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["0,1,2", "GPU-aa,5", "GPU-aa,5"]
