"""ctypes view of oracle/libbs_oracle.so -- the CPU restatement used as the parity checker.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  Nothing under p3arsec_b200/ may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
_LIB = None


def build():
    """(Re)build libbs_oracle.so and, when /root/reference is mounted, oracle/_ref/*."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(ORACLE_DIR, "libbs_oracle.so")
    if not os.path.exists(path):
        build()
    L = ctypes.CDLL(path)
    for sfx, c_fp in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
        p_fp = ctypes.POINTER(c_fp)
        p_int = ctypes.POINTER(ctypes.c_int)
        f = getattr(L, "bs_oracle_cndf_" + sfx)
        f.restype, f.argtypes = c_fp, [c_fp]
        f = getattr(L, "bs_oracle_price_" + sfx)
        f.restype, f.argtypes = c_fp, [c_fp] * 5 + [ctypes.c_int]
        f = getattr(L, "bs_oracle_map_" + sfx)
        f.restype, f.argtypes = None, [ctypes.c_size_t] + [p_fp] * 5 + [p_int, p_fp, ctypes.c_int]
        f = getattr(L, "bs_oracle_errchk_" + sfx)
        f.restype = ctypes.c_ulonglong
        f.argtypes = [ctypes.c_size_t, p_fp, p_fp, ctypes.POINTER(ctypes.c_longlong), ctypes.c_size_t]
        f = getattr(L, "bs_oracle_load_" + sfx)
        f.restype = ctypes.c_long
        f.argtypes = [ctypes.c_char_p, ctypes.c_size_t] + [p_fp] * 6 + [p_int, p_fp, p_fp]
        f = getattr(L, "bs_oracle_write_" + sfx)
        f.restype, f.argtypes = ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, p_fp]
    _LIB = L
    return L


def _np_fp(fp_bytes):
    return np.float32 if fp_bytes == 4 else np.float64


def _sfx(fp_bytes):
    return "f32" if fp_bytes == 4 else "f64"


def _ptr(a, ctype):
    return a.ctypes.data_as(ctypes.POINTER(ctype))


def price_map(spot, strike, rate, vol, tte, otype, fp_bytes=4, nthreads=None):
    """One pass of the reference Map over SoA arrays; returns the prices array."""
    dt = _np_fp(fp_bytes)
    c_fp = ctypes.c_float if fp_bytes == 4 else ctypes.c_double
    arrs = [np.ascontiguousarray(a, dtype=dt) for a in (spot, strike, rate, vol, tte)]
    ot = np.ascontiguousarray(otype, dtype=np.int32)
    n = arrs[0].shape[0]
    out = np.empty(n, dtype=dt)
    if nthreads is None:
        nthreads = len(os.sched_getaffinity(0))
    getattr(lib(), "bs_oracle_map_" + _sfx(fp_bytes))(
        n, *[_ptr(a, c_fp) for a in arrs], _ptr(ot, ctypes.c_int), _ptr(out, c_fp), nthreads)
    return out


def cndf(x, fp_bytes=4):
    return getattr(lib(), "bs_oracle_cndf_" + _sfx(fp_bytes))(float(x))


def errchk(prices, refval, fp_bytes=4, cap=1024):
    """ERR_CHK for one pass: (count, first `cap` offending indices)."""
    dt = _np_fp(fp_bytes)
    c_fp = ctypes.c_float if fp_bytes == 4 else ctypes.c_double
    p = np.ascontiguousarray(prices, dtype=dt)
    r = np.ascontiguousarray(refval, dtype=dt)
    idx = np.full(cap, -1, dtype=np.int64)
    cnt = getattr(lib(), "bs_oracle_errchk_" + _sfx(fp_bytes))(
        p.shape[0], _ptr(p, c_fp), _ptr(r, c_fp), _ptr(idx, ctypes.c_longlong), cap)
    return int(cnt), idx[: min(cnt, cap)].copy()


def load(path, fp_bytes=4):
    """The reference loader + AoS->SoA staging.  Returns a dict of SoA arrays (raises on error)."""
    dt = _np_fp(fp_bytes)
    c_fp = ctypes.c_float if fp_bytes == 4 else ctypes.c_double
    fn = getattr(lib(), "bs_oracle_load_" + _sfx(fp_bytes))
    nullp = ctypes.POINTER(c_fp)()
    n = fn(path.encode(), 0, nullp, nullp, nullp, nullp, nullp, nullp, ctypes.POINTER(ctypes.c_int)(), nullp, nullp)
    if n < 0:
        raise IOError("oracle loader: cannot read header of %s (code %d)" % (path, n))
    names = ("sptprice", "strike", "rate", "divq", "volatility", "otime")
    d = {k: np.empty(n, dtype=dt) for k in names}
    d["otype"] = np.empty(n, dtype=np.int32)
    d["divs"] = np.empty(n, dtype=dt)
    d["dgrefval"] = np.empty(n, dtype=dt)
    got = fn(path.encode(), n, *[_ptr(d[k], c_fp) for k in names], _ptr(d["otype"], ctypes.c_int),
             _ptr(d["divs"], c_fp), _ptr(d["dgrefval"], c_fp))
    if got != n:
        raise IOError("oracle loader: bad row in %s (code %d)" % (path, got))
    d["numOptions"] = int(n)
    return d


def write(path, prices, fp_bytes=4):
    dt = _np_fp(fp_bytes)
    c_fp = ctypes.c_float if fp_bytes == 4 else ctypes.c_double
    p = np.ascontiguousarray(prices, dtype=dt)
    rc = getattr(lib(), "bs_oracle_write_" + _sfx(fp_bytes))(path.encode(), p.shape[0], _ptr(p, c_fp))
    if rc != 0:
        raise IOError("oracle writer failed with %d" % rc)


def read_prices_text(path):
    """Parse a prices file ("%i\\n" + N x "%.18f\\n") into (n, list of the exact text tokens)."""
    with open(path) as f:
        toks = f.read().split()
    return int(toks[0]), toks[1:]


def ref_binary(name):
    """Path of a compiled reference binary under oracle/_ref, or None if it was never built."""
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


def run_ref(name, nthreads, infile, outfile, timeout=600):
    """Run an oracle/_ref binary the way parsecmgmt would; returns (stdout, roi_seconds or None)."""
    exe = ref_binary(name)
    if exe is None:
        raise FileNotFoundError("oracle/_ref/%s not built (run `make -C oracle`)" % name)
    cp = subprocess.run([exe, str(nthreads), infile, outfile], capture_output=True, text=True, timeout=timeout)
    roi = None
    for line in cp.stdout.splitlines():
        if line.startswith("roi.time|"):
            roi = float(line.split("|")[1])
    if cp.returncode != 0:
        raise RuntimeError("%s exited %d: %s" % (name, cp.returncode, cp.stdout[-500:]))
    return cp.stdout, roi
