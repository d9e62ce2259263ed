"""ctypes view of oracle/libsw_oracle.so -- the CPU restatement of the swaptions Map used as the parity checker.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the CPU-baseline legs of the bench tools.
Nothing under p3arsec_b200/ may import this module.
"""
import ctypes
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
_LIB = None

SWAPTION_DTYPE = np.dtype([("dStrike", "f8"), ("dCompounding", "f8"), ("dMaturity", "f8"), ("dTenor", "f8"),
                           ("dPaymentInterval", "f8"), ("dYears", "f8")])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(ORACLE_DIR, "libsw_oracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)
    L = ctypes.CDLL(path)
    pd, vp, ci, cl = ctypes.POINTER(ctypes.c_double), ctypes.c_void_p, ctypes.c_int, ctypes.c_long
    L.sw_oracle_price_one.restype, L.sw_oracle_price_one.argtypes = ci, [pd, vp, ci, ci, pd, pd, cl, cl, ci]
    L.sw_oracle_map.restype, L.sw_oracle_map.argtypes = ci, [ci, vp, ci, ci, pd, pd, cl, cl, ci, pd, pd, ci]
    L.sw_oracle_portfolio.restype, L.sw_oracle_portfolio.argtypes = cl, [ci, cl, vp, pd, pd]
    L.sw_oracle_ranunif.restype, L.sw_oracle_ranunif.argtypes = ctypes.c_double, [ctypes.POINTER(cl)]
    L.sw_oracle_cumnormalinv.restype, L.sw_oracle_cumnormalinv.argtypes = ctypes.c_double, [ctypes.c_double]
    _LIB = L
    return L


def _pd(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def portfolio(n, seed=1979):
    """HJM_Securities.cpp:198,276-296 through the oracle: (swaption_seed, swaptions, yields[n,11], factors[n,3,10])."""
    sw = np.zeros(n, dtype=SWAPTION_DTYPE)
    y = np.empty((n, 11), dtype=np.float64)
    f = np.empty((n, 3, 10), dtype=np.float64)
    s = lib().sw_oracle_portfolio(n, seed, sw.ctypes.data_as(ctypes.c_void_p), _pd(y), _pd(f))
    return int(s), sw, y, f


def price_map(swaptions, yields, factors, swaption_seed, trials, block_size=16, iN=11, iFactors=3, nthreads=None):
    """The Map over swaptions (HJM_Securities.cpp:311-323) on host cores; returns (mean, std_error)."""
    sw = np.ascontiguousarray(swaptions, dtype=SWAPTION_DTYPE)
    n = sw.shape[0]
    y = np.ascontiguousarray(yields, dtype=np.float64).reshape(n, iN)
    f = np.ascontiguousarray(factors, dtype=np.float64).reshape(n, iFactors, iN - 1)
    mean = np.zeros(n, dtype=np.float64)
    err = np.zeros(n, dtype=np.float64)
    ok = lib().sw_oracle_map(n, sw.ctypes.data_as(ctypes.c_void_p), iN, iFactors, _pd(y), _pd(f), int(swaption_seed), int(trials),
                             int(block_size), _pd(mean), _pd(err), nthreads or (os.cpu_count() or 1))
    if ok != 1:
        raise ValueError("sw_oracle_map: a swaption's time indices run off the HJM path")
    return mean, err


def ran_unif(counter):
    c = ctypes.c_long(counter)
    return lib().sw_oracle_ranunif(ctypes.byref(c)), c.value


def cum_normal_inv(u):
    return lib().sw_oracle_cumnormalinv(float(u))


# ---- the compiled reference (oracle/_ref/sw_ref_*: HJM_Securities.cpp + HJM_Swaption_Blocking.cpp, unmodified) ------
_LINE = re.compile(r"Swaption (\d+): \[SwaptionPrice: (\S+) StdError: (\S+)\]")


def have_ref(name="sw_ref_serial"):
    return os.path.exists(os.path.join(REF_DIR, name))


def parse_ref_output(text):
    """The reference's result lines (HJM_Securities.cpp:357-358) -> list of (id, price string, stderr string)."""
    return [(int(m.group(1)), m.group(2), m.group(3)) for m in _LINE.finditer(text)]


def run_ref(ns, sm, nt=1, seed=None, name="sw_ref_serial"):
    """Run the compiled reference; returns (stdout, stderr, roi seconds)."""
    cmd = [os.path.join(REF_DIR, name), "-ns", str(ns), "-sm", str(sm), "-nt", str(nt)]
    if seed is not None:
        cmd += ["-sd", str(seed)]
    p = subprocess.run(cmd, capture_output=True, text=True, check=True)
    m = re.search(r"roi\.time\|([0-9.eE+-]+)", p.stdout)
    return p.stdout, p.stderr, float(m.group(1)) if m else None


def format_lines(mean, err):
    """What the reference prints per swaption (HJM_Securities.cpp:357-358)."""
    def c_fmt(v):  # C printf prints the sign of a NaN ("-nan" is what x86 sqrt of a negative number yields)
        if np.isnan(v):
            return "-nan" if np.signbit(v) else "nan"
        return "%.10f" % v
    return ["Swaption %d: [SwaptionPrice: %s StdError: %s] " % (i, c_fmt(m), c_fmt(e)) for i, (m, e) in enumerate(zip(mean, err))]
