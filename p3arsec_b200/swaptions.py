"""ctypes binding of libsw_gpu.so (include/sw_gpu.h) and a host-side mirror of the reference swaptions driver.

Reference surface mirrored (parsec-ff/pkgs/apps/swaptions/src/HJM_Securities.cpp): `make_portfolio()` is the
set-up loop at :198,:231-297 (RanUnif-driven dYears / dStrike, the yield curve, the factor table); `SwaptionsGPU.price()`
is the Map at :311-323 and fills what the reference stores in parm.dSimSwaptionMeanPrice / dSimSwaptionStdError.
The shared library is the product; there is no CPU fallback: a missing library raises ImportError and a box without
a CUDA device makes sw_gpu_init fail.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsw_gpu.so")
ABI_VERSION = 1

FLAG_IEEE, FLAG_LEAN, FLAG_BATCHED = 1, 2, 4
MAX_N, MAX_FACTORS = 32, 8
BLOCK_SIZE = 16              # PARSEC HJM_type.h, passed at HJM_Securities.cpp:319
DEFAULT_NUM_TRIALS = 102400  # PARSEC HJM_type.h, HJM_Securities.cpp:53
DEFAULT_SEED = 1979          # HJM_Securities.cpp:61
IN, IFACTORS = 11, 3         # HJM_Securities.cpp:56,58

# every symbol include/sw_gpu.h declares
ABI_SYMBOLS = (
    "sw_gpu_abi_version", "sw_gpu_status_string", "sw_gpu_device_count", "sw_gpu_init", "sw_gpu_init_devices", "sw_gpu_price",
    "sw_gpu_set_geometry", "sw_gpu_get_timing", "sw_gpu_num_shards", "sw_gpu_shard", "sw_gpu_last_error", "sw_gpu_fini",
)

# the scalar parm fields in the order of the call at HJM_Securities.cpp:314-318 (struct sw_gpu_swaption)
SWAPTION_DTYPE = np.dtype([("dStrike", "f8"), ("dCompounding", "f8"), ("dMaturity", "f8"), ("dTenor", "f8"),
                           ("dPaymentInterval", "f8"), ("dYears", "f8")])

# HJM_Securities.cpp:231-265
FACTOR_TABLE = np.array([
    [.01, .01, .01, .01, .01, .01, .01, .01, .01, .01],
    [.009048, .008187, .007408, .006703, .006065, .005488, .004966, .004493, .004066, .003679],
    [.001000, .000750, .000500, .000250, .000000, -.000250, -.000500, -.000750, -.001000, -.001250]], dtype=np.float64)


class Timing(ctypes.Structure):
    _fields_ = [("roi_ms", ctypes.c_double), ("wall_ms", ctypes.c_double), ("kernel_launches", ctypes.c_ulonglong),
                ("trials_simulated", ctypes.c_ulonglong), ("h2d_bytes", ctypes.c_ulonglong), ("d2h_bytes", ctypes.c_ulonglong)]


class SwGpuError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__("%s failed: %s (%d)%s" % (where, _status_string(status), status, (": " + detail) if detail else ""))


_lib = None


def load_library(path=None):
    """dlopen libsw_gpu.so and declare the prototypes.  Raises if it was not built -- no fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("SW_GPU_LIB") or LIB_PATH  # SW_GPU_LIB: an alternative build of the same ABI (tuning runs)
    if not os.path.exists(p):
        raise ImportError("%s not found: build it with `make -C p3arsec_b200/csrc` (or __graft_entry__.build()); "
                          "p3arsec_b200 has no CPU fallback" % p)
    L = ctypes.CDLL(p)
    vp, ci, cl, pd = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.POINTER(ctypes.c_double)
    L.sw_gpu_abi_version.restype, L.sw_gpu_abi_version.argtypes = ci, []
    L.sw_gpu_status_string.restype, L.sw_gpu_status_string.argtypes = ctypes.c_char_p, [ci]
    L.sw_gpu_device_count.restype, L.sw_gpu_device_count.argtypes = ci, []
    L.sw_gpu_init.restype, L.sw_gpu_init.argtypes = ci, [ctypes.POINTER(vp), ci, ci, ci, ci]
    L.sw_gpu_init_devices.restype, L.sw_gpu_init_devices.argtypes = ci, [ctypes.POINTER(vp), ctypes.POINTER(ci), ci, ci, ci, ci]
    L.sw_gpu_price.restype, L.sw_gpu_price.argtypes = ci, [vp, ci, vp, pd, pd, cl, cl, ci, ctypes.c_uint, pd, pd]
    L.sw_gpu_set_geometry.restype, L.sw_gpu_set_geometry.argtypes = ci, [vp, ci, ci]
    L.sw_gpu_get_timing.restype, L.sw_gpu_get_timing.argtypes = ci, [vp, ctypes.POINTER(Timing)]
    L.sw_gpu_num_shards.restype, L.sw_gpu_num_shards.argtypes = ci, [vp]
    L.sw_gpu_shard.restype, L.sw_gpu_shard.argtypes = ci, [vp, ci, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]
    L.sw_gpu_last_error.restype, L.sw_gpu_last_error.argtypes = ctypes.c_char_p, [vp]
    L.sw_gpu_fini.restype, L.sw_gpu_fini.argtypes = None, [vp]
    if L.sw_gpu_abi_version() != ABI_VERSION:
        raise ImportError("libsw_gpu.so ABI %d != binding ABI %d" % (L.sw_gpu_abi_version(), ABI_VERSION))
    if path is None:
        _lib = L
    return L


def _status_string(status):
    try:
        return load_library().sw_gpu_status_string(status).decode()
    except Exception:
        return "status"


def device_count():
    return load_library().sw_gpu_device_count()


# ---- the driver's set-up, restated (pure integer / double arithmetic; no GPU involved) -------------------------------
def ran_unif(state):
    """PARSEC RanUnif on a one-element list holding the counter (HJM_Securities.cpp:198,279,281 call it with &seed).
    Python integers reproduce the C `long` arithmetic for non-negative counters (the product stays below 2^63)."""
    ix = state[0]
    state[0] = ix + 1
    ix = (ix * 1513517) % 2147483647
    k1 = ix // 127773
    ix = 16807 * (ix - k1 * 127773) - k1 * 2836
    if ix < 0:
        ix += 2147483647
    return ix * 4.656612875e-10


def make_portfolio(n_swaptions, seed=DEFAULT_SEED):
    """HJM_Securities.cpp:198 and :276-296: returns (swaption_seed, swaptions[SWAPTION_DTYPE], yields[n,11], factors[n,3,10])."""
    if seed < 0:
        raise ValueError("make_portfolio mirrors the C arithmetic for non-negative seeds only")
    st = [int(seed)]
    swaption_seed = int(2147483647 * ran_unif(st))                       # :198
    sw = np.zeros(n_swaptions, dtype=SWAPTION_DTYPE)
    yields = np.empty((n_swaptions, IN), dtype=np.float64)
    for i in range(n_swaptions):
        sw["dYears"][i] = 5.0 + int(60 * ran_unif(st)) * 0.25              # :279
        sw["dStrike"][i] = 0.1 + int(49 * ran_unif(st)) * 0.1              # :281
    sw["dCompounding"] = 0
    sw["dMaturity"] = 1.0
    sw["dTenor"] = 2.0
    sw["dPaymentInterval"] = 1.0
    y = .1                                                                 # :288-290 (running sum, as the reference does)
    yields[:, 0] = y
    for j in range(1, IN):
        y = y + .005
        yields[:, j] = y
    factors = np.ascontiguousarray(np.broadcast_to(FACTOR_TABLE, (n_swaptions,) + FACTOR_TABLE.shape))
    return swaption_seed, sw, yields, factors


class SwaptionsGPU:
    """One sw_gpu_ctx: `num_gpus` devices, a fixed HJM path shape and a capacity in swaptions."""

    def __init__(self, max_swaptions, num_gpus=1, iN=IN, iFactors=IFACTORS, devices=None):
        self._L = load_library()
        self._ctx = ctypes.c_void_p()
        self.max_swaptions, self.iN, self.iFactors = int(max_swaptions), int(iN), int(iFactors)
        if devices is not None:
            arr = (ctypes.c_int * len(devices))(*devices)
            st = self._L.sw_gpu_init_devices(ctypes.byref(self._ctx), arr, len(devices), self.max_swaptions, self.iN, self.iFactors)
        else:
            st = self._L.sw_gpu_init(ctypes.byref(self._ctx), int(num_gpus), self.max_swaptions, self.iN, self.iFactors)
        if st != 0:
            self._ctx = ctypes.c_void_p()
            raise SwGpuError(st, "sw_gpu_init")

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._L.sw_gpu_fini(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st, where):
        if st != 0:
            raise SwGpuError(st, where, self._L.sw_gpu_last_error(self._ctx).decode())

    def set_geometry(self, ctas_per_sm=0, trials_per_thread=0):
        self._check(self._L.sw_gpu_set_geometry(self._ctx, ctas_per_sm, trials_per_thread), "sw_gpu_set_geometry")

    def price(self, swaptions, yields, factors, swaption_seed, trials=DEFAULT_NUM_TRIALS, block_size=BLOCK_SIZE, flags=0):
        """The Map (HJM_Securities.cpp:311-323).  Returns (mean, std_error), one entry per swaption."""
        sw = np.ascontiguousarray(swaptions, dtype=SWAPTION_DTYPE)
        n = sw.shape[0]
        y = np.ascontiguousarray(yields, dtype=np.float64).reshape(n, self.iN)
        f = np.ascontiguousarray(factors, dtype=np.float64).reshape(n, self.iFactors, self.iN - 1)
        mean = np.empty(n, dtype=np.float64)
        err = np.empty(n, dtype=np.float64)
        pd = ctypes.POINTER(ctypes.c_double)
        st = self._L.sw_gpu_price(self._ctx, n, sw.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(pd), f.ctypes.data_as(pd),
                                  int(swaption_seed), int(trials), int(block_size), int(flags), mean.ctypes.data_as(pd),
                                  err.ctypes.data_as(pd))
        self._check(st, "sw_gpu_price")
        return mean, err

    def timing(self):
        t = Timing()
        self._check(self._L.sw_gpu_get_timing(self._ctx, ctypes.byref(t)), "sw_gpu_get_timing")
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def shards(self):
        out = []
        for g in range(self._L.sw_gpu_num_shards(self._ctx)):
            d, f, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            self._check(self._L.sw_gpu_shard(self._ctx, g, ctypes.byref(d), ctypes.byref(f), ctypes.byref(c)), "sw_gpu_shard")
            out.append((d.value, f.value, c.value))
        return out
