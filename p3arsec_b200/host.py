"""ctypes binding of libbs_gpu.so (include/bs_gpu.h) and a thin host-side mirror of the reference driver.

The shared library is the product; this module only lets Python (tests, bench.py) call the same C ABI
the C++ driver `blackscholes_gpu` calls.  There is no CPU fallback anywhere: if the library is missing
or no CUDA device is usable, construction raises.

Reference surface mirrored (parsec-ff/pkgs/apps/blackscholes/src/blackscholes.c): the SoA globals
sptprice/strike/rate/volatility/otime/otype/prices (:102-111) become numpy views of the context's pinned
host buffers; `run(num_runs)` is the ROI (:781-914) -- NUM_RUNS repetitions of the Map (:318).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbs_gpu.so")
ABI_VERSION = 2

# enum bs_gpu_buffer
BUF_SPTPRICE, BUF_STRIKE, BUF_RATE, BUF_VOLATILITY, BUF_OTIME, BUF_OTYPE, BUF_PRICES, BUF_DGREFVAL = range(8)
BUF_NAMES = ("sptprice", "strike", "rate", "volatility", "otime", "otype", "prices", "dgrefval")
# enum bs_gpu_math
MATH_DEFAULT, MATH_IEEE, MATH_FAST, MATH_REFERENCE = 0, 1, 2, 3
# flags
FLAG_NO_HOST_STAGING, FLAG_WITH_DGREFVAL, FLAG_NO_GRAPH, FLAG_ASYNC_DISCOVERY, FLAG_PDL, FLAG_NO_SUBSHARDS = 1, 2, 4, 8, 16, 32

NUM_RUNS = 100  # blackscholes.c:87

# struct DataCont of the CAF Map (blackscholes.c:482-489): the record layout bs_gpu_price_aos() takes
DATACONT_DTYPE = np.dtype([("otype", np.int32), ("sptprice", np.float32), ("strike", np.float32), ("rate", np.float32),
                           ("volatility", np.float32), ("otime", np.float32)])
assert DATACONT_DTYPE.itemsize == 24

# every symbol include/bs_gpu.h declares (tests check the library exports exactly these)
ABI_SYMBOLS = (
    "bs_gpu_abi_version", "bs_gpu_status_string", "bs_gpu_device_count", "bs_gpu_limit_devices", "bs_gpu_init", "bs_gpu_init_ex",
    "bs_gpu_host_buffer", "bs_gpu_mark_dirty", "bs_gpu_price", "bs_gpu_upload", "bs_gpu_run", "bs_gpu_download",
    "bs_gpu_price_aos", "bs_gpu_fill_synthetic", "bs_gpu_read_device", "bs_gpu_errors", "bs_gpu_num_shards", "bs_gpu_shard",
    "bs_gpu_get_timing", "bs_gpu_get_launch", "bs_gpu_last_error", "bs_gpu_fini",
)


class Config(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_size_t),
        ("num_options", ctypes.c_size_t),
        ("fp_bytes", ctypes.c_int),
        ("num_gpus", ctypes.c_int),
        ("devices", ctypes.POINTER(ctypes.c_int)),
        ("flags", ctypes.c_uint),
        ("math", ctypes.c_int),
        ("threads_per_block", ctypes.c_int),
        ("blocks_per_sm", ctypes.c_int),
        ("unroll", ctypes.c_int),
        ("variant", ctypes.c_int),
    ]


class Timing(ctypes.Structure):
    _fields_ = [
        ("h2d_ms", ctypes.c_double),
        ("roi_ms", ctypes.c_double),
        ("d2h_ms", ctypes.c_double),
        ("wall_ms", ctypes.c_double),
        ("pipeline_ms", ctypes.c_double),
        ("kernel_launches", ctypes.c_ulonglong),
        ("h2d_bytes", ctypes.c_ulonglong),
        ("d2h_bytes", ctypes.c_ulonglong),
    ]


class BsGpuError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__("%s failed: %s (%d)%s" % (where, _status_string(status), status, (": " + detail) if detail else ""))


_lib = None


def load_library(path=None):
    """dlopen libbs_gpu.so and declare the prototypes.  Raises if it was not built -- no fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("BS_GPU_LIB") or LIB_PATH  # BS_GPU_LIB: an alternative build of the same ABI (measurement runs only)
    if not os.path.exists(p):
        raise ImportError("%s not found: build it with `make -C p3arsec_b200/csrc` (or __graft_entry__.build()); "
                          "p3arsec_b200 has no CPU fallback" % p)
    L = ctypes.CDLL(p)
    vp, ci, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    pull = ctypes.POINTER(ctypes.c_ulonglong)
    L.bs_gpu_abi_version.restype, L.bs_gpu_abi_version.argtypes = ci, []
    L.bs_gpu_status_string.restype, L.bs_gpu_status_string.argtypes = ctypes.c_char_p, [ci]
    L.bs_gpu_device_count.restype, L.bs_gpu_device_count.argtypes = ci, []
    L.bs_gpu_limit_devices.restype, L.bs_gpu_limit_devices.argtypes = ci, [ci]
    L.bs_gpu_init.restype, L.bs_gpu_init.argtypes = ci, [ctypes.POINTER(vp), ci, cs, ci]
    L.bs_gpu_init_ex.restype, L.bs_gpu_init_ex.argtypes = ci, [ctypes.POINTER(vp), ctypes.POINTER(Config)]
    L.bs_gpu_host_buffer.restype, L.bs_gpu_host_buffer.argtypes = vp, [vp, ci]
    L.bs_gpu_mark_dirty.restype, L.bs_gpu_mark_dirty.argtypes = ci, [vp]
    L.bs_gpu_price.restype, L.bs_gpu_price.argtypes = ci, [vp, ci, ci, pull]
    L.bs_gpu_upload.restype, L.bs_gpu_upload.argtypes = ci, [vp]
    L.bs_gpu_run.restype, L.bs_gpu_run.argtypes = ci, [vp, ci, ci, pull]
    L.bs_gpu_download.restype, L.bs_gpu_download.argtypes = ci, [vp]
    L.bs_gpu_price_aos.restype, L.bs_gpu_price_aos.argtypes = ci, [vp, vp, cs, vp, ci]
    L.bs_gpu_fill_synthetic.restype, L.bs_gpu_fill_synthetic.argtypes = ci, [vp, ctypes.c_ulonglong]
    L.bs_gpu_read_device.restype, L.bs_gpu_read_device.argtypes = ci, [vp, ci, cs, cs, vp]
    L.bs_gpu_errors.restype, L.bs_gpu_errors.argtypes = ctypes.c_longlong, [vp, ctypes.POINTER(ctypes.c_longlong), cs]
    L.bs_gpu_num_shards.restype, L.bs_gpu_num_shards.argtypes = ci, [vp]
    L.bs_gpu_shard.restype, L.bs_gpu_shard.argtypes = ci, [vp, ci, ctypes.POINTER(ci), ctypes.POINTER(cs), ctypes.POINTER(cs)]
    L.bs_gpu_get_timing.restype, L.bs_gpu_get_timing.argtypes = ci, [vp, ctypes.POINTER(Timing)]
    L.bs_gpu_get_launch.restype, L.bs_gpu_get_launch.argtypes = ci, [vp, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]
    L.bs_gpu_last_error.restype, L.bs_gpu_last_error.argtypes = ctypes.c_char_p, [vp]
    L.bs_gpu_fini.restype, L.bs_gpu_fini.argtypes = None, [vp]
    if L.bs_gpu_abi_version() != ABI_VERSION:
        raise ImportError("libbs_gpu.so ABI %d != binding ABI %d" % (L.bs_gpu_abi_version(), ABI_VERSION))
    if path is None:
        _lib = L
    return L


def _status_string(status):
    try:
        return load_library().bs_gpu_status_string(status).decode()
    except Exception:  # library itself missing
        return "status"


def device_count():
    return load_library().bs_gpu_device_count()


class BlackScholesGPU:
    """One bs_gpu_ctx.  `devices` is a list of CUDA ordinals (one contiguous shard each)."""

    def __init__(self, num_options, fp_bytes=4, devices=None, num_gpus=None, math=MATH_DEFAULT, host_staging=True,
                 with_dgrefval=True, use_graph=True, threads_per_block=0, blocks_per_sm=0, unroll=0, variant=0, async_discovery=False, pdl=False, subshards=True):
        self._L = load_library()
        self._ctx = ctypes.c_void_p()
        if devices is None:
            devices = list(range(num_gpus or 1))
        self.devices = list(devices)
        self.n = int(num_options)
        self.fp_bytes = int(fp_bytes)
        self.dtype = np.float32 if fp_bytes == 4 else np.float64
        self.host_staging = host_staging
        dev_arr = (ctypes.c_int * len(self.devices))(*self.devices)
        cfg = Config()
        cfg.struct_size = ctypes.sizeof(Config)
        cfg.num_options = self.n
        cfg.fp_bytes = self.fp_bytes
        cfg.num_gpus = len(self.devices)
        cfg.devices = dev_arr
        cfg.flags = (0 if host_staging else FLAG_NO_HOST_STAGING) | (FLAG_WITH_DGREFVAL if with_dgrefval else 0) | \
                    (0 if use_graph else FLAG_NO_GRAPH) | (FLAG_ASYNC_DISCOVERY if async_discovery else 0) | (FLAG_PDL if pdl else 0) | (0 if subshards else FLAG_NO_SUBSHARDS)
        cfg.math = math
        cfg.threads_per_block = threads_per_block
        cfg.blocks_per_sm = blocks_per_sm
        cfg.unroll = unroll
        cfg.variant = variant
        st = self._L.bs_gpu_init_ex(ctypes.byref(self._ctx), ctypes.byref(cfg))
        if st != 0:
            self._ctx = ctypes.c_void_p()
            raise BsGpuError(st, "bs_gpu_init_ex")
        self._views = {}

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._views = {}
            self._L.bs_gpu_fini(self._ctx)
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
            raise BsGpuError(st, where, self._L.bs_gpu_last_error(self._ctx).decode())

    # -- the pinned SoA arrays (reference globals) ------------------------------------------------
    def host(self, which):
        """numpy view (no copy) of a pinned host stream; `which` is a BUF_* index or its name."""
        if isinstance(which, str):
            which = BUF_NAMES.index(which)
        if which not in self._views:
            ptr = self._L.bs_gpu_host_buffer(self._ctx, which)
            if not ptr:
                raise BsGpuError(-5, "bs_gpu_host_buffer", "no host staging in this context")
            ctype = ctypes.c_int32 if which == BUF_OTYPE else (ctypes.c_float if self.fp_bytes == 4 else ctypes.c_double)
            buf = (ctype * max(self.n, 1)).from_address(ptr)
            self._views[which] = np.frombuffer(buf, dtype=np.int32 if which == BUF_OTYPE else self.dtype)[: self.n]
        return self._views[which]

    def set_inputs(self, sptprice, strike, rate, volatility, otime, otype, dgrefval=None):
        for which, a in ((BUF_SPTPRICE, sptprice), (BUF_STRIKE, strike), (BUF_RATE, rate), (BUF_VOLATILITY, volatility),
                         (BUF_OTIME, otime), (BUF_OTYPE, otype)):
            self.host(which)[:] = a
        if dgrefval is not None:
            self.host(BUF_DGREFVAL)[:] = dgrefval
        self.mark_dirty()

    def mark_dirty(self):
        self._check(self._L.bs_gpu_mark_dirty(self._ctx), "bs_gpu_mark_dirty")

    @property
    def prices(self):
        return self.host(BUF_PRICES)

    # -- the hot path -----------------------------------------------------------------------------
    def price(self, num_runs=NUM_RUNS, err_chk=False):
        """bs_gpu_price: H2D (if dirty) -> num_runs launches -> D2H.  Returns the total error count."""
        errs = ctypes.c_ulonglong(0)
        self._check(self._L.bs_gpu_price(self._ctx, num_runs, 1 if err_chk else 0, ctypes.byref(errs)), "bs_gpu_price")
        return errs.value

    def price_aos(self, records, num_runs=NUM_RUNS):
        """bs_gpu_price_aos: the CAF Map's message (DATACONT_DTYPE records, blackscholes.c:482-489) -> prices."""
        rec = np.ascontiguousarray(records, dtype=DATACONT_DTYPE)
        out = np.empty(rec.shape[0], dtype=np.float32)
        self._check(self._L.bs_gpu_price_aos(self._ctx, rec.ctypes.data_as(ctypes.c_void_p), rec.shape[0],
                                             out.ctypes.data_as(ctypes.c_void_p), num_runs), "bs_gpu_price_aos")
        return out

    def upload(self):
        self._check(self._L.bs_gpu_upload(self._ctx), "bs_gpu_upload")

    def run(self, num_runs=NUM_RUNS, err_chk=False):
        errs = ctypes.c_ulonglong(0)
        self._check(self._L.bs_gpu_run(self._ctx, num_runs, 1 if err_chk else 0, ctypes.byref(errs)), "bs_gpu_run")
        return errs.value

    def download(self):
        self._check(self._L.bs_gpu_download(self._ctx), "bs_gpu_download")

    def fill_synthetic(self, first_index=0):
        self._check(self._L.bs_gpu_fill_synthetic(self._ctx, first_index), "bs_gpu_fill_synthetic")

    def read_device(self, which, first, count):
        if isinstance(which, str):
            which = BUF_NAMES.index(which)
        out = np.empty(count, dtype=np.int32 if which == BUF_OTYPE else self.dtype)
        self._check(self._L.bs_gpu_read_device(self._ctx, which, first, count, out.ctypes.data_as(ctypes.c_void_p)), "bs_gpu_read_device")
        return out

    def read_device_into(self, which, first, dst):
        """bs_gpu_read_device straight into `dst` (a contiguous numpy array, e.g. a slice of a pinned host stream)."""
        if isinstance(which, str):
            which = BUF_NAMES.index(which)
        assert dst.flags["C_CONTIGUOUS"] and dst.dtype == (np.int32 if which == BUF_OTYPE else self.dtype)
        self._check(self._L.bs_gpu_read_device(self._ctx, which, first, dst.shape[0], dst.ctypes.data_as(ctypes.c_void_p)), "bs_gpu_read_device")

    def errors(self, cap=65536):
        idx = np.empty(cap, dtype=np.int64)
        n = self._L.bs_gpu_errors(self._ctx, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), cap)
        if n < 0:
            raise BsGpuError(int(n), "bs_gpu_errors")
        return idx[:n].copy()

    # -- introspection ----------------------------------------------------------------------------
    def shards(self):
        out = []
        for g in range(self._L.bs_gpu_num_shards(self._ctx)):
            dev, first, count = ctypes.c_int(), ctypes.c_size_t(), ctypes.c_size_t()
            self._check(self._L.bs_gpu_shard(self._ctx, g, ctypes.byref(dev), ctypes.byref(first), ctypes.byref(count)), "bs_gpu_shard")
            out.append((dev.value, first.value, count.value))
        return out

    def timing(self):
        t = Timing()
        self._check(self._L.bs_gpu_get_timing(self._ctx, ctypes.byref(t)), "bs_gpu_get_timing")
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def launch(self):
        m, t, b = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self._L.bs_gpu_get_launch(self._ctx, ctypes.byref(m), ctypes.byref(t), ctypes.byref(b)), "bs_gpu_get_launch")
        return {"math": {1: "ieee", 2: "fast", 3: "reference"}.get(m.value, str(m.value)), "threads_per_block": t.value, "blocks": b.value}


def bytes_per_option(fp_bytes, err_chk=False):
    """ALGORITHMIC bytes per option per run: five fptype reads + int32 otype read + one fptype write
    (SURVEY.md 8d / blackscholes.c:328-331); +fptype for the DGrefval stream under ERR_CHK."""
    return 5 * fp_bytes + 4 + fp_bytes + (fp_bytes if err_chk else 0)


# ------------------------------------------------------------------------------------------------
# include/bs_io.h -- loader / writer (blackscholes.c:696-739,:760-767 and :923-947)
# ------------------------------------------------------------------------------------------------
IO_SYMBOLS = ("bs_io_open", "bs_io_load", "bs_io_close", "bs_io_write_prices", "bs_io_soa_write", "bs_io_soa_matches", "bs_io_is_soa")
IO_ERR_OPEN, IO_ERR_READ, IO_ERR_WRITE, IO_ERR_CLOSE = -1, -2, -3, -4
_io_ready = False


class BsIoError(IOError):
    def __init__(self, status, where):
        self.status = status
        msg = {IO_ERR_OPEN: "Unable to open file", IO_ERR_READ: "Unable to read from file",
               IO_ERR_WRITE: "Unable to write to file", IO_ERR_CLOSE: "Unable to close file"}.get(status, "error %d" % status)
        super().__init__("%s: %s" % (where, msg))


def _io():
    global _io_ready
    L = load_library()
    if not _io_ready:
        vp, ci, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        L.bs_io_open.restype, L.bs_io_open.argtypes = ci, [ctypes.c_char_p, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_longlong)]
        L.bs_io_load.restype, L.bs_io_load.argtypes = ci, [vp, ci, cs] + [vp] * 5 + [vp] + [vp] * 3 + [ci]
        L.bs_io_close.restype, L.bs_io_close.argtypes = ci, [vp]
        L.bs_io_write_prices.restype, L.bs_io_write_prices.argtypes = ci, [ctypes.c_char_p, ci, cs, vp, ci]
        L.bs_io_format_price.restype, L.bs_io_format_price.argtypes = ci, [ctypes.c_double, ctypes.c_char_p, cs]
        L.bs_io_soa_write.restype, L.bs_io_soa_write.argtypes = ci, [ctypes.c_char_p, ci, cs] + [vp] * 7 + [ctypes.c_char_p]
        L.bs_io_soa_matches.restype, L.bs_io_soa_matches.argtypes = ci, [ctypes.c_char_p, ctypes.c_char_p, ci]
        L.bs_io_is_soa.restype, L.bs_io_is_soa.argtypes = ci, [vp]
        _io_ready = True
    return L


def read_header(path):
    """numOptions as the reference's fscanf("%i") reads it."""
    L = _io()
    f, n = ctypes.c_void_p(), ctypes.c_longlong()
    st = L.bs_io_open(path.encode(), ctypes.byref(f), ctypes.byref(n))
    if st != 0:
        raise BsIoError(st, path)
    L.bs_io_close(f)
    return n.value


def load_options(path, fp_bytes=4, into=None, nthreads=0):
    """Parse an input file into SoA arrays.  `into` may be a BlackScholesGPU whose pinned host buffers
    receive the rows directly (then marked dirty); otherwise fresh numpy arrays are returned."""
    L = _io()
    f, n = ctypes.c_void_p(), ctypes.c_longlong()
    st = L.bs_io_open(path.encode(), ctypes.byref(f), ctypes.byref(n))
    if st != 0:
        raise BsIoError(st, path)
    try:
        count = n.value
        if count < 0:
            raise BsIoError(IO_ERR_READ, path)
        dt = np.float32 if fp_bytes == 4 else np.float64
        if into is not None:
            if into.n != count or into.fp_bytes != fp_bytes:
                raise ValueError("context was created for %d options of %d bytes, file has %d" % (into.n, into.fp_bytes, count))
            d = {k: into.host(k) for k in ("sptprice", "strike", "rate", "volatility", "otime", "otype", "dgrefval")}
        else:
            d = {k: np.empty(count, dtype=dt) for k in ("sptprice", "strike", "rate", "volatility", "otime", "dgrefval")}
            d["otype"] = np.empty(count, dtype=np.int32)
        d["divq"] = np.empty(count, dtype=dt)
        d["divs"] = np.empty(count, dtype=dt)
        ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        st = L.bs_io_load(f, fp_bytes, count, ptr(d["sptprice"]), ptr(d["strike"]), ptr(d["rate"]), ptr(d["volatility"]),
                          ptr(d["otime"]), ptr(d["otype"]), ptr(d["dgrefval"]), ptr(d["divq"]), ptr(d["divs"]), nthreads)
        if st != 0:
            raise BsIoError(st, path)
        if into is not None:
            into.mark_dirty()
        d["numOptions"] = count
        return d
    finally:
        L.bs_io_close(f)


def write_soa(path, d, source_path=None):
    """Write the binary SoA side-car of a loaded option set (dict as returned by load_options)."""
    L = _io()
    fp_bytes = d["sptprice"].dtype.itemsize
    ptr = lambda a: np.ascontiguousarray(a).ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    ref = d.get("dgrefval")
    st = L.bs_io_soa_write(path.encode(), fp_bytes, len(d["sptprice"]), ptr(d["sptprice"]), ptr(d["strike"]), ptr(d["rate"]),
                           ptr(d["volatility"]), ptr(d["otime"]), ptr(d["otype"]), ptr(ref) if ref is not None else None,
                           source_path.encode() if source_path else None)
    if st != 0:
        raise BsIoError(st, path)


def soa_matches(soa_path, source_path, fp_bytes=4):
    return bool(_io().bs_io_soa_matches(soa_path.encode(), source_path.encode(), fp_bytes))


def write_prices(path, prices, nthreads=0):
    L = _io()
    p = np.ascontiguousarray(prices)
    if p.dtype not in (np.float32, np.float64):
        raise ValueError("prices must be float32 or float64")
    st = L.bs_io_write_prices(path.encode(), p.dtype.itemsize, p.shape[0], p.ctypes.data_as(ctypes.c_void_p), nthreads)
    if st != 0:
        raise BsIoError(st, path)


def format_price(x):
    """The writer's rendering of one value ("%.18f\\n")."""
    L = _io()
    buf = ctypes.create_string_buffer(512)
    n = L.bs_io_format_price(float(x), buf, 512)
    if n < 0:
        raise BsIoError(n, "format_price")
    return buf.value.decode()
