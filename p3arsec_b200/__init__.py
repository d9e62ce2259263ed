"""p3arsec_b200 -- B200-native blackscholes Map behind the P3ARSEC driver surface.

Product = p3arsec_b200/lib/libbs_gpu.so (C ABI in include/bs_gpu.h, CUDA in csrc/) and the drop-in driver
p3arsec_b200/bin/blackscholes_gpu.  `host` binds the same C ABI for Python callers.
"""
from . import host  # noqa: F401
from .host import BlackScholesGPU, BsGpuError, load_library, device_count  # noqa: F401

__all__ = ["host", "BlackScholesGPU", "BsGpuError", "load_library", "device_count"]
