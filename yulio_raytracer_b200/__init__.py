"""device_cuda — B200-native render device behind the Yulio / Embree-example-renderer Device API.

The product is the C-ABI shared library `lib/libyrt_device_cuda.so` (sources in csrc/, entry points in
include/yrt_device.h). `devapi.Device` is a thin ctypes view of it used by bench.py and the tests.
"""
from .devapi import Device, FrameStats, CUDA_LIB, DECLARED_SYMBOLS  # noqa: F401
