"""TEST INFRASTRUCTURE ONLY. Opens the reference CPU device (oracle/_ref/liboracle_singleray.so =
the reference's own device_singleray sources + oracle/embree2_shim.cpp) through the same ctypes
class the product uses. Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs."""
import ctypes as C
import os

import numpy as np

from yulio_raytracer_b200.devapi import Device

ORACLE_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "liboracle_singleray.so")

TRACE_DTYPE = np.dtype([("org", "f4", 3), ("tnear", "f4"), ("dir", "f4", 3), ("tfar", "f4"),
                        ("t", "f4"), ("u", "f4"), ("v", "f4"), ("geomID", "i4"), ("primID", "i4"),
                        ("kind", "i4"), ("Ng_x", "f4"), ("Ng_y", "f4")])
assert TRACE_DTYPE.itemsize == 64


def available() -> bool:
    return os.path.exists(ORACLE_LIB)


ORACLE_FAST_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "liboracle_singleray_fast.so")


def open_oracle(num_threads: int = 0, cfg: str = "", fast: bool = False) -> Device:
    """fast=True: the -O3 -ffast-math build of the same sources (oracle/build_ref.py), for bench.py's CPU baseline only — never parity."""
    return Device(ORACLE_FAST_LIB if fast else ORACLE_LIB, num_threads=num_threads, cfg=cfg)


class RayLog:
    """Records every rtcIntersect/rtcOccluded call the reference makes (single-threaded devices only)."""

    def __init__(self, dev: Device, capacity: int):
        self.dev = dev
        self.buf = np.zeros(capacity, TRACE_DTYPE)
        dev.lib.yrt_shim_trace_begin.argtypes = [C.c_void_p, C.c_size_t]
        dev.lib.yrt_shim_trace_end.restype = C.c_size_t

    def __enter__(self):
        self.dev.lib.yrt_shim_trace_begin(self.buf.ctypes.data, len(self.buf))
        return self

    def __exit__(self, *a):
        n = self.dev.lib.yrt_shim_trace_end()
        if n > len(self.buf):
            raise RuntimeError(f"ray log overflow: {n} > {len(self.buf)}")
        self.records = self.buf[:n]
