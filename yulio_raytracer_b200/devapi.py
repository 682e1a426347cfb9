"""ctypes binding of the device C-ABI declared in include/yrt_device.h.

The product library is `lib/libyrt_device_cuda.so` (built from csrc/ by __graft_entry__.build()).
The class is library-agnostic on purpose: oracle/oracle_device.py (test infrastructure) points it at
a second library that exports the very same symbols forwarded to the reference's CPU
`embree::Device` (devices/device/device.h:126-329), so the parity tests drive both with one script.
Nothing in this package references oracle/.

Method names and argument meaning follow the reference's Device interface one to one
(rtNewCamera, rtSetFloat3, rtCommit, rtRenderFrame, ...); a failed call raises RuntimeError like
the reference throws std::runtime_error (devices/device_singleray/api/singleray_device.cpp:190..435).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_HERE)
CUDA_LIB = os.path.join(_HERE, "lib", "libyrt_device_cuda.so")

H = C.c_void_p
_P = C.c_char_p


class FrameStats(C.Structure):
    _fields_ = [
        ("render_ms", C.c_double), ("build_ms", C.c_double), ("host_ms", C.c_double),
        ("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("trace_ms", C.c_double), ("node_visits", C.c_uint64), ("tri_tests", C.c_uint64),
        ("num_triangles", C.c_uint64), ("num_nodes", C.c_uint64),
        ("num_gpus", C.c_uint32), ("reserved", C.c_uint32),
        ("closest_ms", C.c_double), ("shadow_ms", C.c_double), ("shade_ms", C.c_double), ("raygen_film_ms", C.c_double),
        ("closest_launches", C.c_uint64), ("shadow_launches", C.c_uint64),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("sort_ms", C.c_double), ("bvh_builds", C.c_uint64),
        ("resolve_ms", C.c_double), ("node_visits_shadow", C.c_uint64), ("tri_tests_shadow", C.c_uint64), ("path_vertices", C.c_uint64),
        ("shade_launches", C.c_uint64), ("miss_ms", C.c_double), ("errors", C.c_uint64),
    ]


_SIGS = {
    # name: (restype, argtypes after the device pointer)
    "yrtNewCamera": (H, [_P]), "yrtNewData": (H, [_P, C.c_size_t, C.c_void_p]),
    "yrtNewDataFromFile": (H, [_P, _P, C.c_size_t, C.c_size_t]),
    "yrtNewImage": (H, [_P, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int]),
    "yrtNewImageFromFile": (H, [_P]), "yrtNewTexture": (H, [_P]), "yrtNewMaterial": (H, [_P]),
    "yrtNewShape": (H, [_P]), "yrtNewLight": (H, [_P]),
    "yrtNewShapePrimitive": (H, [H, H, C.c_void_p, C.c_int]),
    "yrtNewLightPrimitive": (H, [H, H, C.c_void_p]),
    "yrtTransformPrimitive": (H, [H, C.c_void_p]),
    "yrtNewScene": (H, [_P]), "yrtSetPrimitive": (C.c_int, [H, C.c_size_t, H]),
    "yrtUpdatePrimitive": (C.c_int, [H, C.c_size_t, H, C.c_void_p, C.c_void_p]),
    "yrtNewToneMapper": (H, [_P]), "yrtNewRenderer": (H, [_P]),
    "yrtNewFrameBuffer": (H, [_P, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]),
    "yrtMapFrameBuffer": (C.c_void_p, [H, C.c_int]), "yrtUnmapFrameBuffer": (C.c_int, [H, C.c_int]),
    "yrtSwapBuffers": (C.c_int, [H]), "yrtIncRef": (C.c_int, [H]), "yrtDecRef": (C.c_int, [H]),
    "yrtSetBool1": (C.c_int, [H, _P, C.c_int]), "yrtSetBool2": (C.c_int, [H, _P] + [C.c_int] * 2),
    "yrtSetBool3": (C.c_int, [H, _P] + [C.c_int] * 3), "yrtSetBool4": (C.c_int, [H, _P] + [C.c_int] * 4),
    "yrtSetInt1": (C.c_int, [H, _P, C.c_int]), "yrtSetInt2": (C.c_int, [H, _P] + [C.c_int] * 2),
    "yrtSetInt3": (C.c_int, [H, _P] + [C.c_int] * 3), "yrtSetInt4": (C.c_int, [H, _P] + [C.c_int] * 4),
    "yrtSetPointer": (C.c_int, [H, _P, C.c_void_p]),
    "yrtSetFloat1": (C.c_int, [H, _P, C.c_float]), "yrtGetFloat1": (C.c_int, [H, _P, C.POINTER(C.c_float)]),
    "yrtSetFloat2": (C.c_int, [H, _P] + [C.c_float] * 2), "yrtSetFloat3": (C.c_int, [H, _P] + [C.c_float] * 3),
    "yrtGetFloat3": (C.c_int, [H, _P] + [C.POINTER(C.c_float)] * 3),
    "yrtSetFloat4": (C.c_int, [H, _P] + [C.c_float] * 4),
    "yrtSetArray": (C.c_int, [H, _P, _P, H, C.c_size_t, C.c_size_t, C.c_size_t]),
    "yrtSetString": (C.c_int, [H, _P, _P]), "yrtGetString": (C.c_int, [H, _P, C.c_char_p, C.c_size_t]),
    "yrtSetImage": (C.c_int, [H, _P, H]), "yrtSetTexture": (C.c_int, [H, _P, H]),
    "yrtSetTransform": (C.c_int, [H, _P, C.c_void_p]), "yrtGetTransform": (C.c_int, [H, _P, C.c_void_p]),
    "yrtClear": (C.c_int, [H]), "yrtCommit": (C.c_int, [H]),
    "yrtRenderFrame": (C.c_int, [H, H, H, H, H, C.c_int]),
    "yrtPick": (C.c_int, [H, C.c_float, C.c_float, H] + [C.POINTER(C.c_float)] * 3),
    "yrtxGetFrameStats": (C.c_int, [C.POINTER(FrameStats)]),
    "yrtxTraceRays": (C.c_int, [H, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]),
}
_OPTIONAL = {
    "yrtxPrimaryRays": (C.c_int, [H, H, H, C.c_void_p, C.c_void_p]),
    "yrtxSampleTable": (C.c_int, [H, H, C.c_int] + [C.POINTER(C.c_int)] * 4 + [C.c_void_p]),
    "yrtxFrameBufferDevice": (C.c_int, [H, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "yrtxSetReadback": (C.c_int, [C.c_int]),
    "yrtxSetOption": (C.c_int, [_P, C.c_long]),
    "yrtxMicrobench": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_double)]),
    "yrtxRenderCubeMap": (C.c_int, [H, C.c_void_p, C.c_size_t, H, H, C.c_void_p, C.c_int]),
    "yrtxReadImage": (C.c_int, [H] + [C.POINTER(C.c_int)] * 3 + [C.c_void_p]),
    "yrtxStripBegin": (C.c_int, [C.c_size_t, C.c_size_t]), "yrtxStripSetWatermark": (C.c_int, [_P]),
    "yrtxStripAddFace": (C.c_int, [H, C.c_int, C.c_int]), "yrtxStripRead": (C.c_int, [C.c_void_p]),
    "yrtxStripEncodeJPEG": (C.c_int, [C.c_int, C.c_int, _P]),
}

_NODEV = {"yrtxHostSampleTable": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 3 + [C.c_void_p]),
          "yrtxDecodePNGFile": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p])}


def decode_png_file(libpath: str, file: str, flip_vertical: bool = False, flip_horizontal: bool = False) -> np.ndarray:
    """yrtxDecodePNGFile of `libpath` (works without a GPU): RGBA8 array [h, w, 4] in the reference's row order."""
    lib = C.CDLL(libpath, mode=C.RTLD_LOCAL)
    fn = lib.yrtxDecodePNGFile
    fn.restype, fn.argtypes = _NODEV["yrtxDecodePNGFile"]
    lib.yrtGetLastError.restype = _P
    w, h = C.c_int(0), C.c_int(0)
    if fn(_b(file), int(flip_vertical), int(flip_horizontal), C.byref(w), C.byref(h), None) != 0:
        raise RuntimeError(lib.yrtGetLastError().decode())
    out = np.zeros((h.value, w.value, 4), np.uint8)
    if fn(_b(file), int(flip_vertical), int(flip_horizontal), C.byref(w), C.byref(h), out.ctypes.data) != 0:
        raise RuntimeError(lib.yrtGetLastError().decode())
    return out


def host_sample_table(libpath: str, filter: str, spp: int, sets: int, max_depth: int, iteration: int = 0):
    """yrtxHostSampleTable of `libpath` (works without a GPU): returns (table[sets, spp_rounded, rec], n1, n2)."""
    lib = C.CDLL(libpath, mode=C.RTLD_LOCAL)
    fn = lib.yrtxHostSampleTable
    fn.restype, fn.argtypes = _NODEV["yrtxHostSampleTable"]
    lib.yrtGetLastError.restype = _P
    o, a, b = C.c_int(0), C.c_int(0), C.c_int(0)
    if fn(_b(filter), spp, sets, max_depth, iteration, C.byref(o), C.byref(a), C.byref(b), None) != 0:
        raise RuntimeError(lib.yrtGetLastError().decode())
    tab = np.zeros((sets, o.value, 5 + a.value + 2 * b.value), np.float32)
    if fn(_b(filter), spp, sets, max_depth, iteration, C.byref(o), C.byref(a), C.byref(b), tab.ctypes.data) != 0:
        raise RuntimeError(lib.yrtGetLastError().decode())
    return tab, a.value, b.value


#: every symbol include/yrt_device.h declares (checked by tests/test_cabi_exports.py)
DECLARED_SYMBOLS = ["yrtCreateDevice", "yrtDestroyDevice", "yrtGetLastError"] + list(_SIGS) + list(_OPTIONAL) + list(_NODEV)


def _b(s) -> Optional[bytes]:
    if s is None:
        return None
    return s if isinstance(s, bytes) else str(s).encode()


class Device:
    """One render device behind the C-ABI. `Device.cuda()` opens the product library."""

    def __init__(self, libpath: str, parms: str = "", num_threads: int = 0, priority: int = 0, cfg: str = ""):
        if not os.path.exists(libpath):
            raise RuntimeError(f"device library missing: {libpath} (run `python -c 'import __graft_entry__ as g; g.build()'`)")
        self.libpath = libpath
        self.lib = C.CDLL(libpath, mode=C.RTLD_LOCAL)
        self.lib.yrtCreateDevice.restype = C.c_void_p
        self.lib.yrtCreateDevice.argtypes = [_P, C.c_size_t, C.c_int, _P]
        self.lib.yrtDestroyDevice.argtypes = [C.c_void_p]
        self.lib.yrtGetLastError.restype = _P
        for name, (res, args) in list(_SIGS.items()) + list(_OPTIONAL.items()):
            fn = getattr(self.lib, name, None)
            if fn is None:
                if name in _OPTIONAL:
                    continue
                raise RuntimeError(f"{libpath} does not export {name}")
            fn.restype = res
            fn.argtypes = [C.c_void_p] + args
        self._keep = []  # numpy buffers that must outlive handles created with copy=False
        self.dev = self.lib.yrtCreateDevice(_b(parms), num_threads, priority, _b(cfg))
        if not self.dev:
            raise RuntimeError("device creation failed: " + self.last_error())

    # ---- construction helpers -------------------------------------------------------
    @classmethod
    def cuda(cls, cfg: str = "", **kw) -> "Device":
        return cls(CUDA_LIB, cfg=cfg, **kw)

    def close(self):
        if getattr(self, "dev", None):
            self.lib.yrtDestroyDevice(self.dev)
            self.dev = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_error(self) -> str:
        m = self.lib.yrtGetLastError()
        return m.decode(errors="replace") if m else ""

    def _h(self, name, *args):
        r = getattr(self.lib, name)(self.dev, *args)
        if not r:
            raise RuntimeError(f"{name[1:]}: {self.last_error()}")
        return r

    def _s(self, name, *args):
        if getattr(self.lib, name)(self.dev, *args) != 0:
            raise RuntimeError(f"{name[1:]}: {self.last_error()}")

    @staticmethod
    def _xfm(t):
        if t is None:
            return None
        a = np.ascontiguousarray(np.asarray(t, dtype=np.float32).reshape(12))
        return a

    # ---- creation (device.h:126-241) ------------------------------------------------
    def rtNewCamera(self, type): return self._h("yrtNewCamera", _b(type))

    def rtNewData(self, type, data) -> int:
        a = np.ascontiguousarray(data)
        return self._h("yrtNewData", _b(type), a.nbytes, a.ctypes.data if a.nbytes else None)

    def rtNewImage(self, type, width, height, data, copy=True):
        a = np.ascontiguousarray(data)
        if not copy:
            self._keep.append(a)
        return self._h("yrtNewImage", _b(type), width, height, a.ctypes.data, 1 if copy else 0)

    def rtNewImageFromFile(self, file): return self._h("yrtNewImageFromFile", _b(file))
    def rtNewTexture(self, type): return self._h("yrtNewTexture", _b(type))
    def rtNewMaterial(self, type): return self._h("yrtNewMaterial", _b(type))
    def rtNewShape(self, type): return self._h("yrtNewShape", _b(type))
    def rtNewLight(self, type): return self._h("yrtNewLight", _b(type))

    def rtNewShapePrimitive(self, shape, material, transform=None, faceCamera=False):
        t = self._xfm(transform)
        return self._h("yrtNewShapePrimitive", shape, material, t.ctypes.data if t is not None else None, int(faceCamera))

    def rtNewLightPrimitive(self, light, material=None, transform=None):
        t = self._xfm(transform)
        return self._h("yrtNewLightPrimitive", light, material, t.ctypes.data if t is not None else None)

    def rtTransformPrimitive(self, prim, transform):
        t = self._xfm(transform)
        return self._h("yrtTransformPrimitive", prim, t.ctypes.data)

    def rtNewScene(self, type="default"): return self._h("yrtNewScene", _b(type))
    def rtSetPrimitive(self, scene, slot, prim): self._s("yrtSetPrimitive", scene, slot, prim)

    def rtUpdatePrimitive(self, scene, slot, prim, camPos, camUp):
        p = np.asarray(camPos, np.float32); u = np.asarray(camUp, np.float32)
        self._s("yrtUpdatePrimitive", scene, slot, prim, p.ctypes.data, u.ctypes.data)

    def rtNewToneMapper(self, type="default"): return self._h("yrtNewToneMapper", _b(type))
    def rtNewRenderer(self, type): return self._h("yrtNewRenderer", _b(type))

    def rtNewFrameBuffer(self, type, width, height, buffers=1):
        return self._h("yrtNewFrameBuffer", _b(type), width, height, buffers, None)

    def rtMapFrameBuffer(self, fb, bufID=-1) -> int: return self._h("yrtMapFrameBuffer", fb, bufID)
    def rtUnmapFrameBuffer(self, fb, bufID=-1): self._s("yrtUnmapFrameBuffer", fb, bufID)
    def rtSwapBuffers(self, fb): self._s("yrtSwapBuffers", fb)
    def rtIncRef(self, h): self._s("yrtIncRef", h)
    def rtDecRef(self, h): self._s("yrtDecRef", h)

    # ---- parameters (device.h:248-312) ----------------------------------------------
    def rtSetBool1(self, h, p, x): self._s("yrtSetBool1", h, _b(p), int(x))
    def rtSetInt1(self, h, p, x): self._s("yrtSetInt1", h, _b(p), int(x))
    def rtSetInt2(self, h, p, x, y): self._s("yrtSetInt2", h, _b(p), int(x), int(y))
    def rtSetInt3(self, h, p, x, y, z): self._s("yrtSetInt3", h, _b(p), int(x), int(y), int(z))
    def rtSetPointer(self, h, p, ptr): self._s("yrtSetPointer", h, _b(p), ptr)
    def rtSetFloat1(self, h, p, x): self._s("yrtSetFloat1", h, _b(p), float(x))
    def rtSetFloat2(self, h, p, x, y): self._s("yrtSetFloat2", h, _b(p), float(x), float(y))
    def rtSetFloat3(self, h, p, x, y, z): self._s("yrtSetFloat3", h, _b(p), float(x), float(y), float(z))
    def rtSetFloat4(self, h, p, x, y, z, w): self._s("yrtSetFloat4", h, _b(p), float(x), float(y), float(z), float(w))

    def rtGetFloat1(self, h, p) -> float:
        x = C.c_float(0)
        self._s("yrtGetFloat1", h, _b(p), C.byref(x))
        return x.value

    def rtGetFloat3(self, h, p):
        x, y, z = C.c_float(0), C.c_float(0), C.c_float(0)
        self._s("yrtGetFloat3", h, _b(p), C.byref(x), C.byref(y), C.byref(z))
        return (x.value, y.value, z.value)

    def rtSetArray(self, h, p, type, data, size, stride=-1, ofs=0):
        self._s("yrtSetArray", h, _b(p), _b(type), data, size, C.c_size_t(stride & (2**64 - 1)).value, ofs)

    def rtSetString(self, h, p, s): self._s("yrtSetString", h, _b(p), _b(s))

    def rtGetString(self, h, p) -> str:
        buf = C.create_string_buffer(1024)
        self._s("yrtGetString", h, _b(p), buf, 1024)
        return buf.value.decode()

    def rtSetImage(self, h, p, img): self._s("yrtSetImage", h, _b(p), img)
    def rtSetTexture(self, h, p, tex): self._s("yrtSetTexture", h, _b(p), tex)

    def rtSetTransform(self, h, p, t):
        a = self._xfm(t)
        self._s("yrtSetTransform", h, _b(p), a.ctypes.data)

    def rtGetTransform(self, h, p) -> np.ndarray:
        out = np.zeros(12, np.float32)
        self._s("yrtGetTransform", h, _b(p), out.ctypes.data)
        return out

    def rtClear(self, h): self._s("yrtClear", h)
    def rtCommit(self, h): self._s("yrtCommit", h)

    # ---- render calls (device.h:322-329) ---------------------------------------------
    def rtRenderFrame(self, renderer, camera, scene, tonemapper, framebuffer, accumulate=0):
        self._s("yrtRenderFrame", renderer, camera, scene, tonemapper, framebuffer, int(accumulate))

    def set_option(self, key: str, value: int):
        """yrtxSetOption: change an integer cfg key ("lanes", "tracectas", "shadectas", "syncmin", "timers", "verbose") on a live device."""
        self._s("yrtxSetOption", _b(key), int(value))

    def microbench(self, kind: int, nbytes: int = 0) -> float:
        """yrtxMicrobench: 0 FP32 FMA TFLOP/s, 1 read GB/s over an nbytes working set (L1 bypassed), 2 G warp-instructions/s."""
        v = C.c_double(0)
        self._s("yrtxMicrobench", int(kind), int(nbytes), C.byref(v))
        return v.value

    def render_cube_map(self, renderer, cameras, scene, tonemapper, framebuffers, accumulate=0):
        """yrtxRenderCubeMap: the faces of one viewpoint (1..12 cameras, one frame buffer each) as one wavefront."""
        n = len(cameras)
        assert n == len(framebuffers)
        cams = (C.c_void_p * n)(*cameras); fbs = (C.c_void_p * n)(*framebuffers)
        self._s("yrtxRenderCubeMap", renderer, cams, n, scene, tonemapper, fbs, int(accumulate))

    def rtPick(self, camera, x: float, y: float, scene):
        """(hit, (px, py, pz)) — device.h:329."""
        a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
        r = self.lib.yrtPick(self.dev, camera, C.c_float(x), C.c_float(y), scene, C.byref(a), C.byref(b), C.byref(c))
        if r < 0:
            raise RuntimeError(f"rtPick: {self.last_error()}")
        return r == 1, (a.value, b.value, c.value)

    # ---- convenience ---------------------------------------------------------------
    def read_framebuffer(self, fb, fmt: str, width: int, height: int) -> np.ndarray:
        """Map, copy out (honouring the reference's row strides, api/framebuffer.h:106,146,195), unmap."""
        ptr = self.rtMapFrameBuffer(fb)
        try:
            if fmt == "RGB_FLOAT32":
                n = width * height * 3
                a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n,)).copy()
                return a.reshape(height, width, 3)
            if fmt == "RGBA8":
                a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(width * height * 4,)).copy()
                return a.reshape(height, width, 4)
            if fmt == "RGB8":
                stride = (3 * width + 3) // 4 * 4
                a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(stride * height,)).copy()
                return a.reshape(height, stride)[:, : 3 * width].reshape(height, width, 3)
            raise ValueError(fmt)
        finally:
            self.rtUnmapFrameBuffer(fb)

    def frame_stats(self) -> FrameStats:
        st = FrameStats()
        self._s("yrtxGetFrameStats", C.byref(st))
        return st

    def trace_rays(self, scene, rays: np.ndarray, closest: bool = True):
        """rays: (n,8) float32 {org,tnear,dir,tfar}; returns ((n,8) hit records as float32 view, kernel ms)."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        hits = np.zeros((n, 8), np.float32)
        hits.view(np.int32)[:, 3:5] = -1
        ms = C.c_float(0)
        self._s("yrtxTraceRays", scene, n, rays.ctypes.data, hits.ctypes.data, int(closest), 0, C.byref(ms))
        return hits, ms.value

    def sample_table(self, renderer, scene=None, iteration=0):
        """Returns (table[sets, spp, rec], n1, n2) with rec = 5 + n1 + 2*n2 (yrtxSampleTable)."""
        sets, spp, n1, n2 = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        self._s("yrtxSampleTable", renderer, scene, iteration, C.byref(sets), C.byref(spp), C.byref(n1), C.byref(n2), None)
        rec = 5 + n1.value + 2 * n2.value
        tab = np.zeros((sets.value, spp.value, rec), np.float32)
        self._s("yrtxSampleTable", renderer, scene, iteration, C.byref(sets), C.byref(spp), C.byref(n1), C.byref(n2), tab.ctypes.data)
        return tab, n1.value, n2.value

    def framebuffer_device(self, fb):
        """(device pointer, bytes, row stride) of the frame the last rtRenderFrame produced (yrtxFrameBufferDevice)."""
        ptr, nbytes, stride = C.c_void_p(0), C.c_size_t(0), C.c_size_t(0)
        self._s("yrtxFrameBufferDevice", fb, C.byref(ptr), C.byref(nbytes), C.byref(stride))
        return ptr.value, nbytes.value, stride.value

    def set_readback(self, each_frame: bool):
        self._s("yrtxSetReadback", int(each_frame))

    def read_image(self, image) -> np.ndarray:
        """Pixels of an image handle as stored (yrtxReadImage): [h, w, channels] uint8 or float32."""
        w, h, f = C.c_int(0), C.c_int(0), C.c_int(0)
        self._s("yrtxReadImage", image, C.byref(w), C.byref(h), C.byref(f), None)
        ch, dt = {0: (3, np.uint8), 1: (4, np.uint8), 2: (3, np.float32), 3: (4, np.float32)}[f.value]
        out = np.zeros((h.value, w.value, ch), dt)
        self._s("yrtxReadImage", image, C.byref(w), C.byref(h), C.byref(f), out.ctypes.data)
        return out

    # ---- stereo cube-map strip on the device (include/yrt_device.h; renderer.cpp:620-725) ----
    def strip_begin(self, face_w: int, face_h: int): self._s("yrtxStripBegin", face_w, face_h)
    def strip_set_watermark(self, png_file): self._s("yrtxStripSetWatermark", _b(png_file) if png_file else None)
    def strip_add_face(self, fb, cube_face_index: int, watermark: bool = False): self._s("yrtxStripAddFace", fb, int(cube_face_index), int(watermark))
    def strip_read(self, face_w: int, face_h: int) -> np.ndarray:
        out = np.zeros((face_h, 12 * face_w, 3), np.uint8)
        self._s("yrtxStripRead", out.ctypes.data)
        return out
    def strip_encode_jpeg(self, file: str, quality: int = 90, cube_face_index: int = -1): self._s("yrtxStripEncodeJPEG", int(cube_face_index), int(quality), _b(file))

    def trace_rays_device(self, scene, rays_ptr: int, hits_ptr: int, n: int, closest: bool = True) -> float:
        """Device-resident rays/hits (8 floats each); returns the CUDA-event time of the traversal kernel in ms."""
        ms = C.c_float(0)
        self._s("yrtxTraceRays", scene, n, rays_ptr, hits_ptr, int(closest), 1, C.byref(ms))
        return ms.value

    def primary_rays(self, renderer, camera, fb, width, height, spp):
        rays = np.zeros((height * width * spp, 8), np.float32)
        sets = np.zeros(height * width, np.int32)
        self._s("yrtxPrimaryRays", renderer, camera, fb, rays.ctypes.data, sets.ctypes.data)
        return rays, sets
