"""The Linux re-host of the Yulio front end (frontend/yulio_rt.cpp: StartRT / WaitRT / StopRT / GetLastErrorRT / GetCurrentStatusRT,
the reference's own Collada loader + vendored Assimp above the device boundary — SURVEY §8f-1, §3.1, §3.2).

CPU tests drive it on the reference CPU back end (the oracle library exports the reference's `create`); the GPU test renders the
same .dae through libdevice_cuda.so and compares the two 12W x H stereo cube-map strips."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests import dae_scene

# a blocked StartRT/StopRT sits in native code: let pytest-timeout end the run instead of waiting forever
pytestmark = pytest.mark.timeout(300, method="thread")

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "yulio_raytracer_b200", "lib")
FRONTEND = os.path.join(LIB, "libyulio_rt.so")
RT_TEST = os.path.join(LIB, "rt_test")
ORACLE = os.path.join(REPO, "oracle", "_ref", "liboracle_singleray.so")


class ParamsRT(ctypes.Structure):           # devices/renderer/YulioRT.h:36-51
    _fields_ = [("renderer", ctypes.c_char_p), ("size", ctypes.c_int), ("depth", ctypes.c_int), ("tMaxShadowRay", ctypes.c_float),
                ("spp", ctypes.c_int), ("ambientlight", ctypes.c_float * 3), ("eyeSeparation", ctypes.c_float), ("toeIn", ctypes.c_bool),
                ("zeroParallax", ctypes.c_float), ("jpegQuality", ctypes.c_int), ("debug", ctypes.c_bool), ("threadsPriority", ctypes.c_int),
                ("waterMark", ctypes.c_bool), ("faceCullingMode", ctypes.c_char_p)]


class StatusRT(ctypes.Structure):           # YulioRT.h:28-34
    _fields_ = [("state", ctypes.c_int), ("progress", ctypes.c_float), ("lastError", ctypes.c_int)]


def params(size, spp, depth, debug=False, toe_in=True):
    return ParamsRT(b"pathtracer", size, depth, 120.0, spp, (ctypes.c_float * 3)(.83, .95, .98), 2.5, toe_in, 75.0, 90, debug, 0, False, b"default")


def need_frontend():
    for f in (FRONTEND, RT_TEST):
        if not os.path.exists(f):
            pytest.fail(f"{f} missing: run `python frontend/build_frontend.py` where /root/reference is mounted")


def run_rt_test(dae, size, spp, depth, device_lib=None, extra_env=None):
    env = dict(os.environ)
    env.pop("YULIO_RT_DEVICE_LIB", None)
    if device_lib:
        env["YULIO_RT_DEVICE_LIB"] = device_lib
    env.update(extra_env or {})
    return subprocess.run([RT_TEST, dae, str(size), str(spp), str(depth)], env=env, capture_output=True, text=True, timeout=600)


@pytest.fixture(scope="module")
def front():
    need_frontend()
    os.environ["YULIO_RT_DEVICE_LIB"] = ORACLE
    lib = ctypes.CDLL(FRONTEND)
    lib.StartRT.restype = ctypes.c_bool; lib.StartRT.argtypes = [ctypes.c_char_p, ctypes.POINTER(ParamsRT)]
    lib.WaitRT.restype = ctypes.c_bool; lib.StopRT.restype = ctypes.c_bool; lib.StopRT.argtypes = [ctypes.c_bool]
    lib.GetLastErrorRT.restype = ctypes.c_int; lib.GetCurrentStatusRT.argtypes = [ctypes.POINTER(StatusRT)]
    yield lib
    os.environ.pop("YULIO_RT_DEVICE_LIB", None)


def test_exports():
    need_frontend()
    lib = ctypes.CDLL(FRONTEND)
    for name in ("StartRT", "WaitRT", "StopRT", "GetLastErrorRT", "GetCurrentStatusRT"):      # YulioRT.h:53-57
        assert hasattr(lib, name)


def test_error_codes(front, tmp_path):
    assert not front.WaitRT() and not front.StopRT(True)                                       # nothing running (renderer.cpp:1612-1640)
    assert not front.StartRT(None, None) and front.GetLastErrorRT() == 2                       # MissingColladaFile
    assert not front.StartRT(b"scene.obj", None) and front.GetLastErrorRT() == 2               # wrong extension (renderer.cpp:1542-1547)
    # a .dae without cameras -> InvalidColladaFormat from the worker (renderer.cpp:1500-1503)
    dae = dae_scene.write_scene(str(tmp_path), "nocam", views=())
    p = params(16, 1, 2)
    assert front.StartRT(dae.encode(), ctypes.byref(p))
    assert front.WaitRT()
    assert front.GetLastErrorRT() == 3


def test_cube_map_on_reference_backend(front, tmp_path):
    """12 cameras per tagged viewpoint, 12W x H strip per viewpoint named <dae>_<camera name>, 12 face images with debug=true."""
    dae = dae_scene.write_scene(str(tmp_path), "room", views=(("A", (0, 60, 0)), ("B", (-90, 70, 40))))
    p = params(16, 1, 3, debug=True)
    assert front.StartRT(dae.encode(), ctypes.byref(p))
    assert not front.StartRT(dae.encode(), ctypes.byref(p)) and front.GetLastErrorRT() == 1    # RenderingIsInProgress
    assert front.WaitRT()
    st = StatusRT(); front.GetCurrentStatusRT(ctypes.byref(st))
    assert (st.state, st.progress) == (4, 1.0)                                                 # Done
    for v in "AB":
        strip = dae_scene.read_ppm(str(tmp_path / f"room_{v}.ppm"))
        assert strip.shape == (16, 12 * 16, 3) and strip.max() > 0
        faces = [dae_scene.read_ppm(str(tmp_path / f"room_{v}_{n}_image_{e}.ppm")) for e in ("left", "right")
                 for n in ("front", "right", "back", "left", "top", "bottom")]
        # strip segments: Left Right Up Down Back Front of cameras 6-11, then of cameras 0-5 (renderer.cpp:677-710)
        order = [3, 1, 4, 5, 2, 0]
        for seg in range(12):
            src = (6 if seg < 6 else 0) + order[seg % 6]
            assert np.array_equal(strip[:, seg * 16:(seg + 1) * 16], faces[src])


def test_stop_discards_results(front, tmp_path):
    dae = dae_scene.write_scene(str(tmp_path), "room")
    p = params(64, 64, 8)
    assert front.StartRT(dae.encode(), ctypes.byref(p))
    assert front.StopRT(False)
    st = StatusRT(); front.GetCurrentStatusRT(ctypes.byref(st))
    assert st.state == 3 and not os.path.exists(str(tmp_path / "room_A.ppm"))                  # Stopped, nothing kept


def test_stop_keeps_results_and_progress_is_monotonic(front, tmp_path):
    """StopRT(true) keeps what was written so far (renderer.cpp:728-736); GetCurrentStatusRT reports (stage + tile fraction) / stages,
    which never decreases during a run (renderer.cpp:99-225)."""
    import time
    dae = dae_scene.write_scene(str(tmp_path), "room", views=(("A", (0, 60, 0)), ("B", (-90, 70, 40)), ("C", (60, 50, -40))))
    p = params(24, 4, 4, debug=True)
    assert front.StartRT(dae.encode(), ctypes.byref(p))
    seen = []
    st = StatusRT()
    for _ in range(400):
        front.GetCurrentStatusRT(ctypes.byref(st)); seen.append(st.progress)
        if os.path.exists(str(tmp_path / "room_A.ppm")) or st.state in (3, 4):
            break
        time.sleep(0.01)
    assert front.StopRT(True)
    front.GetCurrentStatusRT(ctypes.byref(st))
    assert st.state in (3, 4) and st.progress == 1.0
    assert all(b >= a for a, b in zip(seen, seen[1:])), seen
    assert os.path.exists(str(tmp_path / "room_A.ppm"))                                         # the finished viewpoint stays


@pytest.mark.gpu
def test_cube_map_cuda_matches_reference_backend(tmp_path):
    """Same .dae, same loader, same front end: device_cuda against the reference CPU device. Equal sample tables and per-path
    decisions -> the RGB8 strips agree to quantisation except where libm/CUDA last-bit differences flip a path decision.
    device_cuda's own output is the JPEG strip assembled and encoded on the GPU; YULIO_RT_HOST_STRIP=1 selects the
    host assembly + .ppm so that the frames can also be compared exactly."""
    need_frontend()
    from PIL import Image
    d1, d2, d3 = tmp_path / "cuda", tmp_path / "cpu", tmp_path / "cuda_jpg"
    dae1 = dae_scene.write_scene(str(d1), "room", scene_scale=1.5)
    dae2 = dae_scene.write_scene(str(d2), "room", scene_scale=1.5)
    dae3 = dae_scene.write_scene(str(d3), "room", scene_scale=1.5)
    r1 = run_rt_test(dae1, 64, 16, 6, extra_env={"YULIO_RT_HOST_STRIP": "1"})
    assert r1.returncode == 0, r1.stdout + r1.stderr
    r2 = run_rt_test(dae2, 64, 16, 6, device_lib=ORACLE)
    assert r2.returncode == 0, r2.stdout + r2.stderr
    a = dae_scene.read_ppm(str(d1 / "room_A.ppm")).astype(np.int32)
    b = dae_scene.read_ppm(str(d2 / "room_A.ppm")).astype(np.int32)
    assert a.shape == b.shape == (64, 12 * 64, 3)
    diff = np.abs(a - b)
    assert diff.mean() <= 0.05 and (diff > 2).mean() <= 0.005, (diff.mean(), (diff > 2).mean(), diff.max())
    r3 = run_rt_test(dae3, 64, 16, 6)                                               # strip + JPEG on the device
    assert r3.returncode == 0, r3.stdout + r3.stderr
    assert not os.path.exists(str(d3 / "room_A.ppm"))
    j = np.asarray(Image.open(str(d3 / "room_A.jpg")).convert("RGB")).astype(np.float64)
    assert j.shape == a.shape
    from tests.test_formats import _libjpeg_psnr, _psnr
    assert _psnr(j, a) >= _libjpeg_psnr(a, 90, str(tmp_path / "ref.jpg")) - 1.5     # jpegQuality 90, 4:2:0: as good as libjpeg on this image


@pytest.mark.gpu
def test_cube_map_call_equals_the_per_face_loop(tmp_path):
    """The front end renders a viewpoint with ONE yrtxRenderCubeMap call on device_cuda; YULIO_RT_PER_FACE=1 keeps the reference's literal
    loop (12 x update / commit / rtRenderFrame). Same frames -> byte-identical JPEG strips, incl. a scene with a camera-aligned billboard
    and two viewpoints."""
    need_frontend()
    d1, d2 = tmp_path / "batched", tmp_path / "perface"
    views = (("A", (0, 60, 0)), ("B", (-90, 70, 40)))
    dae1 = dae_scene.write_scene(str(d1), "room", views=views)
    dae2 = dae_scene.write_scene(str(d2), "room", views=views)
    r1 = run_rt_test(dae1, 48, 8, 6)
    r2 = run_rt_test(dae2, 48, 8, 6, extra_env={"YULIO_RT_PER_FACE": "1"})
    assert r1.returncode == 0 and r2.returncode == 0, r1.stdout + r1.stderr + r2.stdout + r2.stderr
    for v in "AB":
        a, b = open(str(d1 / f"room_{v}.jpg"), "rb").read(), open(str(d2 / f"room_{v}.jpg"), "rb").read()
        assert len(a) > 1000 and a == b, f"viewpoint {v}: strips differ"
