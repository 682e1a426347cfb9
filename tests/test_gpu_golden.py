"""GPU tests against the committed golden fixtures (tests/golden/, generated from the reference by make_golden.py) and
against the live oracle on further configurations: cameras, shading, textures, framebuffers, accumulation, the row-band
partition, billboards, edge cases. All calls go through the C-ABI (yulio_raytracer_b200.devapi -> libyrt_device_cuda.so).

Stated tolerances:
  primary rays    |d org| <= 4 ulp(max |org|) + 1e-6 * eye separation scale, |d dir| <= 4e-7 (unit vectors; libm acosf/sinf/cosf/
                  atanf of CUDA vs glibc differ in the last bit)                                  [BASELINE: "stated ULP bound"]
  hit records     bit-exact
  images          equal spp, identical sample tables: mean relative error <= 2e-4 and <= 0.5 % of pixels beyond 1e-2 (a last-bit
                  difference in a transcendental can flip a discrete decision of a single path: lobe choice, Russian roulette)
"""
import os

import numpy as np
import pytest

from tests import scenes
from tests.golden import make_golden as mg
from tests.test_gpu_parity import assert_hits_bit_exact, ids, image_close

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "reference_outputs.npz"))


def assert_rays_close(g, o, scale):
    assert g.shape == o.shape
    assert np.array_equal(g[:, 3], o[:, 3]) and np.array_equal(g[:, 7], o[:, 7])          # tnear, tfar
    d_org = np.abs(g[:, 0:3] - o[:, 0:3]).max()
    d_dir = np.abs(g[:, 4:7] - o[:, 4:7]).max()
    assert d_org <= 4 * np.spacing(np.float32(scale)) + 1e-6, f"origin differs by {d_org}"
    assert d_dir <= 4e-7, f"direction differs by {d_dir}"


def test_primary_rays_pinhole_golden(cuda_dev, gold):
    s = scenes.cornell(cuda_dev, 16, 16, 4, 1)
    rays, _ = cuda_dev.primary_rays(s.renderer, s.camera, s.framebuffer, 16, 16, 4)
    assert_rays_close(rays, gold["primary_pinhole"], 800.0)
    # the pinhole camera has no transcendental per ray: bit-exact
    assert np.array_equal(rays.view(np.uint32), gold["primary_pinhole"].view(np.uint32))


@pytest.mark.parametrize("face", range(12))
def test_primary_rays_stereo_golden(cuda_dev, gold, face):
    s = scenes.spheres(cuda_dev, "mirror", 16, 16, 4, 1, face=face, num=8)
    rays, _ = cuda_dev.primary_rays(s.renderer, s.camera, s.framebuffer, 16, 16, 4)
    assert_rays_close(rays, gold[f"primary_stereo_{face}"], 200.0)


@pytest.mark.parametrize("face", [1, 8])
def test_primary_rays_stereo_toe_in_golden(cuda_dev, gold, face):
    s = scenes.spheres(cuda_dev, "mirror", 16, 16, 4, 1, face=None, num=8)
    pos, target, up = s.view
    cam = scenes.stereo_camera(cuda_dev, face, pos, target, up, toe_in=True, eye_separation=20.0, zero_parallax=200.0)
    rays, _ = cuda_dev.primary_rays(s.renderer, cam, s.framebuffer, 16, 16, 4)
    assert_rays_close(rays, gold[f"primary_stereo_toein_{face}"], 200.0)


def test_hits_golden(cuda_dev, gold):
    s = scenes.cornell(cuda_dev, 32, 32, 1, 1)
    assert_hits_bit_exact(cuda_dev.trace_rays(s.scene, gold["cornell_rays"], True)[0], gold["cornell_hits"])
    occ = cuda_dev.trace_rays(s.scene, gold["cornell_segments"], False)[0]
    assert np.array_equal(ids(occ)[:, 0], gold["cornell_occluded"])
    sp = scenes.soup(cuda_dev, 2000, seed=11, extent=20.0, meshes=2, cull=True, edge=0.8)
    assert_hits_bit_exact(cuda_dev.trace_rays(sp.scene, gold["soup_rays"], True)[0], gold["soup_hits"])


@pytest.mark.parametrize("name,build", [
    ("img_cornell_48_16spp_d2", lambda d: scenes.cornell(d, 48, 48, 16, 2)),
    ("img_cornell_32_4spp_d5", lambda d: scenes.cornell(d, 32, 32, 4, 5)),
    ("img_spheres_glass_32_16spp_d8", lambda d: scenes.spheres(d, "glass", 32, 32, 16, 8, face=None, num=12)),
    ("img_spheres_mirror_f3_32_8spp_d8", lambda d: scenes.spheres(d, "mirror", 32, 32, 8, 8, face=3, num=12)),
    ("img_atrium_f0_32_4spp_d4", lambda d: scenes.atrium(d, 32, 32, 4, 4, face=0, detail=2, tex_size=32)),
    ("img_atrium_f4_32_4spp_d4", lambda d: scenes.atrium(d, 32, 32, 4, 4, face=4, detail=2, tex_size=32)),
    ("img_atrium_f7_32_4spp_d4", lambda d: scenes.atrium(d, 32, 32, 4, 4, face=7, detail=2, tex_size=32)),
])
def test_images_golden(cuda_dev, gold, name, build):
    img = mg.render(cuda_dev, build(cuda_dev))
    image_close(img, gold[name])


def test_texture_fetch_golden(cuda_dev, gold):
    """G6: bilinear/nearest x RGBA8/RGB8 x invert, with border extrapolation, through a depth-1 ambient-lit card."""
    for k, (filt, inv, ch, seed) in enumerate(mg.TEXTURE_CASES):
        img = mg.render(cuda_dev, mg.texture_card(cuda_dev, mg.card_pixels(ch, seed), filt, inv))
        assert np.abs(img - gold[f"img_texcard_{k}"]).max() <= 2e-6, (k, np.abs(img - gold[f"img_texcard_{k}"]).max())
    img = mg.render(cuda_dev, mg.texture_card(cuda_dev, mg.card_pixels(4, 21), "bilinear", False, s0=(0.3, -0.2), ds=(2.5, 1.5)))
    assert np.abs(img - gold["img_texcard_scaled"]).max() <= 2e-6


# ---- live oracle comparisons beyond the fixtures ------------------------------------------------------------------
def both(cuda_dev, oracle_dev, build, **kw):
    out = []
    for d in (cuda_dev, oracle_dev):
        s = build(d)
        d.rtRenderFrame(s.renderer, s.camera, s.scene, s.tonemapper, s.framebuffer, kw.get("accumulate", 0))
        out.append((d, s))
    return out


@pytest.mark.parametrize("fmt", ["RGB8", "RGBA8", "RGB_FLOAT32"])
def test_framebuffer_formats(cuda_dev, oracle_dev, fmt):
    """FrameBufferRGB8/RGBA8/RGBFloat32::set (api/framebuffer.h:127-129,171-178,220-226) incl. the padded RGB8 stride (w = 50)."""
    w, h = 50, 34
    imgs = []
    for d in (cuda_dev, oracle_dev):
        s = scenes.cornell(d, w, h, 4, 2, fmt=fmt)
        d.rtRenderFrame(s.renderer, s.camera, s.scene, s.tonemapper, s.framebuffer, 0)
        imgs.append(d.read_framebuffer(s.framebuffer, fmt, w, h).astype(np.float64))
    if fmt == "RGB_FLOAT32":
        image_close(imgs[0], imgs[1])
    else:
        diff = np.abs(imgs[0] - imgs[1])
        assert diff.max() <= 1 and (diff > 0).mean() < 0.01     # truncation to 8 bit can flip on a last-bit difference


def test_gamma_and_vignetting(cuda_dev, oracle_dev):
    imgs = []
    for d in (cuda_dev, oracle_dev):
        s = scenes.cornell(d, 40, 40, 4, 2)
        s.tonemapper = scenes.tonemapper(d, gamma=2.2, vignetting=True)
        d.rtRenderFrame(s.renderer, s.camera, s.scene, s.tonemapper, s.framebuffer, 0)
        imgs.append(d.read_framebuffer(s.framebuffer, "RGB_FLOAT32", 40, 40))
    image_close(imgs[0], imgs[1], mean_tol=5e-4)


def test_accumulation_iterations(cuda_dev, oracle_dev):
    """accumulate != 0 keeps the AccuBuffer and advances the sampler chunk offset (integratorrenderer.cpp:67-69, sampler.cpp:93-97)."""
    imgs = []
    for d in (cuda_dev, oracle_dev):
        s = scenes.cornell(d, 32, 32, 4, 3)
        for it in range(3):
            d.rtRenderFrame(s.renderer, s.camera, s.scene, s.tonemapper, s.framebuffer, it)
        imgs.append(d.read_framebuffer(s.framebuffer, "RGB_FLOAT32", 32, 32))
    image_close(imgs[0], imgs[1])


def test_finite_shadow_ray_length_and_dome_light(cuda_dev, oracle_dev):
    """The dome-light shadow hack with a finite tMaxShadowRay and the pinned jitter hash (pathtraceintegrator.cpp:147-158; pins P1)."""
    imgs = []
    for d in (cuda_dev, oracle_dev):
        s = scenes.atrium(d, 40, 40, 8, 6, face=2, detail=3, tex_size=64, tmax_shadow=300.0)
        d.rtRenderFrame(s.renderer, s.camera, s.scene, s.tonemapper, s.framebuffer, 0)
        imgs.append(d.read_framebuffer(s.framebuffer, "RGB_FLOAT32", 40, 40))
    assert imgs[1].mean() > 0.05
    image_close(imgs[0], imgs[1])
    assert cuda_dev.frame_stats().rays_closest + cuda_dev.frame_stats().rays_shadow == oracle_dev.frame_stats().rays_closest


def test_billboard_update_primitive(cuda_dev, oracle_dev):
    """rtUpdatePrimitive of a faceCamera primitive + scene re-commit per cube face (singleray_device.cpp:354-398, renderer.cpp:551-559)."""
    imgs = {}
    for d in (cuda_dev, oracle_dev):
        s = scenes.atrium(d, 32, 32, 4, 3, face=0, detail=2, tex_size=32)
        out = []
        for face, cam in scenes.render_cube_map(d, s, faces=(1, 9)):
            out.append(d.read_framebuffer(s.framebuffer, "RGB_FLOAT32", 32, 32))
        imgs[d is cuda_dev] = out
    for a, b in zip(imgs[True], imgs[False]):
        image_close(a, b)


def test_row_band_partition_matches_reference_servers(cuda_dev, oracle_dev):
    """Multi-GPU partition at full C1 size: device `r` of `world` (cfg serverID/serverCount) renders exactly what server r of the
    reference's network device renders (rtSetInt1(NULL, "serverID"/"serverCount"), api/singleray_device.cpp:505-508;
    api/swapchain.h:57-70). NB the per-tile sample-set LCG is seeded with the server id and advances once per RENDERED pixel
    (integratorrenderer.cpp:134,149), so - in the reference as here - a banded frame is statistically, not bitwise, equal to the
    single-device frame; the union of the bands still covers every row exactly once."""
    from yulio_raytracer_b200 import Device, bands
    w = h = 512; spp = 4; world = 3
    covered = np.zeros(h, int)
    try:
        for r in range(world):
            d = Device.cuda(cfg=f"serverID={r},serverCount={world}")
            oracle_dev.rtSetInt1(None, "serverID", r); oracle_dev.rtSetInt1(None, "serverCount", world)
            parts = []
            for dev in (d, oracle_dev):
                sr = scenes.cornell(dev, w, h, spp, 2)
                dev.rtRenderFrame(sr.renderer, sr.camera, sr.scene, sr.tonemapper, sr.framebuffer, 0)
                parts.append(dev.read_framebuffer(sr.framebuffer, "RGB_FLOAT32", w, h))
            rows = bands.active_rows(h, r, world)
            covered[rows] += 1
            image_close(parts[0][: len(rows)], parts[1][: len(rows)])
            assert not parts[0][len(rows):].any()                                   # rows past the compacted bands stay untouched
            d.close()
    finally:
        oracle_dev.rtSetInt1(None, "serverID", 0); oracle_dev.rtSetInt1(None, "serverCount", 1)
    assert (covered == 1).all()


def test_determinism_and_readback_modes(cuda_dev):
    s = scenes.spheres(cuda_dev, "glass", 64, 64, 8, 8, face=5, num=16)
    a = mg.render(cuda_dev, s)
    b = mg.render(cuda_dev, s)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    cuda_dev.set_readback(False)
    try:
        c = mg.render(cuda_dev, s)            # copied on rtMapFrameBuffer instead
    finally:
        cuda_dev.set_readback(True)
    assert np.array_equal(a.view(np.uint32), c.view(np.uint32))


def test_edge_cases(cuda_dev, oracle_dev):
    # empty scene: every ray misses; an ambient light is seen directly (pathtraceintegrator.cpp:79-92)
    for d in (cuda_dev, oracle_dev):
        prims = [scenes.ambient_light(d, (0.25, 0.5, 0.75))]
        cam = scenes.pinhole(d, (0, 0, 0), (0, 0, 1), (0, 1, 0), 60.0, 1.0)
        s = scenes._bundle(d, prims, cam, scenes.pathtracer(d, 2, 3), 20, 12)
        img = mg.render(d, s)
        assert np.allclose(img, (0.25, 0.5, 0.75), atol=1e-6)
    # degenerate and out-of-range triangles are never hit; ragged sizes (1 pixel, non-multiple-of-16 frames)
    rays = scenes.random_rays(2000, seed=3, extent=4.0)
    res = []
    for d in (cuda_dev, oracle_dev):
        pos = [(0, 0, 0), (4, 0, 0), (0, 4, 0), (1, 1, 1), (1, 1, 1), (1, 1, 1), (float("nan"), 0, 0), (4, 4, 4)]
        tri = [(0, 1, 2), (3, 4, 5), (0, 1, 6), (0, 1, 99), (1, 2, 7)]
        m = d.rtNewMaterial("matte"); d.rtCommit(m)
        sc = scenes.make_scene(d, [d.rtNewShapePrimitive(scenes.add_mesh(d, pos, tri), m, None)])
        res.append(d.trace_rays(sc, rays, True)[0])
    assert_hits_bit_exact(res[0], res[1])
    assert set(np.unique(ids(res[0])[:, 1])) <= {-1, 0, 4}
    for w, h in ((1, 1), (17, 5), (33, 47)):
        imgs = []
        for d in (cuda_dev, oracle_dev):
            s = scenes.cornell(d, w, h, 2, 2)
            imgs.append(mg.render(d, s))
        image_close(imgs[0], imgs[1], frac_tol=0.02)


def test_error_behaviour(cuda_dev):
    """std::runtime_error of the reference -> RuntimeError (api/singleray_device.cpp:190..435)."""
    for call, arg in ((cuda_dev.rtNewCamera, "fisheye"), (cuda_dev.rtNewMaterial, "lava"), (cuda_dev.rtNewShape, "torus"),
                      (cuda_dev.rtNewLight, "laser"), (cuda_dev.rtNewRenderer, "gpt"), (cuda_dev.rtNewScene, "twolevel"),
                      (cuda_dev.rtNewTexture, "trilinear")):
        with pytest.raises(RuntimeError):
            call(arg)
    with pytest.raises(RuntimeError):
        cuda_dev.rtNewFrameBuffer("RGB16", 4, 4)
    cam = cuda_dev.rtNewCamera("pinhole")
    cuda_dev.rtSetFloat1(None, "angle", 1.0)                   # NULL handle: silently ignored (singleray_device.cpp:474-479)
    with pytest.raises(RuntimeError):
        cuda_dev.rtSetFloat1(cam, None, 1.0)                   # NULL property: error
    assert cuda_dev.rtGetFloat1(cam, "angle") == 0.0           # unset parameter reads back as zero
    cuda_dev.rtSetFloat1(cam, "angle", 33.0)
    assert cuda_dev.rtGetFloat1(cam, "angle") == 33.0
    cuda_dev.rtDecRef(cam)


def test_debug_renderer_id_image(cuda_dev, oracle_dev):
    """renderers/debugrenderer.cpp:66-148 at maxDepth 1: hash of geomID + primID through pixel corners (the G8 ID image)."""
    imgs = []
    for d in (cuda_dev, oracle_dev):
        s = scenes.cornell(d, 64, 48, 1, 1, fmt="RGB8")
        r = d.rtNewRenderer("debug"); d.rtSetInt1(r, "maxDepth", 1); d.rtSetInt1(r, "sampler.spp", 1); d.rtCommit(r)
        d.rtRenderFrame(r, s.camera, s.scene, s.tonemapper, s.framebuffer, 0)
        imgs.append(d.read_framebuffer(s.framebuffer, "RGB8", 64, 48))
    assert np.array_equal(imgs[0], imgs[1])


def test_plugin_boundary_with_a_reference_side_caller():
    """lib/plugin_smoke only knows devices/device/device.h: it dlopens a back end, dlsym("create")s it as
    Device::rtCreateDevice does (devices/device/device.cpp:24-35) and renders through the C++ virtual interface. The same
    binary drives the CUDA plugin and the reference's CPU back end; the frames must agree."""
    import subprocess
    from oracle import oracle_device
    from yulio_raytracer_b200.devapi import CUDA_LIB
    libdir = os.path.dirname(CUDA_LIB)
    exe, plugin = os.path.join(libdir, "plugin_smoke"), os.path.join(libdir, "libdevice_cuda.so")
    if not (os.path.exists(exe) and os.path.exists(plugin)):
        pytest.fail("adapter not built: run yulio_raytracer_b200/adapter/build_adapter.py where the reference is mounted")
    means = []
    for lib in (plugin, oracle_device.ORACLE_LIB):
        out = subprocess.run([exe, lib, "64"], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        line = [l for l in out.stdout.splitlines() if l.startswith("mean")][-1]
        means.append(np.array([float(v) for v in line.split()[1:]]))
    assert means[1].min() > 0.01
    assert np.allclose(means[0], means[1], rtol=2e-3), means


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["chunk=4096", "bvh=0", "splitleaves=0", "collapse=0", "collapse=1,ctri=20", "collapse=1,ctri=300", "bvh=0,collapse=0", "lanes=2", "syncmin=0", "syncmin=1000000000", "trav=0", "sort=1,sortmin=1",
                                 "refill=1,trinum=1,triden=4", "shadectas=3,tracectas=2"])
def test_frame_is_independent_of_scheduling_and_acceleration_structure(cuda_dev, cfg):
    """Radiance is accumulated per path in the reference's order and hits are the (t, geomID, primID) minimum, so the frame must be
    bit-identical whatever the wavefront chunking, queue order, traversal schedule, launch geometry or BVH builder."""
    from yulio_raytracer_b200 import Device
    ref = None
    for d in (cuda_dev, Device.cuda(cfg=cfg)):
        s = scenes.atrium(d, 48, 40, 8, 6, face=2, detail=4, tex_size=32)
        for _, _ in scenes.render_cube_map(d, s, faces=[2]):
            img = d.read_framebuffer(s.framebuffer, "RGB_FLOAT32", 48, 40)
        if ref is None:
            ref = img
        else:
            assert np.array_equal(ref.view(np.uint32), img.view(np.uint32)), cfg
            d.close()
