"""Parity of the objects that were built but had no test of their own: point / spot / directional / distant lights
(lights/pointlight.h, spotlight.h, directionallight.h, distantlight.h), the depth-of-field camera (cameras/depthoffieldcamera.h), the disk
shape (shapes/disk.h:46-67), the backplate (integrators/pathtraceintegrator.cpp:79-84) and 1-pixel textures (pin P5). Same call sequence
on device_cuda and on the oracle; images at equal spp within the stated bound of tests/test_gpu_parity.py, ray counts identical."""
import numpy as np
import pytest

from tests import scenes
from tests.test_gpu_parity import assert_hits_bit_exact, image_close

pytestmark = pytest.mark.gpu
W = H = 48


def _cornell_with(dev, lights, camera=None, extra_prims=(), spp=8, depth=3, **kw):
    prims = scenes.cornell_prims(dev) + list(extra_prims(dev) if callable(extra_prims) else extra_prims) + lights(dev)
    cam = camera(dev) if camera else scenes.pinhole(dev, (278, 273, -800), (278, 273, 0), (0, 1, 0), 37.0, W / H)
    return scenes._bundle(dev, prims, cam, scenes.pathtracer(dev, spp, depth, **kw), W, H)


def _light(dev, kind, **p):
    l = dev.rtNewLight(kind)
    for k, v in p.items():
        if isinstance(v, (tuple, list)):
            dev.rtSetFloat3(l, k, *v)
        else:
            dev.rtSetFloat1(l, k, float(v))
    dev.rtCommit(l)
    return dev.rtNewLightPrimitive(l, None, None)


def _both(cuda_dev, oracle_dev, build, mean_tol=2e-4):
    imgs = []
    for d in (cuda_dev, oracle_dev):
        s = build(d)
        d.rtRenderFrame(s.renderer, s.camera, s.scene, s.tonemapper, s.framebuffer, 0)
        imgs.append(d.read_framebuffer(s.framebuffer, "RGB_FLOAT32", W, H))
    assert imgs[1].mean() > 1e-3, "the oracle image is black: the test lights nothing"
    image_close(imgs[0], imgs[1], mean_tol=mean_tol)
    sg, so = cuda_dev.frame_stats(), oracle_dev.frame_stats()
    assert sg.rays_closest + sg.rays_shadow == so.rays_closest, "ray counts differ (pathtraceintegrator.cpp:74,161)"
    return imgs


LIGHTS = {
    "pointlight": dict(P=(278, 400, 279), I=(4e5, 3.6e5, 3e5)),
    "spotlight": dict(P=(278, 540, 279), D=(0.1, -1, 0.05), I=(9e5, 9e5, 8e5), angleMin=20.0, angleMax=40.0),
    "spotlight_hard": dict(P=(278, 540, 279), D=(0, -1, 0), I=(9e5, 9e5, 8e5), angleMin=35.0, angleMax=35.0),
    "directionallight": dict(D=(0.3, -1, 0.6), E=(3, 3, 2.5)),
    "distantlight": dict(D=(0.2, -1, 0.5), L=(900, 850, 800), halfAngle=3.0),
}


@pytest.mark.parametrize("name", sorted(LIGHTS))
@pytest.mark.parametrize("tmax", [float("inf"), 150.0])
def test_analytic_lights(cuda_dev, oracle_dev, name, tmax):
    kind = name.split("_")[0]
    _both(cuda_dev, oracle_dev, lambda d: _cornell_with(d, lambda dd: [_light(dd, kind, **LIGHTS[name])], tmax_shadow=tmax))


def test_transformed_lights(cuda_dev, oracle_dev):
    """Light::transform through rtNewLightPrimitive's transform (lights/*.h)."""
    xfm = np.array([0, 0, 1, 0, 1, 0, -1, 0, 0, 20, -30, 10], np.float32)

    def lights(d):
        out = []
        for kind, p in (("pointlight", LIGHTS["pointlight"]), ("spotlight", LIGHTS["spotlight"]), ("distantlight", LIGHTS["distantlight"])):
            l = d.rtNewLight(kind)
            for k, v in p.items():
                d.rtSetFloat3(l, k, *v) if isinstance(v, tuple) else d.rtSetFloat1(l, k, float(v))
            d.rtCommit(l)
            out.append(d.rtNewLightPrimitive(l, None, xfm))
        return out
    _both(cuda_dev, oracle_dev, lambda d: _cornell_with(d, lights, tmax_shadow=150.0))


def test_depth_of_field_camera(cuda_dev, oracle_dev):
    def cam(d):
        c = d.rtNewCamera("depthoffield")
        d.rtSetTransform(c, "local2world", scenes.look_at((278, 273, -800), (278, 273, 0), (0, 1, 0)))
        d.rtSetFloat1(c, "angle", 37.0); d.rtSetFloat1(c, "aspectRatio", W / H)
        d.rtSetFloat1(c, "lensRadius", 25.0); d.rtSetFloat1(c, "focalDistance", 1000.0)
        d.rtCommit(c)
        return c
    lights = lambda d: scenes.quad_light(d, (213, 548.77, 227), (130, 0, 0), (0, 0, 105), (50, 50, 50))
    _both(cuda_dev, oracle_dev, lambda d: _cornell_with(d, lights, camera=cam))
    # primary rays of the lens camera: origins on the lens disk, bit-comparable up to sinf/cosf of the disk sample
    sg = _cornell_with(cuda_dev, lights, camera=cam)
    rays, _ = cuda_dev.primary_rays(sg.renderer, sg.camera, sg.framebuffer, W, H, 8)
    r = np.linalg.norm(rays[:, 0:3] - np.array([278, 273, -800], np.float32), axis=1)
    assert r.max() <= 25.0 * (1 + 1e-5) and r.max() > 20.0 and np.allclose(np.linalg.norm(rays[:, 4:7], axis=1), 1.0, atol=1e-5)


def test_disk_shape(cuda_dev, oracle_dev):
    def disk(d):
        m = d.rtNewMaterial("matte"); d.rtSetFloat3(m, "reflectance", .7, .6, .2); d.rtCommit(m)
        s = d.rtNewShape("disk")
        d.rtSetFloat3(s, "P", 278, 120, 279); d.rtSetFloat1(s, "h", 60.0); d.rtSetFloat1(s, "r", 150.0); d.rtSetInt1(s, "numTriangles", 24)
        d.rtCommit(s)
        return [d.rtNewShapePrimitive(s, m, None)]
    lights = lambda d: scenes.quad_light(d, (213, 548.77, 227), (130, 0, 0), (0, 0, 105), (50, 50, 50))
    _both(cuda_dev, oracle_dev, lambda d: _cornell_with(d, lights, extra_prims=disk))
    # and the tessellation itself: primary hits on the disk bit-exact (geometry built inside each device)
    sg, so = (_cornell_with(d, lights, extra_prims=disk) for d in (cuda_dev, oracle_dev))
    rays, _ = cuda_dev.primary_rays(sg.renderer, sg.camera, sg.framebuffer, W, H, 8)
    hg, _ = cuda_dev.trace_rays(sg.scene, rays, closest=True)
    ho, _ = oracle_dev.trace_rays(so.scene, rays, closest=True)
    assert (ho.view(np.int32)[:, 3] == 8).sum() > 100          # geomID 8 = the disk (after the 8 Cornell meshes)
    assert_hits_bit_exact(hg, ho)


def test_backplate(cuda_dev, oracle_dev):
    """Unbent paths that leave the scene read the backplate image at the pixel position (pathtraceintegrator.cpp:79-84); bent ones the
    environment lights."""
    rng = np.random.default_rng(5)
    plate = rng.random((20, 30, 3)).astype(np.float32)

    def build(d):
        img = d.rtNewImage("RGB_FLOAT32", 30, 20, plate)
        prims = scenes.spheres_prims(d, "glass", 12) + [scenes.ambient_light(d, (.4, .5, .6))]
        cam = scenes.pinhole(d, (-200, 100, 200), (0, 100, 200), (0, 1, 0), 90.0, W / H)
        return scenes._bundle(d, prims, cam, scenes.pathtracer(d, 8, 4, backplate=img), W, H)
    imgs = _both(cuda_dev, oracle_dev, build)
    assert imgs[0][: H // 4].std() > 0.05                      # the noise image shows through the sky part of the frame


def test_one_pixel_textures(cuda_dev, oracle_dev):
    """Pin P5: 1x1 (and 1xN, Nx1) images under the bilinear filter — the missing-texture fallback (singleray_device.cpp:250) and the sample
    scene's own 1x1 JPEGs — render their single colour on both sides instead of reading past the allocation."""
    def build(d):
        prims = []
        for k, (w, h) in enumerate(((1, 1), (1, 5), (6, 1))):
            px = np.full((h, w, 4), 255, np.uint8); px[..., k % 3] = 40 + 60 * k; px[..., 3] = 255
            tex, _ = scenes.texture(d, px)
            p, n, u, t = scenes.grid_quad((100 + 150 * k, 50, 300), (120, 0, 0), (0, 300, 0), 2, 2, (3, 2))
            prims.append(d.rtNewShapePrimitive(scenes.add_mesh(d, p, t, normals=n, uvs=u), scenes.uber(d, tex), None))
        prims += scenes.cornell_prims(d) + scenes.quad_light(d, (213, 548.77, 227), (130, 0, 0), (0, 0, 105), (50, 50, 50))
        cam = scenes.pinhole(d, (278, 273, -800), (278, 273, 0), (0, 1, 0), 37.0, W / H)
        return scenes._bundle(d, prims, cam, scenes.pathtracer(d, 8, 3), W, H)
    a = _both(cuda_dev, oracle_dev, build)[0]
    b = _both(cuda_dev, oracle_dev, build)[0]
    assert np.array_equal(a, b), "frames with 1-pixel textures must be reproducible"


@pytest.mark.parametrize("depth,spp", [(2, 1), (4, 2)])
def test_debug_renderer_with_diffuse_bounces(cuda_dev, oracle_dev, depth, spp):
    """renderers/debugrenderer.cpp:104-121 with maxDepth > 1: per-tile Random(tile * 1024) drawn in pixel / sample / bounce order (pin P6:
    u before v). The ID colour of the last hit must match the oracle's; a last-bit difference of sinf / cosf can move a bounce ray across
    a triangle edge, so up to 1 % of the pixels may show a neighbouring primitive."""
    imgs = []
    for d in (cuda_dev, oracle_dev):
        s = scenes.cornell(d, 64, 48, 1, 1, fmt="RGB8")
        r = d.rtNewRenderer("debug"); d.rtSetInt1(r, "maxDepth", depth); d.rtSetInt1(r, "sampler.spp", spp); d.rtCommit(r)
        d.rtRenderFrame(r, s.camera, s.scene, s.tonemapper, s.framebuffer, 0)
        imgs.append(d.read_framebuffer(s.framebuffer, "RGB8", 64, 48))
    differ = (imgs[0] != imgs[1]).any(axis=-1).mean()
    assert differ <= 0.01, f"{differ:.2%} of the pixels differ"
    assert len(np.unique(imgs[1].reshape(-1, 3), axis=0)) > 5


def test_tangent_arrays_and_bump_map(cuda_dev, oracle_dev):
    """Per-vertex tangent_x / tangent_y arrays (trianglemesh_full.cpp:252-270, read by BrushedMetal's anisotropic lobe) and the Obj
    material's map_Bump (materials/obj.h:52-56: Ns = normalize(b.x Tx + b.y Ty + b.z Ns))."""
    rng = np.random.default_rng(11)
    bump = np.clip(rng.normal(128, 40, (16, 16, 3)), 0, 255).astype(np.uint8); bump[..., 2] = 230

    def build(d):
        prims = []
        # a tilted quad carrying tangent arrays, brushed metal
        p, n, u, t = scenes.grid_quad((150, 20, 350), (260, 0, -60), (0, 240, 40), 3, 3)
        mesh = d.rtNewShape("trianglemesh")
        keep = []
        for key, ty, arr, stride in (("positions", "float3", p, 12), ("normals", "float3", n, 12), ("texcoords", "float2", u, 8), ("indices", "int3", t, 12),
                                     ("tangent_x", "float3", np.tile(np.array([[1.0, 0.2, 0.1]], np.float32), (len(p), 1)) * (1 + p[:, :1] / 500), 12),
                                     ("tangent_y", "float3", np.tile(np.array([[-0.1, 1.0, 0.3]], np.float32), (len(p), 1)), 12)):
            a = np.ascontiguousarray(arr, np.int32 if ty == "int3" else np.float32)
            h = d.rtNewData("immutable", a); keep.append(h)
            d.rtSetArray(mesh, key, ty, h, len(a), stride, 0)
        d.rtCommit(mesh)
        for h in keep:
            d.rtDecRef(h)
        prims.append(d.rtNewShapePrimitive(mesh, scenes.material(d, "BrushedMetal", reflectance=(.8, .8, .9), eta=(1.4, 1.2, 1.1), k=(5.0, 4.6, 4.2), roughnessX=.02, roughnessY=.3), None))
        # a bump-mapped Obj floor patch
        tex, _ = scenes.texture(d, bump)
        m = d.rtNewMaterial("obj")
        d.rtSetFloat3(m, "Kd", .7, .7, .6); d.rtSetFloat3(m, "Ks", .3, .3, .3); d.rtSetFloat1(m, "Ns", 20.0); d.rtSetTexture(m, "map_Bump", tex)
        d.rtCommit(m)
        p2, n2, u2, t2 = scenes.grid_quad((60, 1, 400), (430, 0, 0), (0, 0, -380), 2, 2, (3, 3))
        prims.append(d.rtNewShapePrimitive(scenes.add_mesh(d, p2, t2, normals=n2, uvs=u2), m, None))
        prims += scenes.cornell_prims(d) + scenes.quad_light(d, (213, 548.77, 227), (130, 0, 0), (0, 0, 105), (50, 50, 50)) + [scenes.ambient_light(d, (.2, .25, .3))]
        cam = scenes.pinhole(d, (278, 273, -800), (278, 273, 0), (0, 1, 0), 37.0, W / H)
        return scenes._bundle(d, prims, cam, scenes.pathtracer(d, 16, 4), W, H)
    _both(cuda_dev, oracle_dev, build, mean_tol=4e-4)


def _motion_scene(d, spp=16, depth=3):
    """models/sphere_motion.xml in miniature: a static and a moving sphere (dPdt), a quad whose far edge drops over the shutter interval
    ("motions" array, under a transform), the quad light and a dome light."""
    prims = []
    for P, dPdt, mat in (((150, 100, 280), (0, 0, 0), scenes.material(d, "matte", reflectance=(.7, .2, .2))),
                         ((400, 100, 330), (-90, 0, -60), scenes.material(d, "MetallicPaint", shadeColor=(.1, .5, .2), glitterColor=(.8, .8, .6), glitterSpread=.3, eta=1.45))):
        s = d.rtNewShape("sphere")
        d.rtSetFloat3(s, "P", *P); d.rtSetFloat3(s, "dPdt", *dPdt); d.rtSetFloat1(s, "r", 90.0)
        d.rtSetInt1(s, "numTheta", 14); d.rtSetInt1(s, "numPhi", 14)
        d.rtCommit(s)
        prims.append(d.rtNewShapePrimitive(s, mat, None))
    mesh = d.rtNewShape("trianglemesh")
    arrays = (("positions", "float3", [(-120, 0, -120), (120, 180, -120), (120, 180, 120), (-120, 0, 120)]),
              ("motions", "float3", [(0, 0, 0), (0, -70, 0), (0, -70, 0), (0, 0, 0)]),
              ("normals", "float3", [(0, 1, 0)] * 4), ("texcoords", "float2", [(0, 0), (1, 0), (1, 1), (0, 1)]), ("indices", "int3", [(0, 1, 2), (2, 3, 0)]))
    keep = []
    for key, ty, vals in arrays:
        a = np.ascontiguousarray(vals, np.int32 if ty == "int3" else np.float32)
        h = d.rtNewData("immutable", a); keep.append(h)
        d.rtSetArray(mesh, key, ty, h, len(a), a.shape[1] * 4, 0)
    d.rtCommit(mesh)
    for h in keep:
        d.rtDecRef(h)
    xfm = np.array([0, 0, 1, 0, 1, 0, -1, 0, 0, 300, 0, 150], np.float32)               # rotated and translated: motion vectors turn with it
    prims.append(d.rtNewShapePrimitive(mesh, scenes.material(d, "matte", reflectance=(.3, .4, .8)), xfm))
    fp, fn, fu, ft = scenes.grid_quad((-400, 0, 900), (1400, 0, 0), (0, 0, -1300), 2, 2)
    prims.append(d.rtNewShapePrimitive(scenes.add_mesh(d, fp, ft, normals=fn, uvs=fu), scenes.material(d, "matte", reflectance=(.6, .6, .6)), None))
    prims += scenes.quad_light(d, (213, 548.77, 227), (130, 0, 0), (0, 0, 105), (50, 50, 50)) + [scenes.ambient_light(d, (.2, .25, .3))]
    cam = scenes.pinhole(d, (278, 273, -800), (278, 150, 0), (0, 1, 0), 37.0, W / H)
    return scenes._bundle(d, prims, cam, scenes.pathtracer(d, spp, depth, tmax_shadow=400.0), W, H)


def test_motion_blur(cuda_dev, oracle_dev):
    """Linear vertex motion over the shutter interval (trianglemesh_full.cpp:29-32,104-110,150-164; sphere.h:62): the ray's time comes from the
    sample (integratorrenderer.cpp:160), travels with the path and its shadow rays, moving triangles are interpolated per ray."""
    imgs = _both(cuda_dev, oracle_dev, _motion_scene, mean_tol=4e-4)
    # the frame does not depend on the scheduling or the BVH builder
    from yulio_raytracer_b200 import Device
    for cfg in ("bvh=0,chunk=1000", "lanes=2,refill=1,trinum=1,triden=4"):
        d = Device.cuda(cfg=cfg)
        s = _motion_scene(d)
        d.rtRenderFrame(s.renderer, s.camera, s.scene, s.tonemapper, s.framebuffer, 0)
        assert np.array_equal(d.read_framebuffer(s.framebuffer, "RGB_FLOAT32", W, H).view(np.uint32), imgs[0].view(np.uint32)), cfg
        d.close()
