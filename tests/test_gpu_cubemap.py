"""yrtxRenderCubeMap (the 12 faces of a viewpoint as one wavefront) against the per-face rtRenderFrame loop of the reference's front end
(devices/renderer/renderer.cpp:543-632): every frame must be bit-identical, whatever the chunking and the row-band partition, and the
scene commit of faces 1..11 (same camera origin -> same billboard vertices) must not rebuild the BVH."""
import numpy as np
import pytest

from tests import scenes

pytestmark = pytest.mark.gpu


def _per_face(dev, s, cams, w, h):
    out = []
    for cam in cams:
        org = dev.rtGetFloat3(cam, "origin")
        for j, p in enumerate(s.prims):
            dev.rtUpdatePrimitive(s.scene, j, p, org, s.view[2])
        dev.rtCommit(s.scene)
        dev.rtRenderFrame(s.renderer, cam, s.scene, s.tonemapper, s.framebuffer, 0)
        out.append(dev.read_framebuffer(s.framebuffer, "RGB_FLOAT32", w, h))
    return out


@pytest.mark.parametrize("cfg", ["", "chunk=3000", "serverID=1,serverCount=3"])
def test_cube_map_equals_twelve_frames(cfg):
    from yulio_raytracer_b200 import Device
    dev = Device.cuda(cfg=cfg)
    w, h = 40, 36
    s = scenes.atrium(dev, w, h, 4, 6, face=0, detail=4, tex_size=32)
    cams = scenes.cube_cameras(dev, s)
    ref = _per_face(dev, s, cams, w, h)
    builds = dev.frame_stats().bvh_builds
    assert builds <= 2, f"per-face commits rebuilt the BVH {builds} times (expected: initial build + at most one billboard turn)"
    fbs = [dev.rtNewFrameBuffer("RGB_FLOAT32", w, h, 1) for _ in cams]
    scenes.render_cube_map_batched(dev, s, cams, fbs)
    st = dev.frame_stats()
    for f, fb in enumerate(fbs):
        img = dev.read_framebuffer(fb, "RGB_FLOAT32", w, h)
        assert np.array_equal(img.view(np.uint32), ref[f].view(np.uint32)), f"face {f} differs ({cfg})"
    assert st.rays_closest > 0 and st.bvh_builds == builds
    # subsets and a repeated call
    scenes.render_cube_map_batched(dev, s, [cams[7], cams[2]], [fbs[0], fbs[1]])
    assert np.array_equal(dev.read_framebuffer(fbs[0], "RGB_FLOAT32", w, h).view(np.uint32), ref[7].view(np.uint32))
    assert np.array_equal(dev.read_framebuffer(fbs[1], "RGB_FLOAT32", w, h).view(np.uint32), ref[2].view(np.uint32))
    dev.close()


def test_cube_map_rejects_bad_arguments(cuda_dev):
    s = scenes.cornell(cuda_dev, 16, 16, 1, 2)
    other = cuda_dev.rtNewFrameBuffer("RGB_FLOAT32", 8, 8, 1)
    with pytest.raises(RuntimeError):
        cuda_dev.render_cube_map(s.renderer, [s.camera, s.camera], s.scene, s.tonemapper, [s.framebuffer, other])     # sizes differ
    with pytest.raises(RuntimeError):
        cuda_dev.render_cube_map(s.renderer, [s.camera, s.camera], s.scene, s.tonemapper, [s.framebuffer, s.framebuffer])   # one buffer twice
    with pytest.raises(RuntimeError):
        cuda_dev.render_cube_map(s.renderer, [s.camera] * 13, s.scene, s.tonemapper, [s.framebuffer] * 13)


def test_cube_map_debug_renderer_and_rgb8(cuda_dev):
    d = cuda_dev
    s = scenes.spheres(d, "mirror", 24, 24, 2, 3, face=0, fmt="RGB8", num=8)
    cams = scenes.cube_cameras(d, s, faces=[0, 5, 9])
    fbs = [d.rtNewFrameBuffer("RGB8", 24, 24, 1) for _ in cams]
    for r in (s.renderer, None):
        if r is None:
            r = d.rtNewRenderer("debug"); d.rtSetInt1(r, "maxDepth", 1); d.rtSetInt1(r, "sampler.spp", 1); d.rtCommit(r)
        ref = []
        for cam in cams:
            d.rtRenderFrame(r, cam, s.scene, s.tonemapper, s.framebuffer, 0)
            ref.append(d.read_framebuffer(s.framebuffer, "RGB8", 24, 24))
        d.render_cube_map(r, cams, s.scene, s.tonemapper, fbs)
        for k, fb in enumerate(fbs):
            assert np.array_equal(d.read_framebuffer(fb, "RGB8", 24, 24), ref[k])


def test_two_chunk_lanes_render_the_same_frame():
    """A call of >= 2^22 paths is cut into chunks that run on two streams with their own wavefront state (cfg lanes, default 2):
    bit-identical to one lane, whatever the chunk size."""
    from yulio_raytracer_b200 import Device
    ref = None
    for cfg in ("lanes=1", "lanes=2", "lanes=2,chunk=700000"):
        d = Device.cuda(cfg=cfg)
        s = scenes.atrium(d, 192, 160, 256, 6, face=5, detail=4, tex_size=32)          # 7.9 M paths
        cams = scenes.cube_cameras(d, s, faces=[5, 10])
        fbs = [d.rtNewFrameBuffer("RGB_FLOAT32", 192, 160, 1) for _ in cams]
        scenes.render_cube_map_batched(d, s, cams, fbs)
        st = d.frame_stats()
        imgs = [d.read_framebuffer(fb, "RGB_FLOAT32", 192, 160) for fb in fbs]
        if ref is None:
            ref = (imgs, st.rays_closest, st.rays_shadow)
        else:
            assert (st.rays_closest, st.rays_shadow) == ref[1:], cfg
            for a, b in zip(imgs, ref[0]):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), cfg
        d.close()
