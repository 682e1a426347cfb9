#!/usr/bin/env python3
"""Generates the golden fixtures in this directory by RUNNING THE REFERENCE (oracle/_ref/liboracle_singleray.so = the
reference's devices/device_singleray sources compiled where they lie + oracle/embree2_shim.cpp). Run in the build
container (where /root/reference is mounted and oracle/build_ref.py has run):

    python tests/golden/make_golden.py

The fixtures pin (a) the oracle itself (tests/test_oracle_golden.py, CPU) and (b) the CUDA device (tests/test_gpu_*.py).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import oracle_device  # noqa: E402
from tests import scenes  # noqa: E402
from yulio_raytracer_b200.devapi import host_sample_table  # noqa: E402

TABLE_CASES = [("bspline", 1, 2, 0), ("bspline", 16, 2, 0), ("bspline", 16, 2, 1), ("box", 4, 3, 0), ("none", 8, 8, 2), ("bspline", 64, 8, 0)]
HASH_CASES = [("bspline", 256, 10, 0), ("bspline", 64, 10, 3)]


def tile_order(width, height, spp):
    """(y, x, s) in the order IntegratorRenderer::RenderJob::renderTile visits them with one thread."""
    out = []
    ntx = (width + 15) // 16; nty = (height + 15) // 16
    for tile in range(ntx * nty):
        tx, ty = (tile % ntx) * 16, (tile // ntx) * 16
        for dy in range(16):
            y = ty + dy
            if y >= height:
                continue
            for dx in range(16):
                x = tx + dx
                if x >= width:
                    continue
                for s in range(spp):
                    out.append((y, x, s))
    return np.array(out)


def logged_primary_rays(dev, s, width, height, spp):
    """Primary rays of a maxDepth = 1 frame, reordered to index ((y*width + x)*spp + s)."""
    with oracle_device.RayLog(dev, width * height * spp * 4 + 16) as log:
        dev.rtRenderFrame(s.renderer, s.camera, s.scene, s.tonemapper, s.framebuffer, 0)
    rec = log.records[log.records["kind"] == 0]
    assert len(rec) == width * height * spp, (len(rec), width * height * spp)
    order = tile_order(width, height, spp)
    idx = (order[:, 0] * width + order[:, 1]) * spp + order[:, 2]
    rays = np.zeros((width * height * spp, 8), np.float32)
    rays[idx, 0:3] = rec["org"]; rays[idx, 3] = rec["tnear"]; rays[idx, 4:7] = rec["dir"]; rays[idx, 7] = rec["tfar"]
    return rays


def texture_card(dev, tex_pixels, filtering, invert, size=48, material="MatteTextured", s0=(0.0, 0.0), ds=(1.0, 1.0)):
    """A textured unit quad filling a pinhole view under a white ambient light, depth 1, no pixel filter, 1 spp:
    every pixel equals the texture fetch at its centre (G6 texture-fetch parity through the Device API)."""
    tex, _ = scenes.texture(dev, tex_pixels, filtering, invert)
    m = dev.rtNewMaterial(material)
    dev.rtSetTexture(m, "Kd", tex); dev.rtSetFloat2(m, "s0", *s0); dev.rtSetFloat2(m, "ds", *ds)
    dev.rtCommit(m)
    quad = scenes.add_mesh(dev, [(-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0)], [(0, 1, 2), (0, 2, 3)],
                           normals=[(0, 0, -1)] * 4, uvs=[(-0.25, -0.25), (1.25, -0.25), (1.25, 1.25), (-0.25, 1.25)])
    prims = [dev.rtNewShapePrimitive(quad, m, None), scenes.ambient_light(dev, (1, 1, 1))]
    cam = scenes.pinhole(dev, (0, 0, -1), (0, 0, 0), (0, 1, 0), 90.0, 1.0)
    r = scenes.pathtracer(dev, 1, 1, filter="none")
    return scenes._bundle(dev, prims, cam, r, size, size)


def render(dev, s):
    dev.rtRenderFrame(s.renderer, s.camera, s.scene, s.tonemapper, s.framebuffer, 0)
    return dev.read_framebuffer(s.framebuffer, s.format, s.width, s.height)


TEXTURE_CASES = [("bilinear", False, 4, 7), ("nearest", False, 4, 8), ("bilinear", True, 3, 9), ("nearest", True, 3, 10)]


def card_pixels(channels, seed, size=8):
    return np.random.default_rng(seed).integers(0, 256, (size, size + 3, channels), dtype=np.uint8)


def main():
    dev = oracle_device.open_oracle(num_threads=1)
    out = {}
    # G4 sample tables
    tabs = {}
    for f, spp, depth, it in TABLE_CASES:
        t, n1, n2 = host_sample_table(oracle_device.ORACLE_LIB, f, spp, 64, depth, it)
        tabs[f"{f}_{spp}_{depth}_{it}"] = t
    for f, spp, depth, it in HASH_CASES:
        t, _, _ = host_sample_table(oracle_device.ORACLE_LIB, f, spp, 64, depth, it)
        tabs[f"sha256_{f}_{spp}_{depth}_{it}"] = np.frombuffer(hashlib.sha256(t.tobytes()).digest(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "sample_tables.npz"), **tabs)

    # G1 primary rays: pinhole (cornell) and the 12 stereo cube cameras (spheres view), with toe-in variants
    w = h = 16; spp = 4
    s = scenes.cornell(dev, w, h, spp, 1)
    out["primary_pinhole"] = logged_primary_rays(dev, s, w, h, spp)
    for face in range(12):
        s = scenes.spheres(dev, "mirror", w, h, spp, 1, face=face, num=8)
        out[f"primary_stereo_{face}"] = logged_primary_rays(dev, s, w, h, spp)
    for face in (1, 8):
        s = scenes.spheres(dev, "mirror", w, h, spp, 1, face=None, num=8)
        pos, target, up = s.view
        s.camera = scenes.stereo_camera(dev, face, pos, target, up, toe_in=True, eye_separation=20.0, zero_parallax=200.0)
        out[f"primary_stereo_toein_{face}"] = logged_primary_rays(dev, s, w, h, spp)

    # G2 / G3 hit records and occlusion bits
    s = scenes.cornell(dev, 32, 32, 1, 1)
    rng = np.random.default_rng(5)
    rays = np.zeros((4096, 8), np.float32)
    rays[:, 0:3] = rng.uniform((20, 20, 20), (530, 530, 530), (4096, 3))
    d = rng.normal(0, 1, (4096, 3)); rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True); rays[:, 7] = np.inf
    out["cornell_rays"] = rays
    out["cornell_hits"] = dev.trace_rays(s.scene, rays, True)[0]
    seg = rays.copy(); seg[:, 7] = rng.uniform(10, 400, 4096)
    out["cornell_segments"] = seg
    out["cornell_occluded"] = dev.trace_rays(s.scene, seg, False)[0].view(np.int32)[:, 3].copy()
    sp = scenes.soup(dev, 2000, seed=11, extent=20.0, meshes=2, cull=True, edge=0.8)
    rays = scenes.random_rays(4096, seed=12, extent=20.0)
    out["soup_rays"] = rays
    out["soup_hits"] = dev.trace_rays(sp.scene, rays, True)[0]

    # G7 images (small, low spp: identical sample tables on both sides make them comparable pixel by pixel)
    out["img_cornell_48_16spp_d2"] = render(dev, scenes.cornell(dev, 48, 48, 16, 2))
    out["img_cornell_32_4spp_d5"] = render(dev, scenes.cornell(dev, 32, 32, 4, 5))
    out["img_spheres_glass_32_16spp_d8"] = render(dev, scenes.spheres(dev, "glass", 32, 32, 16, 8, face=None, num=12))
    out["img_spheres_mirror_f3_32_8spp_d8"] = render(dev, scenes.spheres(dev, "mirror", 32, 32, 8, 8, face=3, num=12))
    for face in (0, 4, 7):
        out[f"img_atrium_f{face}_32_4spp_d4"] = render(dev, scenes.atrium(dev, 32, 32, 4, 4, face=face, detail=2, tex_size=32))
    # G6 texture fetch cards
    for k, (filt, inv, ch, seed) in enumerate(TEXTURE_CASES):
        out[f"img_texcard_{k}"] = render(dev, texture_card(dev, card_pixels(ch, seed), filt, inv))
    out["img_texcard_scaled"] = render(dev, texture_card(dev, card_pixels(4, 21), "bilinear", False, s0=(0.3, -0.2), ds=(2.5, 1.5)))
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    for k, v in out.items():
        print(f"{k:40s} {v.shape} {v.dtype}")
    print("wrote", os.path.join(HERE, "reference_outputs.npz"), os.path.getsize(os.path.join(HERE, "reference_outputs.npz")), "bytes")


if __name__ == "__main__":
    main()
