"""Workloads of BASELINE.json (C1..C5) as Device-API call sequences, shared by bench.py, the parity tests and smoke().

Every builder talks ONLY to the Device API (rtNewShape / rtSetArray / rtNewMaterial / rtNewLight / rtCommit ...),
with the call sequences the reference's loaders and front end make (cited per function), so the same function
populates the CUDA device and any other library that exports include/yrt_device.h (the tests pass the CPU oracle).
Nothing here reads /root/reference at run time: geometry is either the public Cornell-box measurement data or
procedural (seeded); the .dae files of C3 / C4 are stripped from the reference mount, so their stand-ins
are procedural interiors textured with the REAL image sets of those scenes, shipped as a data pack under data/
(tools/make_data_pack.py: models/Sponza/*.JPG and the 152 images of sample_scene/22 Frederick St. good_tempo/) and read
through rtNewImageFromFile exactly as rtLoadTexture does (devices/device/loaders/loaders.cpp:29-61).
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace

import numpy as np

F = np.float32
DATA_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")


# ------------------------------------------------------------------------------------------------
# helpers mirroring the loaders' API call sequences
# ------------------------------------------------------------------------------------------------
def look_at(pos, target, up) -> np.ndarray:
    """AffineSpace3f::lookAtPoint (common/math/affinespace.h:73-78) as the 12-float column-major array."""
    pos, target, up = (np.asarray(v, np.float64) for v in (pos, target, up))
    z = target - pos; z /= np.linalg.norm(z)
    u = np.cross(up, z); u /= np.linalg.norm(u)
    v = np.cross(z, u); v /= np.linalg.norm(v)
    return np.concatenate([u, v, z, pos]).astype(F)


def add_mesh(dev, positions, indices, normals=None, uvs=None, cull=False, uv_stride12=False):
    """devices/device/loaders/obj_loader.cpp:353-378, xml_loader.cpp:444-468, ColladaLoader.cpp:600-625."""
    positions = np.ascontiguousarray(positions, F).reshape(-1, 3)
    indices = np.ascontiguousarray(indices, np.int32).reshape(-1, 3)
    mesh = dev.rtNewShape("trianglemesh")
    dp = dev.rtNewData("immutable", positions)
    dev.rtSetArray(mesh, "positions", "float3", dp, len(positions), 12, 0)
    di = dev.rtNewData("immutable", indices)
    dev.rtSetArray(mesh, "indices", "int3", di, len(indices), 12, 0)
    keep = [dp, di]
    if normals is not None:
        normals = np.ascontiguousarray(normals, F).reshape(-1, 3)
        dn = dev.rtNewData("immutable", normals)
        dev.rtSetArray(mesh, "normals", "float3", dn, len(normals), 12, 0)
        keep.append(dn)
    if uvs is not None:
        uvs = np.ascontiguousarray(uvs, F).reshape(-1, 2)
        if uv_stride12:      # the Collada loader hands aiVector3D texture coordinates with stride 12 (ColladaLoader.cpp:620)
            uv3 = np.zeros((len(uvs), 3), F); uv3[:, :2] = uvs
            du = dev.rtNewData("immutable", uv3)
            dev.rtSetArray(mesh, "texcoords", "float2", du, len(uvs), 12, 0)
        else:
            du = dev.rtNewData("immutable", uvs)
            dev.rtSetArray(mesh, "texcoords", "float2", du, len(uvs), 8, 0)
        keep.append(du)
    if cull:
        dev.rtSetBool1(mesh, "cullBackFaces", True)
    dev.rtSetString(mesh, "accel", "default")
    dev.rtCommit(mesh)
    for d in keep:
        dev.rtDecRef(d)
    return mesh


def texture(dev, pixels: np.ndarray, filtering="bilinear", invert=False):
    """rtLoadTexture (devices/device/loaders/loaders.cpp:47-61) with the decoded pixels handed to rtNewImage."""
    h, w, c = pixels.shape
    if pixels.dtype == np.uint8:
        fmt = "RGB8" if c == 3 else "RGBA8"
    else:
        fmt = "RGB_FLOAT32" if c == 3 else "RGBA_FLOAT32"
    img = dev.rtNewImage(fmt, w, h, pixels, copy=True)
    tex = dev.rtNewTexture(filtering)
    dev.rtSetImage(tex, "image", img)
    dev.rtSetBool1(tex, "invert", invert)
    dev.rtCommit(tex)
    return tex, img


def decode_image_host(path) -> np.ndarray:
    """A JPEG / PNG file decoded on the host with libjpeg-turbo / libpng (PIL) into what the reference's readers store:
    RGBA8, alpha 255 where the file has none, row 0 = bottom scanline (common/image/jpeg.cpp:53-63, freeimage.cpp:40-78)."""
    from PIL import Image
    with Image.open(path) as im:
        px = np.asarray(im.convert("RGBA"), np.uint8)
    return np.ascontiguousarray(px[::-1])


def file_texture(dev, path, filtering="bilinear", pixels_from=None):
    """rtLoadTexture (devices/device/loaders/loaders.cpp:47-61): image file -> image handle -> texture handle.
    On device_cuda the file goes through rtNewImageFromFile (nvJPEG / the PNG reader of csrc/image_codecs.cu). A library without
    the codecs (the tests' CPU oracle: FreeImage / libjpeg-turbo are Windows binaries in the mount) gets the pixels through
    rtNewImage: the ones `pixels_from` (a device_cuda Device) decoded — bit-identical textures for the parity tests — or a
    host decode with libjpeg-turbo / libpng, the reference's own decoders."""
    if getattr(dev.lib, "yrtxReadImage", None) is not None:
        img = dev.rtNewImageFromFile(path)
    else:
        key = os.path.abspath(path)
        if pixels_from is not None:
            h = pixels_from.rtNewImageFromFile(path)
            px = pixels_from.read_image(h); pixels_from.rtDecRef(h)
        else:
            px = decode_image_host(key)
        img = dev.rtNewImage("RGBA8" if px.shape[2] == 4 else "RGB8", px.shape[1], px.shape[0], px, copy=True)
    tex = dev.rtNewTexture(filtering)
    dev.rtSetImage(tex, "image", img)
    dev.rtCommit(tex)
    return tex, img


def data_files(sub, exts=(".jpg", ".jpeg", ".png")):
    d = os.path.join(DATA_DIR, sub)
    if not os.path.isdir(d):
        raise RuntimeError(f"data pack missing: {d} (tools/make_data_pack.py copies it from the reference mount)")
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.lower().endswith(exts)]


def pathtracer(dev, spp, depth, tmax_shadow=math.inf, up=(0, 1, 0), jitter=0.2, filter=None, backplate=None, min_contribution=None):
    """createGlobalObjects (devices/renderer/renderer.cpp:352-361)."""
    r = dev.rtNewRenderer("pathtracer")
    if depth >= 0:
        dev.rtSetInt1(r, "maxDepth", depth)
    dev.rtSetFloat1(r, "tMaxShadowRay", tmax_shadow)
    dev.rtSetFloat1(r, "tMaxShadowJitter", jitter)
    dev.rtSetFloat3(r, "up", *up)
    dev.rtSetInt1(r, "sampler.spp", spp)
    if filter is not None:
        dev.rtSetString(r, "filter", filter)
    if backplate is not None:
        dev.rtSetImage(r, "backplate", backplate)
    if min_contribution is not None:
        dev.rtSetFloat1(r, "minContribution", min_contribution)
    dev.rtCommit(r)
    return r


def tonemapper(dev, gamma=1.0, vignetting=False):
    t = dev.rtNewToneMapper("default")
    dev.rtSetFloat1(t, "gamma", gamma)
    dev.rtSetBool1(t, "vignetting", vignetting)
    dev.rtCommit(t)
    return t


def pinhole(dev, pos, target, up, fov, aspect):
    """createCamera (devices/renderer/renderer.cpp:309-319)."""
    c = dev.rtNewCamera("pinhole")
    dev.rtSetTransform(c, "local2world", look_at(pos, target, up))
    dev.rtSetFloat1(c, "angle", fov)
    dev.rtSetFloat1(c, "aspectRatio", aspect)
    dev.rtCommit(c)
    return c


def stereo_camera(dev, face, pos, target, up, toe_in=False, eye_separation=None, scene_scale=None, zero_parallax=None):
    """The 12 stereo cube cameras of a viewpoint (devices/renderer/renderer.cpp:746-757; ColladaLoader.cpp:470-505)."""
    c = dev.rtNewCamera("stereo")
    dev.rtSetTransform(c, "local2world", look_at(pos, target, up))
    dev.rtSetInt1(c, "cubeFaceIndex", face)
    dev.rtSetFloat3(c, "origin", *pos)
    dev.rtSetFloat3(c, "lookAt", *target)
    dev.rtSetFloat3(c, "up", *up)
    if eye_separation is not None:
        dev.rtSetFloat1(c, "eyeSeparation", eye_separation)
    if scene_scale is not None:
        dev.rtSetFloat1(c, "sceneScale", scene_scale)
    if zero_parallax is not None:
        dev.rtSetFloat1(c, "zeroParallaxDistance", zero_parallax)
    dev.rtSetBool1(c, "toeIn", toe_in)
    dev.rtCommit(c)
    return c


def make_scene(dev, prims):
    """createScene (devices/renderer/renderer.cpp:334-343)."""
    sc = dev.rtNewScene("default")
    dev.rtSetString(sc, "accel", "default")
    for i, p in enumerate(prims):
        dev.rtSetPrimitive(sc, i, p)
    dev.rtCommit(sc)
    return sc


def quad_light(dev, P, U, V, L):
    """-quadlight (devices/renderer/renderer.cpp:1118-1140): two triangle lights."""
    P, U, V = (np.asarray(v, F) for v in (P, U, V))
    out = []
    for v0, v1, v2 in ((P + U + V, P + U, P), (P + U + V, P, P + V)):
        l = dev.rtNewLight("trianglelight")
        dev.rtSetFloat3(l, "v0", *v0); dev.rtSetFloat3(l, "v1", *v1); dev.rtSetFloat3(l, "v2", *v2)
        dev.rtSetFloat3(l, "L", *L)
        dev.rtCommit(l)
        out.append(dev.rtNewLightPrimitive(l, None, None))
    return out


def ambient_light(dev, L):
    """-ambientlight (devices/renderer/renderer.cpp:1026-1032)."""
    l = dev.rtNewLight("ambientlight")
    dev.rtSetFloat3(l, "L", *L)
    dev.rtCommit(l)
    return dev.rtNewLightPrimitive(l, None, None)


def _bundle(dev, prims, camera, renderer, width, height, fmt="RGB_FLOAT32", **extra):
    s = SimpleNamespace(prims=prims, camera=camera, renderer=renderer, width=width, height=height, format=fmt, **extra)
    s.scene = make_scene(dev, prims)
    s.tonemapper = tonemapper(dev)
    s.framebuffer = dev.rtNewFrameBuffer(fmt, width, height, 1)
    return s


# ------------------------------------------------------------------------------------------------
# C1: the Cornell box (public measurement data, http://www.graphics.cornell.edu/online/box/data.html), grouped by
# material run exactly as an OBJ loader flushing one mesh per `usemtl` would (obj_loader.cpp:185-192,316-379)
# ------------------------------------------------------------------------------------------------
_CORNELL_GROUPS = [
    ("white", [[(552.8, 0, 0), (0, 0, 0), (0, 0, 559.2), (549.6, 0, 559.2)],
               [(290, 0, 114), (240, 0, 272), (82, 0, 225), (130, 0, 65)],
               [(472, 0, 406), (314, 0, 456), (265, 0, 296), (423, 0, 247)]]),
    ("white", [[(556, 548.8, 0), (556, 548.8, 559.2), (0, 548.8, 559.2), (0, 548.8, 0)]]),
    ("white", [[(549.6, 0, 559.2), (0, 0, 559.2), (0, 548.8, 559.2), (556, 548.8, 559.2)]]),
    ("green", [[(0, 0, 559.2), (0, 0, 0), (0, 548.8, 0), (0, 548.8, 559.2)]]),
    ("red", [[(552.8, 0, 0), (549.6, 0, 559.2), (556, 548.8, 559.2), (556, 548.8, 0)]]),
    ("white", [[(130, 165, 65), (82, 165, 225), (240, 165, 272), (290, 165, 114)],
               [(290, 0, 114), (290, 165, 114), (240, 165, 272), (240, 0, 272)],
               [(130, 0, 65), (130, 165, 65), (290, 165, 114), (290, 0, 114)],
               [(82, 0, 225), (82, 165, 225), (130, 165, 65), (130, 0, 65)],
               [(240, 0, 272), (240, 165, 272), (82, 165, 225), (82, 0, 225)]]),
    ("white", [[(423, 330, 247), (265, 330, 296), (314, 330, 456), (472, 330, 406)]]),
    ("white", [[(423, 0, 247), (423, 330, 247), (472, 330, 406), (472, 0, 406)],
               [(472, 0, 406), (472, 330, 406), (314, 330, 456), (314, 0, 456)],
               [(314, 0, 456), (314, 330, 456), (265, 330, 296), (265, 0, 296)],
               [(265, 0, 296), (265, 330, 296), (423, 330, 247), (423, 0, 247)]]),
]
_CORNELL_KD = {"white": (1, 1, 1), "red": (1, 0, 0), "green": (0, 1, 0), "blue": (0, 0, 1)}


def cornell_prims(dev):
    mats = {}
    for name, kd in _CORNELL_KD.items():            # loadMTL (obj_loader.cpp:222-277)
        m = dev.rtNewMaterial("obj")
        dev.rtSetFloat3(m, "Ka", 0, 0, 0); dev.rtSetFloat3(m, "Kd", *kd); dev.rtSetFloat3(m, "Ks", 0, 0, 0)
        dev.rtCommit(m)
        mats[name] = m
    prims = []
    for mat, quads in _CORNELL_GROUPS:
        pos, tri = [], []
        for q in quads:                                # triangle fan (obj_loader.cpp:333-341)
            b = len(pos); pos.extend(q)
            tri.append((b, b + 1, b + 2)); tri.append((b, b + 2, b + 3))
        prims.append(dev.rtNewShapePrimitive(add_mesh(dev, pos, tri), mats[mat], None))
    return prims


def cornell(dev, width=512, height=512, spp=16, depth=2, fmt="RGB_FLOAT32", **kw):
    """models/cornell_box.ecs: -quadlight 213 548.77 227 130 0 0 0 0 105 50 50 50; -vp 278 273 -800 -vi 278 273 0 -fov 37;
    pathtracer { depth = 2 }."""
    prims = cornell_prims(dev) + quad_light(dev, (213, 548.77, 227), (130, 0, 0), (0, 0, 105), (50, 50, 50))
    cam = pinhole(dev, (278, 273, -800), (278, 273, 0), (0, 1, 0), 37.0, width / height)
    return _bundle(dev, prims, cam, pathtracer(dev, spp, depth, **kw), width, height, fmt)


# ------------------------------------------------------------------------------------------------
# C2: models/sphere_glass.xml / sphere_mirror.xml + sphere_view.ecs. `lines.ppm` (512x512 RGB8 line pattern) does not travel
# to the GPU box, so a seeded procedural line pattern of the same size/format stands in for it.
# ------------------------------------------------------------------------------------------------
def lines_image(size=512, seed=7) -> np.ndarray:
    rng = np.random.default_rng(seed)
    img = np.full((size, size, 4), 255, np.uint8)
    img[:, :, :3] = 235
    for k in range(0, size, 32):
        img[k:k + 3, :, :3] = rng.integers(0, 90, 3, dtype=np.uint8)
        img[:, k:k + 3, :3] = rng.integers(0, 90, 3, dtype=np.uint8)
    yy, xx = np.mgrid[0:size, 0:size]
    sun = ((xx - 0.7 * size) ** 2 + (yy - 0.25 * size) ** 2) < (0.04 * size) ** 2
    img[sun, :3] = 255
    return img


def spheres_prims(dev, kind="glass", num=50):
    tex_pixels = lines_image()
    tex, img = texture(dev, tex_pixels)
    if kind == "glass":                                 # <code>"glass"</code> transmission 1 1 1, etaOutside 1, etaInside 1.45
        m = dev.rtNewMaterial("glass")
        dev.rtSetFloat3(m, "transmission", 1, 1, 1); dev.rtSetFloat1(m, "etaOutside", 1.0); dev.rtSetFloat1(m, "etaInside", 1.45)
    else:                                               # <code>"mirror"</code> with the ignored "reflectivity" key
        m = dev.rtNewMaterial("mirror")
        dev.rtSetFloat3(m, "reflectivity", 1, 1, 1)
    dev.rtCommit(m)
    sph = dev.rtNewShape("sphere")                      # xml_loader.cpp:471-488
    dev.rtSetFloat3(sph, "P", 0, 100, 0); dev.rtSetFloat3(sph, "dPdt", 0, 0, 0); dev.rtSetFloat1(sph, "r", 100.0)
    dev.rtSetInt1(sph, "numTheta", num); dev.rtSetInt1(sph, "numPhi", num)
    dev.rtCommit(sph)
    floor_mat = dev.rtNewMaterial("MatteTextured")
    dev.rtSetTexture(floor_mat, "Kd", tex); dev.rtSetFloat2(floor_mat, "s0", 0, 0); dev.rtSetFloat2(floor_mat, "ds", 1, 1)
    dev.rtCommit(floor_mat)
    floor = add_mesh(dev, [(-1000, 0, -1000), (1000, 0, -1000), (1000, 0, 1000), (-1000, 0, 1000)], [(0, 1, 2), (2, 3, 0)],
                     normals=[(0, 1, 0)] * 4, uvs=[(0, 0), (1, 0), (1, 1), (0, 1)])
    hdri = dev.rtNewLight("hdrilight")                  # xml_loader.cpp:383-393
    dev.rtSetTransform(hdri, "local2world", np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], F))
    dev.rtSetFloat3(hdri, "L", 2.0, 1.5, 1.2)
    dev.rtSetImage(hdri, "image", img)
    dev.rtCommit(hdri)
    ident = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], F)
    return [dev.rtNewShapePrimitive(sph, m, ident), dev.rtNewShapePrimitive(floor, floor_mat, ident),
            dev.rtNewLightPrimitive(hdri, None, ident)]


def spheres(dev, kind="glass", width=1024, height=1024, spp=64, depth=8, face=None, fmt="RGB_FLOAT32", num=50, **kw):
    """sphere_view.ecs: -vp -200 100 200 -vi 0 100 200 -vu 0 1 0 -fov 90 -depth 8 -ambientlight 1 0 0 -stereo."""
    prims = spheres_prims(dev, kind, num) + [ambient_light(dev, (1, 0, 0))]
    pos, target, up = (-200, 100, 200), (0, 100, 200), (0, 1, 0)
    cam = pinhole(dev, pos, target, up, 90.0, width / height) if face is None else stereo_camera(dev, face, pos, target, up)
    return _bundle(dev, prims, cam, pathtracer(dev, spp, depth, **kw), width, height, fmt, view=(pos, target, up))


# ------------------------------------------------------------------------------------------------
# C3/C4 stand-in ("atrium"): a procedural architectural interior in the style of the missing Collada scenes — tessellated
# floor/walls/ceiling with a skylight, a colonnade, textured Uber materials (incl. RGBA cut-outs with alpha), one
# thin-glass pane, one reflective Uber, back-face-culled single-sided meshes, a camera-facing billboard and the dome
# (ambient) light with a finite tMaxShadowRay. Procedural textures replace the stripped image sets.
# ------------------------------------------------------------------------------------------------
def grid_quad(origin, du, dv, nu, nv, uv_scale=(1.0, 1.0)):
    """Tessellated parallelogram origin + s*du + t*dv, CCW seen from cross(du, dv)."""
    origin, du, dv = (np.asarray(v, np.float64) for v in (origin, du, dv))
    s, t = np.meshgrid(np.linspace(0, 1, nu + 1), np.linspace(0, 1, nv + 1), indexing="xy")
    pos = origin + s[..., None] * du + t[..., None] * dv
    n = np.cross(du, dv); n /= np.linalg.norm(n)
    uv = np.stack([s * uv_scale[0], t * uv_scale[1]], -1)
    idx = np.arange((nu + 1) * (nv + 1)).reshape(nv + 1, nu + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel(), idx[1:, :-1].ravel()
    tri = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)])
    return pos.reshape(-1, 3).astype(F), np.tile(n, (pos.shape[0] * pos.shape[1], 1)).astype(F), uv.reshape(-1, 2).astype(F), tri.astype(np.int32)


def merge(parts):
    pos, nrm, uv, tri, base = [], [], [], [], 0
    for p, n, u, t in parts:
        pos.append(p); nrm.append(n); uv.append(u); tri.append(t + base); base += len(p)
    return np.concatenate(pos), np.concatenate(nrm), np.concatenate(uv), np.concatenate(tri)


def box_parts(lo, hi, n, uv_scale=(1.0, 1.0)):
    """Six outward-facing tessellated faces of an axis-aligned box."""
    x0, y0, z0 = lo; x1, y1, z1 = hi
    dx, dy, dz = x1 - x0, y1 - y0, z1 - z0
    return [grid_quad((x0, y0, z1), (dx, 0, 0), (0, dy, 0), n, n, uv_scale),      # +z
            grid_quad((x1, y0, z0), (-dx, 0, 0), (0, dy, 0), n, n, uv_scale),     # -z
            grid_quad((x1, y0, z1), (0, 0, -dz), (0, dy, 0), n, n, uv_scale),     # +x
            grid_quad((x0, y0, z0), (0, 0, dz), (0, dy, 0), n, n, uv_scale),      # -x
            grid_quad((x0, y1, z1), (dx, 0, 0), (0, 0, -dz), n, n, uv_scale),     # +y
            grid_quad((x0, y0, z0), (dx, 0, 0), (0, 0, dz), n, n, uv_scale)]      # -y


def checker_texture(size, seed, alpha_holes=False, channels=4):
    rng = np.random.default_rng(seed)
    base = rng.integers(60, 255, 3)
    yy, xx = np.mgrid[0:size, 0:size]
    cells = ((xx // (size // 8)) + (yy // (size // 8))) % 2
    noise = rng.integers(0, 40, (size, size, 1))
    rgb = np.clip(base[None, None, :] * (0.55 + 0.45 * cells[..., None]) + noise - 20, 0, 255).astype(np.uint8)
    if channels == 3:
        return rgb
    a = np.full((size, size, 1), 255, np.uint8)
    if alpha_holes:
        r2 = ((xx % (size // 4)) - size // 8) ** 2 + ((yy % (size // 4)) - size // 8) ** 2
        a[r2 < (size // 11) ** 2] = 0
        a[(r2 >= (size // 11) ** 2) & (r2 < (size // 9) ** 2)] = 128
    return np.concatenate([rgb, a], -1)


def uber(dev, tex=None, diffuse=None, roughness=None, reflectivity=None, eta=None):
    """Uber material as ColladaLoader.cpp:300-370 sets it up (keys: SURVEY Appendix A)."""
    m = dev.rtNewMaterial("Uber")
    if tex is not None:
        dev.rtSetTexture(m, "Kd", tex)
    if diffuse is not None:
        dev.rtSetFloat3(m, "diffuse", *diffuse)
    if roughness is not None:
        dev.rtSetFloat1(m, "roughness", roughness)
    if reflectivity is not None:
        dev.rtSetFloat1(m, "reflectivity", reflectivity)
    if eta is not None:
        dev.rtSetFloat1(m, "eta", eta)
    dev.rtCommit(m)
    return m


def atrium_prims(dev, detail=8, tex_size=256, tex_set=None, pixels_from=None):
    """About 12*detail^2*(columns+shell) triangles; detail=8 -> ~30 k, detail=24 -> ~270 k, detail=56 -> ~1.4 M.
    tex_set="sponza": the surfaces carry the real models/Sponza/*.JPG images of the data pack (read through rtNewImageFromFile);
    the two RGBA cut-outs stay procedural (the Sponza image set has no alpha channel)."""
    W, H, D = 2400.0, 900.0, 1600.0                       # centimetres, sceneScale 1
    if tex_set == "sponza":
        files = data_files("sponza")
        pick = lambda name: file_texture(dev, next(f for f in files if os.path.basename(f).lower().startswith(name)), pixels_from=pixels_from)[0]
        texs = [pick("kamen.jpg"), pick("01_stub"), pick("01_s_ba"), pick("reljef"), pick("kamen-stup"), pick("sp_luk")]
        tex_rgb = pick("x01_st")
        extra = [pick(n) for n in ("00_skap", "01_s_kap", "01_st_kp", "prozor1", "sky", "vrata_ko", "vrata_kr")]
    else:
        texs = [texture(dev, checker_texture(tex_size, 100 + i))[0] for i in range(6)]
        tex_rgb = texture(dev, checker_texture(tex_size, 300, channels=3))[0]
        extra = []
    tex_cut = texture(dev, checker_texture(tex_size, 200, alpha_holes=True))[0]
    prims = []

    def add(parts, mat, cull=False, xfm=None, face_camera=False):
        p, n, u, t = merge(parts)
        mesh = add_mesh(dev, p, t, normals=n, uvs=u, cull=cull, uv_stride12=True)
        prims.append(dev.rtNewShapePrimitive(mesh, mat, xfm, face_camera))

    n = detail
    # shell (inward-facing, single sided -> back-face culled like Collada single-sided meshes)
    add([grid_quad((0, 0, D), (W, 0, 0), (0, 0, -D), 3 * n, 2 * n, (6, 4))], uber(dev, texs[0], roughness=0.4), cull=True)      # floor (+y)
    add([grid_quad((0, 0, 0), (W, 0, 0), (0, H, 0), 3 * n, n, (6, 2))], uber(dev, texs[1]), cull=True)                          # -z wall faces +z
    add([grid_quad((W, 0, D), (-W, 0, 0), (0, H, 0), 3 * n, n, (6, 2))], uber(dev, texs[2]), cull=True)                         # +z wall faces -z
    add([grid_quad((0, 0, D), (0, 0, -D), (0, H, 0), 2 * n, n, (4, 2))], uber(dev, tex_rgb), cull=True)                         # -x wall faces +x
    add([grid_quad((W, 0, 0), (0, 0, D), (0, H, 0), 2 * n, n, (4, 2))], uber(dev, texs[3], reflectivity=0.35), cull=True)       # +x wall faces -x
    # ceiling with a skylight: four strips around the opening (face down)
    sx0, sx1, sz0, sz1 = 0.3 * W, 0.7 * W, 0.3 * D, 0.7 * D
    ceil = [grid_quad((0, H, 0), (W, 0, 0), (0, 0, sz0), 3 * n, n), grid_quad((0, H, sz1), (W, 0, 0), (0, 0, D - sz1), 3 * n, n),
            grid_quad((0, H, sz0), (sx0, 0, 0), (0, 0, sz1 - sz0), n, n), grid_quad((sx1, H, sz0), (W - sx1, 0, 0), (0, 0, sz1 - sz0), n, n)]
    add(ceil, uber(dev, diffuse=(0.8, 0.8, 0.78)), cull=True)
    # colonnade: two rows of box columns (double sided)
    col_mat = uber(dev, texs[4], roughness=0.0)
    cols = []
    for i in range(6):
        x = 300.0 + i * 360.0
        for z in (350.0, D - 350.0):
            cols += box_parts((x - 40, 0, z - 40), (x + 40, H, z + 40), max(2, n // 2), (1, 6))
    add(cols, col_mat)
    # cut-out screens (RGBA alpha) between some columns, a thin glass pane, a matte default-material plinth
    screens = [grid_quad((300.0 + i * 360.0 + 40, 0, 350.0), (280, 0, 0), (0, 400, 0), n, n, (2, 3)) for i in (1, 3)]
    add(screens, uber(dev, tex_cut))
    glass = dev.rtNewMaterial("ThinDielectric")
    dev.rtSetFloat3(glass, "transmission", 0.9, 0.95, 0.9); dev.rtSetFloat1(glass, "eta", 1.5); dev.rtSetFloat1(glass, "thickness", 0.5)
    dev.rtCommit(glass)
    add([grid_quad((1000, 0, D - 350.0), (400, 0, 0), (0, 500, 0), 2, 2)], glass)
    matte = dev.rtNewMaterial("matte"); dev.rtSetFloat3(matte, "reflectance", 0.5, 0.5, 0.5); dev.rtCommit(matte)
    add(box_parts((1100, 0, 700), (1300, 120, 900), max(2, n // 2)), matte)
    for k, t in enumerate(extra):                          # the rest of the image set: hangings on the -z wall and plinths along +x
        if k < 4:
            add([grid_quad((250.0 + 520.0 * k, 250, 2.0), (300, 0, 0), (0, 400, 0), max(2, n // 4), max(2, n // 4))], uber(dev, t))
        else:
            add(box_parts((2200, 0, 300.0 + 300 * (k - 4)), (2320, 160, 420.0 + 300 * (k - 4)), max(2, n // 4)), uber(dev, t, roughness=0.6))
    # camera-facing billboard (YULIO_CAMERA_ALIGNED_ mesh, ColladaLoader.cpp:629-634): unit quad in its local xz plane
    bb = [grid_quad((-60, 0, -90), (120, 0, 0), (0, 0, 180), 1, 1)]
    xfm = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 1700, 100, 800], F)
    add(bb, uber(dev, tex_cut), xfm=xfm, face_camera=True)
    prims.append(ambient_light(dev, (0.83, 0.95, 0.98)))      # DLL default dome light (YulioRT.h:42)
    return prims


ATRIUM_VIEW = ((1250.0, 160.0, 820.0), (1250.0, 160.0, 0.0), (0.0, 1.0, 0.0))


def atrium(dev, width=1024, height=1024, spp=64, depth=10, face=0, detail=8, fmt="RGB_FLOAT32", tmax_shadow=120.0, tex_size=256,
           tex_set=None, pixels_from=None, **kw):
    prims = atrium_prims(dev, detail, tex_size, tex_set, pixels_from)
    pos, target, up = ATRIUM_VIEW
    cam = stereo_camera(dev, face, pos, target, up) if face is not None else pinhole(dev, pos, target, up, 90.0, width / height)
    s = _bundle(dev, prims, cam, pathtracer(dev, spp, depth, tmax_shadow=tmax_shadow, **kw), width, height, fmt, view=ATRIUM_VIEW)
    return s


# ------------------------------------------------------------------------------------------------
# C4 stand-in ("office"): the sample scene's .dae is stripped from the mount but its 152 texture images are not. A procedural
# open-plan office in the sample scene's units (inches, sceneScale 1) carries every one of them on an Uber / ThinDielectric
# material set up the way the Collada loader does (ColladaLoader.cpp:205-401: Kd texture, roughness = 1 - shininess,
# reflectivity; single-sided meshes are back-face culled, opacity < 1 -> ThinDielectric with eta 1.4, thickness 1): a culled shell
# with skylights and a window band, desks, cabinets and shelves (the 114 JPEGs), posters, screens and partitions (the 38 PNGs, 30 of
# them RGBA cut-outs), one YULIO_CAMERA_ALIGNED_ billboard, the dome light of the DLL defaults (YulioRT.h:37-50).
# ------------------------------------------------------------------------------------------------
OFFICE_VIEW = ((600.0, 63.0, 410.0), (600.0, 63.0, 0.0), (0.0, 1.0, 0.0))


def office_prims(dev, detail=40, pixels_from=None, max_textures=None):
    W, H, D = 1180.0, 157.0, 790.0                         # inches: 30 m x 4 m x 20 m
    files = data_files("sample_scene")
    if max_textures is not None:                            # small test scenes: a deterministic subset that keeps both kinds
        pngs = [f for f in files if f.lower().endswith(".png")][:max(2, max_textures // 4)]
        files = pngs + [f for f in files if not f.lower().endswith(".png")][:max_textures - len(pngs)]
    png = [f for f in files if f.lower().endswith(".png")]
    jpg = [f for f in files if not f.lower().endswith(".png")]
    big = lambda stem: next((f for f in jpg if os.path.basename(f).startswith(stem)), jpg[0])
    tex = lambda f: file_texture(dev, f, pixels_from=pixels_from)[0]
    prims = []
    used = set()

    def add(parts, mat, cull=False, xfm=None, face_camera=False):
        p, nn, u, t = merge(parts)
        mesh = add_mesh(dev, p, t, normals=nn, uvs=u, cull=cull, uv_stride12=True)
        prims.append(dev.rtNewShapePrimitive(mesh, mat, xfm, face_camera))

    def mat_for(f, k):
        used.add(f)
        # the loader's three Uber flavours, spread deterministically over the image set
        if k % 7 == 3:
            return uber(dev, tex(f), roughness=0.0)                       # shininess 1: perfect dielectric reflection lobe
        if k % 11 == 5:
            return uber(dev, tex(f), roughness=0.6, reflectivity=0.25)
        return uber(dev, tex(f), roughness=1.0 - 0.1 * (k % 5))

    n = detail
    # shell, single sided
    add([grid_quad((0, 0, D), (W, 0, 0), (0, 0, -D), 3 * n, 2 * n, (5, 4))], mat_for(big("PDM_Wood_floor_Cherry_01.jpg"), 0), cull=True)
    add([grid_quad((0, 0, 0), (W, 0, 0), (0, H, 0), 3 * n, n // 2, (8, 1))], mat_for(big("Brick_Antique_01"), 1), cull=True)
    add([grid_quad((0, 0, D), (0, 0, -D), (0, H, 0), 2 * n, n // 2, (6, 1))], mat_for(big("PDM_Concrete_09.jpg"), 2), cull=True)
    add([grid_quad((W, 0, 0), (0, 0, D), (0, H, 0), 2 * n, n // 2, (6, 1))], mat_for(big("PDM_Concrete_09_20"), 4), cull=True)
    # +z wall: a parapet and a lintel around a window band closed by thin glass
    add([grid_quad((W, 0, D), (-W, 0, 0), (0, 40, 0), 3 * n, max(2, n // 8), (8, .3)),
         grid_quad((W, 120, D), (-W, 0, 0), (0, H - 120, 0), 3 * n, max(2, n // 8), (8, .3))], mat_for(big("Metal_Corrogated_Brown"), 6), cull=True)
    glass = dev.rtNewMaterial("ThinDielectric")
    dev.rtSetFloat3(glass, "transmission", 0.9, 0.95, 0.9); dev.rtSetFloat1(glass, "eta", 1.4); dev.rtSetFloat1(glass, "thickness", 1.0)
    dev.rtSetFloat1(glass, "transparency", 0.85); dev.rtCommit(glass)
    add([grid_quad((W, 40, D), (-W, 0, 0), (0, 80, 0), 8, 2)], glass)
    # ceiling with two skylights
    ceil = []
    for (x0, x1) in ((0, 250), (450, 730), (930, W)):
        ceil.append(grid_quad((x0, H, 0), (x1 - x0, 0, 0), (0, 0, D), max(2, int(n * (x1 - x0)) // 400), 2 * n))
    for (x0, x1) in ((250, 450), (730, 930)):
        ceil.append(grid_quad((x0, H, 0), (x1 - x0, 0, 0), (0, 0, 250), n // 2, n // 2))
        ceil.append(grid_quad((x0, H, 540), (x1 - x0, 0, 0), (0, 0, D - 540), n // 2, n // 2))
    add(ceil, mat_for(big("Carpet_Plush_Charcoal"), 8), cull=True)
    # columns
    cols = []
    for i in range(4):
        for z in (200.0, D - 200.0):
            cols += box_parts((150.0 + i * 290.0 - 9, 0, z - 9), (150.0 + i * 290.0 + 9, H, z + 9), max(2, n // 4), (1, 6))
    add(cols, mat_for(big("PDM_Concrete_09.jpg"), 3))
    rest = [f for f in jpg if f not in used]
    # desks (9 x 5), each with its own image on a slab and two legs panels; a PNG "screen" stands on most of them
    m = max(2, n // 5)
    k = 0
    screens = list(png)
    billboard_png = next((f for f in png if os.path.basename(f).startswith("material_20")), png[0])
    if billboard_png in screens and len(screens) > 1:
        screens.remove(billboard_png)
    for iz in range(5):
        for ix in range(9):
            x, z = 90.0 + ix * 118.0, 110.0 + iz * 128.0
            if rest:
                f = rest.pop(0)
                parts = box_parts((x, 27.5, z), (x + 60, 29, z + 30), m) + box_parts((x + 1, 0, z + 1), (x + 3, 27.5, z + 29), max(2, m // 2)) \
                    + box_parts((x + 57, 0, z + 1), (x + 59, 27.5, z + 29), max(2, m // 2))
                add(parts, mat_for(f, k))
            if screens:
                f = screens.pop(0); used.add(f)
                add([grid_quad((x + 15, 29, z + 15), (30, 0, 0), (0, 20, 0), m, m)], uber(dev, tex(f), roughness=0.2 if k % 2 else 0.9))
            k += 1
    # cabinets along the -z and -x walls, shelves along +x
    slots = [((40.0 + i * 56.0, 0, 4), (40.0 + i * 56.0 + 40, 72, 24)) for i in range(20)] + \
            [((4, 0, 60.0 + i * 44.0), (24, 48, 60.0 + i * 44.0 + 36)) for i in range(15)] + \
            [((W - 22, 30.0, 40.0 + i * 36.0), (W - 4, 60, 40.0 + i * 36.0 + 30)) for i in range(20)]
    for lo, hi in slots:
        if not rest:
            break
        f = rest.pop(0)
        add(box_parts(lo, hi, m), mat_for(f, k)); k += 1
    # whatever is left of the image set hangs on the walls as single-sided posters
    i = 0
    for f in rest + screens:
        used.add(f)
        x = 60.0 + (i % 24) * 46.0
        y = 84.0 + 34.0 * (i // 24)
        add([grid_quad((x, y, 1.0), (36, 0, 0), (0, 28, 0), max(2, m // 2), max(2, m // 2))], uber(dev, tex(f)), cull=True); i += 1
    # camera-facing billboard (YULIO_CAMERA_ALIGNED_ mesh, ColladaLoader.cpp:629-634): a quad in its local xz plane
    bb = [grid_quad((-18, 0, -33), (36, 0, 0), (0, 0, 66), 1, 1)]
    xfm = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 760, 34, 300], F)
    used.add(billboard_png)
    add(bb, uber(dev, tex(billboard_png)), xfm=xfm, face_camera=True)
    prims.append(ambient_light(dev, (0.83, 0.95, 0.98)))      # DLL default dome light (YulioRT.h:42)
    return prims, len(used)


def office(dev, width=2048, height=2048, spp=16, depth=10, face=0, detail=40, fmt="RGB_FLOAT32", tmax_shadow=120.0, pixels_from=None,
           max_textures=None, **kw):
    """BASELINE configs[3] stand-in with the DLL defaults (YulioRT.h:37-50): depth 10, tMaxShadowRay 120, dome (.83,.95,.98), eye
    separation 2.5 in, toe-in with zero parallax at 75 in (ColladaLoader.cpp:492-496)."""
    prims, ntex = office_prims(dev, detail, pixels_from, max_textures)
    pos, target, up = OFFICE_VIEW
    stereo = dict(toe_in=True, eye_separation=6.35 * 0.393701, zero_parallax=6.35 * 0.393701 * 30.0, scene_scale=1.0)
    cam = stereo_camera(dev, face, pos, target, up, **stereo) if face is not None else pinhole(dev, pos, target, up, 90.0, width / height)
    s = _bundle(dev, prims, cam, pathtracer(dev, spp, depth, tmax_shadow=tmax_shadow, **kw), width, height, fmt, view=OFFICE_VIEW, stereo=stereo,
                num_textures=ntex)
    return s


def cube_cameras(dev, s, faces=range(12), toe_in=None):
    """The stereo cube cameras of the scene's viewpoint (ColladaLoader.cpp:470-505), with the scene's stereo parameters if it has any."""
    pos, target, up = s.view
    kw = dict(getattr(s, "stereo", {}) or {})
    if toe_in is not None:
        kw["toe_in"] = toe_in
    return [stereo_camera(dev, i, pos, target, up, **kw) for i in faces]


def render_cube_map_batched(dev, s, cams, framebuffers, update_prims=True):
    """One viewpoint through yrtxRenderCubeMap: update the camera-aligned primitives once (the 12 cameras share their origin), commit,
    render all faces as one wavefront."""
    if update_prims:
        org = dev.rtGetFloat3(cams[0], "origin")
        for j, p in enumerate(s.prims):
            dev.rtUpdatePrimitive(s.scene, j, p, org, s.view[2])
        dev.rtCommit(s.scene)
    dev.render_cube_map(s.renderer, cams, s.scene, s.tonemapper, framebuffers, 0)


def render_cube_map(dev, s, faces=range(12), update_prims=True, toe_in=False):
    """The per-viewpoint loop of outputMode (devices/renderer/renderer.cpp:543-632): for every cube face, update all
    primitives against the face camera's origin, commit the scene, render, swap. Yields (face, camera)."""
    pos, target, up = s.view
    for i in faces:
        cam = stereo_camera(dev, i, pos, target, up, toe_in=toe_in)
        if update_prims:
            org = dev.rtGetFloat3(cam, "origin")
            for j, p in enumerate(s.prims):
                dev.rtUpdatePrimitive(s.scene, j, p, org, up)
            dev.rtCommit(s.scene)
        dev.rtRenderFrame(s.renderer, cam, s.scene, s.tonemapper, s.framebuffer, 0)
        dev.rtSwapBuffers(s.framebuffer)
        yield i, cam


# ------------------------------------------------------------------------------------------------
# C5: synthetic triangle soup: blobs of small triangles, centres uniform in a cube; random ray segments
# ------------------------------------------------------------------------------------------------
def soup_arrays(num_tris, seed=12345, extent=1000.0, blob_tris=1000, blob_radius=2.0, edge=0.1):
    rng = np.random.default_rng(seed)
    nblob = max(1, num_tris // blob_tris)
    centres = rng.uniform(0, extent, (nblob, 3))
    owner = rng.integers(0, nblob, num_tris) if num_tris % blob_tris else np.repeat(np.arange(nblob), blob_tris)
    c = centres[owner] + rng.normal(0, blob_radius / 2, (num_tris, 3))
    v = c[:, None, :] + rng.uniform(-edge, edge, (num_tris, 3, 3))
    pos = v.reshape(-1, 3).astype(F)
    tri = np.arange(num_tris * 3, dtype=np.int32).reshape(-1, 3)
    return pos, tri


def soup(dev, num_tris, seed=12345, extent=1000.0, meshes=1, cull=False, edge=0.1, blob_radius=2.0):
    pos, tri = soup_arrays(num_tris, seed, extent, blob_radius=blob_radius, edge=edge)
    matte = dev.rtNewMaterial("matte"); dev.rtCommit(matte)
    prims = []
    per = (num_tris + meshes - 1) // meshes
    for k in range(meshes):
        t = tri[k * per:(k + 1) * per]
        if len(t) == 0:
            continue
        p = pos[t[0, 0]:t[-1, 2] + 1]
        prims.append(dev.rtNewShapePrimitive(add_mesh(dev, p, t - t[0, 0], cull=cull), matte, None))
    s = SimpleNamespace(prims=prims, extent=extent)
    s.scene = make_scene(dev, prims)
    return s


def random_rays(n, seed=67890, extent=1000.0, tfar=math.inf, tfar_uniform=None):
    rng = np.random.default_rng(seed)
    rays = np.zeros((n, 8), F)
    rays[:, 0:3] = rng.uniform(0, extent, (n, 3))
    d = rng.normal(0, 1, (n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 4:7] = d
    rays[:, 7] = tfar if tfar_uniform is None else rng.uniform(0, tfar_uniform, n)
    return rays


# ------------------------------------------------------------------------------------------------
# The materials reachable only through .xml scenes / the regression driver: a Cornell box whose blocks, floor and
# back wall carry Plastic, Metal (rough + polished), BrushedMetal, MetallicPaint and Velvet, lit by the quad light and a dome light.
# Meshes carry normals and texture coordinates so that both tangent constructions are exercised (trianglemesh_full.cpp:252-270,
# trianglemesh_normals.cpp:154-155).
# ------------------------------------------------------------------------------------------------
def material(dev, kind, **params):
    m = dev.rtNewMaterial(kind)
    for k, v in params.items():
        if isinstance(v, (tuple, list)):
            dev.rtSetFloat3(m, k, *v)
        else:
            dev.rtSetFloat1(m, k, float(v))
    dev.rtCommit(m)
    return m


def showroom(dev, width=64, height=64, spp=16, depth=5, fmt="RGB_FLOAT32", **kw):
    mats = {
        "white": material(dev, "Velvet", reflectance=(.6, .3, .35), backScattering=.7, horizonScatteringColor=(.4, .4, .5), horizonScatteringFallOff=6.0),
        "red": material(dev, "Plastic", pigmentColor=(.8, .1, .1), eta=1.5, roughness=.05),
        "green": material(dev, "MetallicPaint", shadeColor=(.1, .5, .2), glitterColor=(.8, .8, .6), glitterSpread=.3, eta=1.45),
        "blue": material(dev, "Metal", reflectance=(.9, .8, .5), eta=(.2, .9, 1.1), k=(3.9, 2.4, 2.2), roughness=.08),
    }
    extra = [material(dev, "BrushedMetal", reflectance=(.8, .8, .9), eta=(1.4, 1.2, 1.1), k=(5.0, 4.6, 4.2), roughnessX=.02, roughnessY=.3),
             material(dev, "Metal", reflectance=(1, 1, 1), eta=(.6, .6, .6), k=(4.8, 4.8, 4.8), roughness=0.0),
             material(dev, "Plastic", pigmentColor=(.2, .3, .9), roughness=0.0)]
    prims = []
    for gi, (mat, quads) in enumerate(_CORNELL_GROUPS):
        for qi, q in enumerate(quads):
            p = np.asarray(q, np.float64)
            n = np.cross(p[1] - p[0], p[3] - p[0]); n /= np.linalg.norm(n)
            uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float64) * (1 + qi)
            m = mats[mat] if not (mat == "white" and gi > 0) else extra[(gi + qi) % len(extra)]
            with_uv = (qi % 2 == 0)
            mesh = add_mesh(dev, p, [(0, 1, 2), (0, 2, 3)], normals=np.tile(n, (4, 1)), uvs=uv if with_uv else None)
            prims.append(dev.rtNewShapePrimitive(mesh, m, None))
    prims += quad_light(dev, (213, 548.77, 227), (130, 0, 0), (0, 0, 105), (40, 40, 40)) + [ambient_light(dev, (.3, .35, .4))]
    cam = pinhole(dev, (278, 273, -800), (278, 273, 0), (0, 1, 0), 37.0, width / height)
    return _bundle(dev, prims, cam, pathtracer(dev, spp, depth, **kw), width, height, fmt)


# ------------------------------------------------------------------------------------------------
# The reference's randomised API stress (devices/renderer/regression.cpp:32-226): one ambient light, random objects — tiny meshes with
# random, possibly degenerate index triples (positions with stride 16, one texcoord per vertex) or tessellated spheres — with random
# materials of its 8 kinds and 32x32 "x*y" textures (RGB8 or float). Seeded here (the reference uses unseeded rand()). The light
# primitive is placed after the shapes: the reference's flat scene indexes its primitive list with geomIDs counted over shapes only
# (api/scene_flat.h:99-105,140-142), so its own order (light first) dereferences the light's null shape on the first hit.
# ------------------------------------------------------------------------------------------------
def _regression_image(dev, rng, width=32, height=32):
    y, x = np.mgrid[0:height, 0:width]
    px = np.stack([(x * y), (y * x), (x + y)], -1).astype(np.int64)
    as_char = ((px + 128) % 256 - 128).astype(np.int8)                  # char(x * y): signed on the reference's platforms
    if rng.random() < 0.5:
        return dev.rtNewImage("RGB8", width, height, np.ascontiguousarray(as_char.view(np.uint8)))
    return dev.rtNewImage("RGB_FLOAT32", width, height, np.ascontiguousarray(as_char.astype(F) / F(255.0)))


def _regression_material(dev, rng):
    r = lambda: float(F(rng.random()))
    k = int(rng.integers(0, 8))
    if k == 0:
        return material(dev, "Matte", reflectance=(r(), r(), r()))
    if k == 1:
        return material(dev, "Plastic", pigmentColor=(r(), r(), r()), eta=1.0 + r(), roughness=0.1 * r())
    if k == 2:
        return material(dev, "Dielectric", transmission=(.5 * r() + .5, .5 * r() + .5, .5 * r() + .5), etaOutside=1.0, etaInside=1.0 + r())
    if k == 3:
        return material(dev, "ThinDielectric", transmission=(.5 * r() + .5, .5 * r() + .5, .5 * r() + .5), eta=1.0 + r(), thickness=.5 * r())
    if k == 4:
        return material(dev, "Mirror", reflectance=(.5 * r() + .5, .5 * r() + .5, .5 * r() + .5))
    if k == 5:
        return material(dev, "Metal", reflectance=(.5 * r() + .5, .5 * r() + .5, .5 * r() + .5), eta=(1 + r(), 1 + r(), 1 + r()),
                        k=(.3 * r(), .3 * r(), .3 * r()), roughness=.3 * r())
    if k == 6:
        return material(dev, "MetallicPaint", shadeColor=(.5 * r() + .5, .5 * r() + .5, .5 * r() + .5), glitterColor=(r(), r(), r()),
                        glitterSpread=.5 + r(), eta=1.0 + r())
    m = dev.rtNewMaterial("MatteTextured")
    t = dev.rtNewTexture("image")
    dev.rtSetImage(t, "image", _regression_image(dev, rng)); dev.rtCommit(t)
    dev.rtSetTexture(m, "Kd", t)
    dev.rtSetFloat2(m, "s0", r(), r()); dev.rtSetFloat2(m, "ds", 5 * r(), 5 * r())
    dev.rtCommit(m)
    return m


def _regression_shape(dev, rng, num_triangles):
    if num_triangles < 20:
        n = num_triangles
        pos = 2.0 * rng.random(3) - 1.0
        positions = np.zeros((n, 4), F); positions[:, :3] = pos + 0.3 * rng.random((n, 3))
        texcoords = rng.random((n, 2)).astype(F)
        indices = rng.integers(0, max(n, 1), (n, 3)).astype(np.int32)
        mesh = dev.rtNewShape("trianglemesh")
        dp, dt, di = dev.rtNewData("immutable", positions), dev.rtNewData("immutable", texcoords), dev.rtNewData("immutable", indices)
        dev.rtSetArray(mesh, "positions", "float3", dp, n, 16, 0)
        dev.rtSetArray(mesh, "texcoords", "float2", dt, n, 8, 0)
        dev.rtSetArray(mesh, "indices", "int3", di, n, 12, 0)
        dev.rtCommit(mesh)
        return mesh
    s = dev.rtNewShape("sphere")
    dev.rtSetFloat3(s, "P", *(float(v) for v in 2.0 * rng.random(3) - 1.0))
    dev.rtSetFloat1(s, "r", 0.2 * float(rng.random()))
    dev.rtSetInt1(s, "numTheta", num_triangles // 20); dev.rtSetInt1(s, "numPhi", 20)
    dev.rtCommit(s)
    return s


def regression(dev, seed, num_objects=10, num_triangles=60, width=32, height=32, spp=4, depth=4, fmt="RGB_FLOAT32"):
    rng = np.random.default_rng(seed)
    prims = []
    for _ in range(num_objects):
        s = int(rng.integers(0, num_triangles)) if num_triangles else 0
        prims.append(dev.rtNewShapePrimitive(_regression_shape(dev, rng, s), _regression_material(dev, rng), None))
    prims.append(ambient_light(dev, (1.0, 1.0, 1.0)))
    cam = pinhole(dev, (0.2, 0.3, -2.6), (0, 0, 0), (0, 1, 0), 50.0, width / height)
    return _bundle(dev, prims, cam, pathtracer(dev, spp, depth), width, height, fmt)
