"""Writes a small COLLADA 1.4.1 scene (+ a PPM texture next to it) that exercises what the reference's Collada front end reads
(devices/device/loaders/ColladaLoader.cpp): a camera node tagged YULIO_FPR_VIEW_ (-> 12 stereo cube cameras, :403-500), a scaled
camera node (-> sceneScale, :441-446), a diffuse-textured effect and a plain diffuse effect (-> Uber, :218-236,349-366), an effect
with <transparency> < 1 (-> ThinDielectric, :278-296), a double-sided effect (-> no back-face culling, :327-338) and a geometry
named YULIO_CAMERA_ALIGNED_* (-> faceCamera billboard, :629-634). The real sample scenes are stripped from the reference mount
(SURVEY F3); this is the stand-in that goes through the same loader. Procedural, seeded; reads nothing from /root/reference."""
from __future__ import annotations

import os

import numpy as np


def _floats(a) -> str:
    return " ".join(repr(float(np.float32(v))) for v in np.asarray(a).ravel())


def _geometry(gid, name, pos, nrm, uv, tris) -> str:
    pos, nrm, uv = np.asarray(pos, np.float32), np.asarray(nrm, np.float32), np.asarray(uv, np.float32)
    tris = np.asarray(tris, np.int32).reshape(-1, 3)
    p = " ".join(f"{i} {i} {i}" for i in tris.ravel())

    def source(sid, arr, names):
        n = arr.shape[1]
        params = "".join(f'<param name="{c}" type="float"/>' for c in names)
        return (f'<source id="{gid}-{sid}"><float_array id="{gid}-{sid}-a" count="{arr.size}">{_floats(arr)}</float_array>'
                f'<technique_common><accessor source="#{gid}-{sid}-a" count="{len(arr)}" stride="{n}">{params}</accessor></technique_common></source>')

    return (f'<geometry id="{gid}" name="{name}"><mesh>{source("pos", pos, "XYZ")}{source("nrm", nrm, "XYZ")}{source("uv", uv, "ST")}'
            f'<vertices id="{gid}-vtx"><input semantic="POSITION" source="#{gid}-pos"/></vertices>'
            f'<triangles count="{len(tris)}" material="m"><input semantic="VERTEX" source="#{gid}-vtx" offset="0"/>'
            f'<input semantic="NORMAL" source="#{gid}-nrm" offset="1"/><input semantic="TEXCOORD" source="#{gid}-uv" offset="2" set="0"/>'
            f'<p>{p}</p></triangles></mesh></geometry>')


def _quad(o, du, dv, n=2):
    """(n x n)-cell quad with normals du x dv and uvs in [0,1]^2."""
    o, du, dv = (np.asarray(v, np.float64) for v in (o, du, dv))
    s = np.linspace(0.0, 1.0, n + 1)
    pos = np.array([o + a * du + b * dv for b in s for a in s])
    uv = np.array([[a, b] for b in s for a in s])
    nn = np.cross(du, dv); nn /= np.linalg.norm(nn)
    nrm = np.tile(nn, (len(pos), 1))
    tris = []
    for j in range(n):
        for i in range(n):
            a = j * (n + 1) + i
            tris += [[a, a + 1, a + n + 2], [a, a + n + 2, a + n + 1]]
    return pos, nrm, uv, np.array(tris)


def _box(lo, hi, inward=False):
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    d = hi - lo
    ex, ey, ez = np.array([d[0], 0, 0]), np.array([0, d[1], 0]), np.array([0, 0, d[2]])
    faces = [(lo, ez, ey), (lo + ex, ey, ez), (lo, ex, ez), (lo + ey, ez, ex), (lo, ey, ex), (lo + ez, ex, ey)]  # outward normals
    P, N, U, T = [], [], [], []
    base = 0
    for o, a, b in faces:
        if inward:
            a, b = b, a
        p, n, u, t = _quad(o, a, b, 2)
        P.append(p); N.append(n); U.append(u); T.append(t + base); base += len(p)
    return np.concatenate(P), np.concatenate(N), np.concatenate(U), np.concatenate(T)


def _effect(eid, diffuse=None, texture=None, transparency=None, double_sided=False, shininess=None, reflectivity=None) -> str:
    params = ""
    if texture is not None:
        params = (f'<newparam sid="{eid}-surf"><surface type="2D"><init_from>{texture}</init_from></surface></newparam>'
                  f'<newparam sid="{eid}-samp"><sampler2D><source>{eid}-surf</source></sampler2D></newparam>')
        dif = f'<diffuse><texture texture="{eid}-samp" texcoord="UVSET0"/></diffuse>'
    else:
        dif = f'<diffuse><color>{_floats(list(diffuse) + [1.0])}</color></diffuse>'
    extra = ""
    if transparency is not None:
        extra += f'<transparent opaque="A_ONE"><color>1 1 1 1</color></transparent><transparency><float>{transparency}</float></transparency>'
    if reflectivity is not None:
        extra += f'<reflective><color>1 1 1 1</color></reflective><reflectivity><float>{reflectivity}</float></reflectivity>'
    shin = f'<specular><color>0.2 0.2 0.2 1</color></specular><shininess><float>{shininess}</float></shininess>' if shininess is not None else ""
    shader = "phong" if shininess is not None else "lambert"
    ds = '<extra><technique profile="GOOGLEEARTH"><double_sided>1</double_sided></technique></extra>' if double_sided else ""
    return (f'<effect id="{eid}"><profile_COMMON>{params}<technique sid="common"><{shader}>{dif}{shin}{extra}</{shader}></technique>{ds}'
            f'</profile_COMMON></effect>')


def _matrix(scale=1.0, t=(0, 0, 0)) -> str:
    m = np.eye(4); m[0, 0] = m[1, 1] = m[2, 2] = scale; m[:3, 3] = t
    return _floats(m)


def write_scene(directory: str, name: str = "room", views=(("A", (0.0, 60.0, 0.0)),), scene_scale: float = 1.0, tex_size: int = 64, seed: int = 3) -> str:
    """Returns the path of <directory>/<name>.dae. One cube-map viewpoint per entry of `views` (name suffix, position)."""
    os.makedirs(directory, exist_ok=True)
    rng = np.random.default_rng(seed)
    # texture: seeded checker with noise, binary PPM (the only 8-bit format both back ends decode without third-party codecs)
    t = np.zeros((tex_size, tex_size, 3), np.uint8)
    yy, xx = np.mgrid[0:tex_size, 0:tex_size]
    chk = ((xx // 8 + yy // 8) % 2).astype(np.uint8)
    t[..., 0] = 60 + 150 * chk; t[..., 1] = 200 - 120 * chk; t[..., 2] = 90
    t = np.clip(t.astype(np.int32) + rng.integers(-20, 20, t.shape), 0, 255).astype(np.uint8)
    with open(os.path.join(directory, "checker.ppm"), "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (tex_size, tex_size)); f.write(t.tobytes())

    geos, nodes, mats, fx = [], [], [], []

    def add(gid, gname, geo, effect, xform=None):
        geos.append(_geometry(gid, gname, *geo))
        fx.append(effect)
        eid = effect.split('"')[1]
        mats.append(f'<material id="{eid}-mat" name="{eid}-mat"><instance_effect url="#{eid}"/></material>')
        nodes.append(f'<node id="{gid}-node" name="{gname}"><matrix>{xform or _matrix()}</matrix><instance_geometry url="#{gid}">'
                     f'<bind_material><technique_common><instance_material symbol="m" target="#{eid}-mat">'
                     f'<bind_vertex_input semantic="UVSET0" input_semantic="TEXCOORD" input_set="0"/></instance_material>'
                     f'</technique_common></bind_material></instance_geometry></node>')

    add("g0", "shell", _box((-200, 0, -200), (200, 160, 200), inward=True), _effect("fx0", texture="img0"))
    add("g1", "block", _box((-120, 0, -130), (-60, 90, -70)), _effect("fx1", diffuse=(0.7, 0.3, 0.2), shininess=0.3))
    add("g2", "pillar", _box((70, 0, 60), (110, 150, 100)), _effect("fx2", diffuse=(0.2, 0.5, 0.8), double_sided=True))
    add("g3", "pane", _quad((-40, 10, 120), (120, 0, 0), (0, 110, 0), 2), _effect("fx3", diffuse=(0.9, 0.95, 1.0), transparency=0.35, double_sided=True))
    add("g4", "YULIO_CAMERA_ALIGNED_tree", _quad((-25, 0, 0), (50, 0, 0), (0, 0, -90), 1), _effect("fx4", texture="img0", double_sided=True),
        _matrix(1.0, (120, 0, -120)))
    add("g5", "shiny", _box((-30, 0, -160), (40, 50, -110)), _effect("fx5", diffuse=(0.8, 0.8, 0.8), shininess=1.0, reflectivity=0.4))

    cams, camnodes = [], []
    for i, (suffix, pos) in enumerate(views):
        cams.append(f'<camera id="cam{i}" name="YULIO_FPR_VIEW_{suffix}"><optics><technique_common><perspective><yfov>60</yfov>'
                    f'<aspect_ratio>1</aspect_ratio><znear>1</znear><zfar>10000</zfar></perspective></technique_common></optics></camera>')
        camnodes.append(f'<node id="cam{i}-node" name="YULIO_FPR_VIEW_{suffix}"><matrix>{_matrix(scene_scale, pos)}</matrix>'
                        f'<instance_camera url="#cam{i}"/></node>')
    doc = ('<?xml version="1.0" encoding="utf-8"?>\n<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">'
           '<asset><unit name="inch" meter="0.0254"/><up_axis>Y_UP</up_axis></asset>'
           f'<library_cameras>{"".join(cams)}</library_cameras>'
           '<library_images><image id="img0" name="img0"><init_from>checker.ppm</init_from></image></library_images>'
           f'<library_effects>{"".join(fx)}</library_effects><library_materials>{"".join(mats)}</library_materials>'
           f'<library_geometries>{"".join(geos)}</library_geometries>'
           f'<library_visual_scenes><visual_scene id="vs" name="vs">{"".join(nodes)}{"".join(camnodes)}</visual_scene></library_visual_scenes>'
           '<scene><instance_visual_scene url="#vs"/></scene></COLLADA>\n')
    path = os.path.join(directory, name + ".dae")
    with open(path, "w") as f:
        f.write(doc)
    return path


def read_ppm(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        data = f.read()
    parts = data.split(b"\n", 3)
    assert parts[0] == b"P6", parts[0]
    w, h = (int(v) for v in parts[1].split())
    return np.frombuffer(parts[3], np.uint8, w * h * 3).reshape(h, w, 3)
