"""Parity at the benchmark sizes (BASELINE.json configs[1..3]) and independent geometric truth.

1. Primary-hit records of ALL 12 stereo cube faces of the C2, C3 and C4 workloads at their full face size (1024^2 / 1024^2 / 2048^2,
   one sample per pixel, exactly the rays yrtRenderFrame generates) against the oracle's rtcIntersect on the SAME rays:
   geomID / primID bit-exact, t / u / v / Ng bit-exact. The mismatch fraction is printed (and asserted to be 0).
2. The oracle and the GPU share one arithmetic contract at the Embree boundary (YRT-PLUECKER-1), so (1) cannot tell whether both are
   geometrically right. An independent brute force settles it: every ray against every triangle in float64 (Moeller-Trumbore on the GPU with
   torch, no BVH, no code shared with the product), nearest hit. Stated bound: the GPU's (geomID, primID) equals the brute-force nearest
   triangle except where the two nearest candidates lie within 1e-5 relative distance of each other (edge / vertex ties between
   neighbouring triangles of a tessellated surface); |t - t64| <= 4 ulp(max(|t|, scene scale)). The tie fraction is printed.
"""
import numpy as np
import pytest

import bench
from tests import scenes
from tests.test_gpu_parity import ids

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wl", ["c2", "c3", "c4"])
def test_primary_hits_on_every_cube_face_at_benchmark_size(wl, oracle_dev):
    oracle_mt = oracle_dev      # the oracle's yrtxTraceRays spreads large batches over the host threads itself (oracle_capi.cpp)
    from yulio_raytracer_b200 import Device
    _, _, size, _, depth, faces = bench.WORKLOADS[wl]
    dev = Device.cuda()
    sg = bench.build_workload(dev, wl, size, 1, depth, "RGB8")
    so = bench.build_workload(oracle_mt, wl, size, 1, depth, "RGB8")
    cams = bench.make_cameras(dev, sg, faces)
    total = bad = hitsum = 0
    for f, cam in enumerate(cams):
        rays, _ = dev.primary_rays(sg.renderer, cam, sg.framebuffer, size, size, 1)
        hg, _ = dev.trace_rays(sg.scene, rays, closest=True)
        ho, _ = oracle_mt.trace_rays(so.scene, rays, closest=True)
        diff = (ids(hg) != ids(ho)).any(axis=1)
        hit = ids(ho)[:, 0] >= 0
        same = hit & ~diff
        bits = (hg[same][:, [0, 1, 2, 5, 6, 7]].view(np.uint32) != ho[same][:, [0, 1, 2, 5, 6, 7]].view(np.uint32)).any(axis=1)
        total += len(rays); bad += int(diff.sum()) + int(bits.sum()); hitsum += int(hit.sum())
    print(f"\n{wl}: {total} primary rays on {len(cams)} faces of {size}x{size}, {hitsum / total:.1%} hit, mismatch fraction {bad / total:.2e}")
    assert hitsum > 0.2 * total
    dev.close()
    assert bad == 0, f"{bad} of {total} primary hit records differ from the oracle"


def _brute_force_f64(tri, rays, cull):
    """Nearest hit of every ray against every triangle in float64 (Moeller-Trumbore, torch on the GPU; no BVH, nothing shared with the
    product). tri: (n,3,3), rays: (m,8) -> (t, triangle index, second-nearest t, |cos| between ray and the hit triangle's normal)."""
    import torch
    T = torch.as_tensor(tri, dtype=torch.float64, device="cuda")
    R = torch.as_tensor(rays, dtype=torch.float64, device="cuda")
    p0, e1, e2 = T[:, 0], T[:, 1] - T[:, 0], T[:, 2] - T[:, 0]
    Nrh = torch.cross(e1, e2, dim=-1); N = Nrh / Nrh.norm(dim=-1, keepdim=True).clamp_min(1e-300)
    C = torch.as_tensor(cull, dtype=torch.bool, device="cuda")
    chunk = max(16, int(1.5e7 // len(T)))
    out = [[], [], [], []]
    for b in range(0, len(R), chunk):
        O, D, tn, tf = R[b:b + chunk, 0:3], R[b:b + chunk, 4:7], R[b:b + chunk, 3], R[b:b + chunk, 7]
        k = len(O)
        P = torch.cross(D[:, None, :].expand(k, len(T), 3), e2[None].expand(k, len(T), 3), dim=-1)
        det = (P * e1[None]).sum(-1)
        inv = 1.0 / det
        S = O[:, None, :] - p0[None]
        u = (S * P).sum(-1) * inv
        Q = torch.cross(S, e1[None].expand(k, len(T), 3), dim=-1)
        v = (Q * D[:, None, :]).sum(-1) * inv
        t = (Q * e2[None]).sum(-1) * inv
        ok = (det != 0) & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > tn[:, None]) & (t <= tf[:, None])
        # back-face cull filter (trianglemesh_full.cpp:101-127): a culled mesh keeps a hit only if dot((p1-p0)x(p2-p0), dir) < 0
        ok &= ~(C[None, :] & ((D[:, None, :] * Nrh[None]).sum(-1) >= 0))
        t = torch.where(ok, t, torch.full_like(t, float("inf")))
        best = torch.topk(t, 2, dim=1, largest=False)
        idx = best.indices[:, 0]
        out[0].append(best.values[:, 0]); out[1].append(idx); out[2].append(best.values[:, 1]); out[3].append((N[idx] * D).sum(-1).abs())
    return [torch.cat(o).cpu().numpy() for o in out]


@pytest.mark.parametrize("scene_kind", ["spheres", "atrium"])
def test_hits_against_float64_brute_force(scene_kind):
    from yulio_raytracer_b200 import Device, workloads as W
    dev = Device.cuda()
    captured = []
    orig = W.add_mesh

    def recording_add_mesh(d, positions, indices, *a, **k):                    # the triangles exactly as handed to the device
        captured.append((np.asarray(positions, np.float32).astype(np.float64).reshape(-1, 3), np.asarray(indices).reshape(-1, 3), bool(k.get("cull", False))))
        return orig(d, positions, indices, *a, **k)
    W.add_mesh = recording_add_mesh
    try:
        if scene_kind == "spheres":
            # the sphere is tessellated inside the device (shapes/sphere.h:51-81): only the floor arrives as a mesh (geomID 1), so the truth set
            # is the floor and the rays point steeply down from far outside the sphere
            s = W.spheres(dev, "glass", 64, 64, 1, 2, face=None)
            geom_of_mesh = [1]
        else:
            s = W.atrium(dev, 64, 64, 1, 2, face=None, detail=10, tex_size=16)
            geom_of_mesh = list(range(len(captured)))
            captured.pop(); geom_of_mesh.pop()                                   # last mesh: the camera-facing billboard (placed by its transform)
    finally:
        W.add_mesh = orig
    tris = np.concatenate([p[t] for p, t, _ in captured])
    owner = np.concatenate([np.full(len(t), g) for g, (_, t, _) in zip(geom_of_mesh, captured)])
    prim = np.concatenate([np.arange(len(t)) for _, t, _ in captured])
    cull = np.concatenate([np.full(len(t), c) for _, t, c in captured])
    rng = np.random.default_rng(2024)
    m = 120_000
    rays = np.zeros((m, 8), np.float32)
    d = rng.normal(0, 1, (m, 3))
    if scene_kind == "spheres":
        rays[:, 0:3] = rng.uniform([-900, 300, -900], [900, 900, 900], (m, 3))
        d[:, 1] = -np.abs(d[:, 1]) - 2.0
    else:
        lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
        rays[:, 0:3] = rng.uniform(lo + 5, hi - 5, (m, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 4:7] = d; rays[:, 7] = np.inf
    hg, _ = dev.trace_rays(s.scene, rays, closest=True)
    t64, i64, t2, cos = _brute_force_f64(tris, rays, cull)
    g_hit = ids(hg)[:, 0] >= 0
    b_hit = np.isfinite(t64)
    scale = float(np.abs(tris).max())
    ulp = float(np.spacing(np.float32(scale)))
    tg = hg[:, 0].astype(np.float64)
    # geometry outside the truth set (the sphere, the billboard) can only make the GPU hit nearer: those rays are not comparable
    other_first = g_hit & ~np.isin(ids(hg)[:, 0], geom_of_mesh)
    comparable = b_hit & ~other_first
    assert comparable.sum() > 0.4 * m
    np.seterr(invalid="ignore")
    missed = comparable & ~g_hit
    tie = comparable & g_hit & ((t2 - t64) <= 1e-5 * np.maximum(1.0, t64))
    # the error of the hit point measured along the triangle normal: |t - t64| * |cos| <= 8 ulp(scene scale)
    bad_t = comparable & g_hit & (np.abs(tg - t64) * cos > 8 * ulp)
    same_id = (ids(hg)[:, 0] == owner[i64]) & (ids(hg)[:, 1] == prim[i64])
    bad_id = comparable & g_hit & ~same_id & ~tie
    n = int(comparable.sum())
    print(f"\n{scene_kind}: {m} rays vs {len(tris)} triangles in float64: {n} comparable hits; edge/tie fraction {tie.sum() / n:.2e}; "
          f"IDs differing outside ties {int(bad_id.sum())}; missed {int(missed.sum())}; hit point beyond 8 ulp {int(bad_t.sum())}")
    assert missed.sum() == 0, "the GPU missed triangles the float64 brute force hits"
    assert bad_id.sum() == 0, "primitive IDs differ from the float64 nearest triangle outside edge/tie cases"
    assert bad_t.sum() == 0, "hit distance differs from the float64 truth"
    assert tie.sum() / n < 5e-3
    dev.close()
