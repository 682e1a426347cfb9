#!/usr/bin/env python3
"""Development A/B on the GPU box: stage times of one workload for several library variants / cfg strings.
    python tools/ab.py <workload> <size> <faces> <lib-or-'default'>[:cfg] ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from yulio_raytracer_b200 import devapi
wl, size, faces = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
_, desc, _, spp, depth = bench.WORKLOADS[wl]
for spec in sys.argv[4:]:
    lib, _, cfg = spec.partition(":")
    path = devapi.CUDA_LIB if lib == "default" else os.path.join(os.path.dirname(devapi.CUDA_LIB), "variants", f"libyrt_{lib}.so")
    dev = devapi.Device(path, cfg=cfg)
    s = bench.build_workload(dev, wl, size, spp, depth, "RGB8")
    acc = {}
    for i in range(faces + 1):
        cam = bench.face_camera(dev, s, i)
        bench.render_face(dev, s, cam)
        st = dev.frame_stats()
        if i == 0: continue
        for k in ("render_ms", "closest_ms", "shadow_ms", "shade_ms", "raygen_film_ms"): acc[k] = acc.get(k, 0.0) + getattr(st, k) / faces
        acc["build_ms"] = st.build_ms
        acc["rays"] = acc.get("rays", 0) + (st.rays_closest + st.rays_shadow) / faces
        if st.node_visits: acc["nodes/ray"] = st.node_visits / (st.rays_closest + st.rays_shadow); acc["tris/ray"] = st.tri_tests / (st.rays_closest + st.rays_shadow)
    print(f"{spec:28s} " + " ".join(f"{k}={v:9.3f}" for k, v in acc.items() if k != "rays") + f" Mrays/s={acc['rays'] / acc['render_ms'] / 1e3:8.1f}", flush=True)
    dev.close()
