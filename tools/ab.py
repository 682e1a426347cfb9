#!/usr/bin/env python3
"""Development A/B on the GPU box: stage times of one workload for several library variants / cfg strings.
    python tools/ab.py <workload> <size> <cube maps> <lib-or-'default'>[:cfg] ...
Every variant renders the same stereo cube map(s) through yrtxRenderCubeMap; the first one is a warm-up."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from yulio_raytracer_b200 import devapi
wl, size, maps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
_, desc, _, spp, depth, faces = bench.WORKLOADS[wl]
for spec in sys.argv[4:]:
    lib, _, cfg = spec.partition(":")
    path = devapi.CUDA_LIB if lib == "default" else os.path.join(os.path.dirname(devapi.CUDA_LIB), "variants", f"libyrt_{lib}.so")
    dev = devapi.Device(path, cfg=cfg)
    dev.set_readback(False)
    s = bench.build_workload(dev, wl, size, spp, depth, "RGB8")
    cams = bench.make_cameras(dev, s, faces)
    fbs = [s.framebuffer] + [dev.rtNewFrameBuffer("RGB8", size, size, 1) for _ in range(len(cams) - 1)]
    acc = {}
    for i in range(maps + 1):
        bench.render_step(dev, s, cams, fbs)
        st = dev.frame_stats()
        if i == 0: continue
        for k in ("render_ms", "closest_ms", "shadow_ms", "shade_ms", "resolve_ms", "miss_ms", "raygen_film_ms"): acc[k] = acc.get(k, 0.0) + getattr(st, k) / maps
        acc["build_ms"] = st.build_ms
        acc["rays"] = acc.get("rays", 0) + (st.rays_closest + st.rays_shadow) / maps
        if st.node_visits: acc["nodes/ray"] = st.node_visits / (st.rays_closest + st.rays_shadow); acc["tris/ray"] = st.tri_tests / (st.rays_closest + st.rays_shadow)
    print(f"{wl} {size} {spec:28s} " + " ".join(f"{k}={v:9.3f}" for k, v in acc.items() if k != "rays") + f" Mrays/s={acc['rays'] / acc['render_ms'] / 1e3:8.1f}", flush=True)
    dev.close()
