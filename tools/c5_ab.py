#!/usr/bin/env python3
"""A/B of the raw traversal on the config-5 soup: the triangle arrays are generated once per size, every variant builds its own scene.
    python tools/c5_ab.py <tris> <log2 rays> <lib-or-'default'>[:cfg] ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from yulio_raytracer_b200 import devapi, workloads as W
tris, lg = int(float(sys.argv[1])), int(sys.argv[2])
n = 1 << lg
t0 = time.perf_counter()
pos, tri = W.soup_arrays(tris)
rays = torch.from_numpy(W.random_rays(n)).cuda()
hits = torch.zeros((n, 8), dtype=torch.float32, device="cuda")
print(f"generated {tris} triangles, {n} rays in {time.perf_counter() - t0:.1f} s", flush=True)
for spec in sys.argv[3:]:
    lib, _, cfg = spec.partition(":")
    path = devapi.CUDA_LIB if lib == "default" else os.path.join(os.path.dirname(devapi.CUDA_LIB), "variants", f"libyrt_{lib}.so")
    d = devapi.Device(path, cfg=cfg)
    matte = d.rtNewMaterial("matte"); d.rtCommit(matte)
    meshes = max(1, tris // 4_000_000); per = (tris + meshes - 1) // meshes
    prims = []
    for k in range(meshes):
        t = tri[k * per:(k + 1) * per]
        if len(t):
            prims.append(d.rtNewShapePrimitive(W.add_mesh(d, pos[t[0, 0]:t[-1, 2] + 1], t - t[0, 0]), matte, None))
    sc = W.make_scene(d, prims)
    ms = [d.trace_rays_device(sc, rays.data_ptr(), hits.data_ptr(), n, True) for _ in range(4)]
    st = d.frame_stats()
    extra = f" nodes/ray {st.node_visits / n:.2f} tris/ray {st.tri_tests / n:.2f}" if st.node_visits else ""
    print(f"c5 {tris:>10d} {spec:28s} best {n / min(ms[1:]) / 1e3:8.1f} Mrays/s  (ms {', '.join(f'{m:.2f}' for m in ms)}) build {st.build_ms:.0f} ms{extra}", flush=True)
    d.close()
