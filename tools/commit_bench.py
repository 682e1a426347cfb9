#!/usr/bin/env python3
"""Times rtCommit(scene) when a billboard moved (the reference rebuilds per cube face). Development tool."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import scenes
from yulio_raytracer_b200 import Device
detail = int(sys.argv[1]) if len(sys.argv) > 1 else 56
dev = Device.cuda(cfg="verbose=1")
s = scenes.atrium(dev, 64, 64, 1, 2, face=0, detail=detail)
pos, target, up = s.view
for face in range(4):
    cam = scenes.stereo_camera(dev, face, pos, target, up)
    org = dev.rtGetFloat3(cam, "origin")
    t0 = time.perf_counter()
    for j, p in enumerate(s.prims):
        dev.rtUpdatePrimitive(s.scene, j, p, org, up)
    t1 = time.perf_counter()
    dev.rtCommit(s.scene)
    t2 = time.perf_counter()
    st = dev.frame_stats()
    print(f"face {face}: update {1e3*(t1-t0):.2f} ms, commit {1e3*(t2-t1):.2f} ms, build_ms {st.build_ms:.3f}, tris {st.num_triangles}", flush=True)
