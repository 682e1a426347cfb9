#!/usr/bin/env python3
"""Where the host time of the reference-facing per-face loop goes (development tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from yulio_raytracer_b200 import Device
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
_, desc, size, spp, depth = bench.WORKLOADS[wl]
dev = Device.cuda(cfg=os.environ.get("YRT_CFG", ""))
s = bench.build_workload(dev, wl, size, spp, depth, "RGB8")
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
for i in range(14):
    if i == 2: T.clear()
    t0 = time.perf_counter(); cam = bench.face_camera(dev, s, i); tick("camera", t0)
    t0 = time.perf_counter()
    org = dev.rtGetFloat3(cam, "origin")
    for j, p in enumerate(s.prims): dev.rtUpdatePrimitive(s.scene, j, p, org, s.view[2])
    tick("update", t0)
    t0 = time.perf_counter(); dev.rtCommit(s.scene); tick("commit", t0)
    t0 = time.perf_counter(); dev.rtRenderFrame(s.renderer, cam, s.scene, s.tonemapper, s.framebuffer, 0); tick("render", t0)
    st = dev.frame_stats(); T["render_ms_device"] = T.get("render_ms_device", 0) + st.render_ms * 1e-3; T["host_ms"] = T.get("host_ms", 0) + st.host_ms * 1e-3
    t0 = time.perf_counter(); dev.rtSwapBuffers(s.framebuffer); p = dev.rtMapFrameBuffer(s.framebuffer); dev.rtUnmapFrameBuffer(s.framebuffer); tick("swap+map", t0)
for k, v in T.items(): print(f"{k:18s} {v / 12 * 1e3:8.3f} ms per face")
