#!/usr/bin/env python3
"""Where the fixed start-up time of a session goes (development tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.perf_counter()
from yulio_raytracer_b200 import Device
from tests import scenes
t1 = time.perf_counter(); print(f"import            {t1 - t0:7.3f} s")
dev = Device.cuda(cfg=os.environ.get("YRT_CFG", ""))
t2 = time.perf_counter(); print(f"yrtCreateDevice   {t2 - t1:7.3f} s")
s = scenes.atrium(dev, 1024, 1024, 64, 10, face=0, detail=8, fmt="RGB8")
t3 = time.perf_counter(); print(f"scene + commit    {t3 - t2:7.3f} s")
for i in range(3):
    t = time.perf_counter()
    dev.rtRenderFrame(s.renderer, s.camera, s.scene, s.tonemapper, s.framebuffer, 0)
    print(f"rtRenderFrame #{i}   {time.perf_counter() - t:7.3f} s   (device {dev.frame_stats().render_ms:.1f} ms)")
t = time.perf_counter(); dev.strip_begin(1024, 1024); dev.strip_add_face(s.framebuffer, 0); dev.strip_encode_jpeg("/tmp/probe.jpg", 90)
print(f"strip + first JPEG  {time.perf_counter() - t:7.3f} s")
t = time.perf_counter(); dev.strip_encode_jpeg("/tmp/probe.jpg", 90); print(f"second JPEG       {time.perf_counter() - t:7.3f} s")
