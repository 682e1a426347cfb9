#!/usr/bin/env python3
"""Where the time of a group-device session goes (development tool).  python tools/group_probe.py <gpus> [size] [spp]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yulio_raytracer_b200 import Device
from tests import scenes
n = int(sys.argv[1]); size = int(sys.argv[2]) if len(sys.argv) > 2 else 1536; spp = int(sys.argv[3]) if len(sys.argv) > 3 else 256
t = time.perf_counter(); dev = Device.cuda(cfg=f"gpus={n}" if n > 1 else ""); print(f"create            {time.perf_counter() - t:7.3f} s")
t = time.perf_counter(); s = scenes.atrium(dev, size, size, spp, 10, face=0, detail=8, fmt="RGB8"); print(f"scene + commit    {time.perf_counter() - t:7.3f} s")
for i, _ in zip(range(4), scenes.render_cube_map(dev, s)):
    pass
t = time.perf_counter()
for i, (f, cam) in enumerate(scenes.render_cube_map(dev, s, faces=[0, 1, 2, 3])):
    t1 = time.perf_counter(); dev.rtMapFrameBuffer(s.framebuffer); dev.rtUnmapFrameBuffer(s.framebuffer); t2 = time.perf_counter()
    st = dev.frame_stats()
    print(f"face {f}: since start {t2 - t:7.3f} s, map {t2 - t1:6.3f} s, device render {st.render_ms:8.1f} ms, rays {st.rays_closest + st.rays_shadow:.3e}")
t = time.perf_counter(); dev.close(); print(f"close             {time.perf_counter() - t:7.3f} s")
