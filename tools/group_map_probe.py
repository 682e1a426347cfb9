#!/usr/bin/env python3
"""Where the time of a group device's yrtMapFrameBuffer goes (2+ GPUs): peer assembly on member 0 + one D2H per frame.
    python tools/group_map_probe.py [gpus] [face size]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yulio_raytracer_b200 import Device, workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
size = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
d = Device.cuda(cfg=f"gpus={n}")
s = W.atrium(d, size, size, 1, 2, face=0, detail=8, fmt="RGB8", tex_size=32)
cams = W.cube_cameras(d, s)
fbs = [s.framebuffer] + [d.rtNewFrameBuffer("RGB8", size, size, 1) for _ in range(11)]
for rep in range(3):
    t0 = time.perf_counter(); W.render_cube_map_batched(d, s, cams, fbs); t1 = time.perf_counter()
    per = []
    for fb in fbs:
        a = time.perf_counter(); d.rtSwapBuffers(fb); d.rtMapFrameBuffer(fb); d.rtUnmapFrameBuffer(fb); per.append((time.perf_counter() - a) * 1e3)
    t2 = time.perf_counter()
    d.strip_begin(size, size)
    a = time.perf_counter()
    for i, fb in enumerate(fbs):
        d.strip_add_face(fb, i)
    t3 = time.perf_counter()
    print(f"rep {rep}: render {1e3 * (t1 - t0):.1f} ms, 12 maps {1e3 * (t2 - t1):.1f} ms ({min(per):.2f}..{max(per):.2f} ms each, {size * size * 3 / 1e6:.1f} MB per frame), "
          f"12 strip_add_face {1e3 * (t3 - a):.1f} ms, d2h bytes {d.frame_stats().d2h_bytes}")
d.close()
