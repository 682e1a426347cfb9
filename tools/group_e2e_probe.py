#!/usr/bin/env python3
"""Where the wall-clock of one stereo cube map goes on the in-process group device (cfg gpus=N) compared with its CUDA-event time.
    python tools/group_e2e_probe.py <gpus> [workload]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from yulio_raytracer_b200 import Device
n = int(sys.argv[1]); wl = sys.argv[2] if len(sys.argv) > 2 else "c4"
_, _, size, spp, depth, faces = bench.WORKLOADS[wl]
d = Device.cuda(cfg=(f"gpus={n}" if n > 1 else "") + ("," + os.environ["YRT_CFG"] if os.environ.get("YRT_CFG") else ""))
s = bench.build_workload(d, wl, size, spp, depth, "RGB8")
fbs = [s.framebuffer] + [d.rtNewFrameBuffer("RGB8", size, size, 1) for _ in range(faces - 1)]
for rep in range(4):
    T = {}
    def lap(name, t0): T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); cams = bench.make_cameras(d, s, faces); lap("cameras", t0)
    t0 = time.perf_counter(); org = d.rtGetFloat3(cams[0], "origin")
    for j, p in enumerate(s.prims): d.rtUpdatePrimitive(s.scene, j, p, org, s.view[2])
    lap("update prims", t0)
    t0 = time.perf_counter(); d.rtCommit(s.scene); lap("commit", t0)
    t0 = time.perf_counter(); d.render_cube_map(s.renderer, cams, s.scene, s.tonemapper, fbs, 0); lap("render call (wall)", t0)
    st = d.frame_stats(); T["render (CUDA events, max over members)"] = st.render_ms; T["host_ms reported (max)"] = st.host_ms
    t0 = time.perf_counter()
    for fb in fbs: d.rtSwapBuffers(fb); d.rtMapFrameBuffer(fb); d.rtUnmapFrameBuffer(fb)
    lap("12 x swap + map", t0)
    t0 = time.perf_counter()
    for c in cams: d.rtDecRef(c)
    lap("release cameras", t0)
    print(f"rep {rep}: " + "; ".join(f"{k} {v:.1f}" for k, v in T.items()), flush=True)
d.close()
