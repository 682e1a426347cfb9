#!/usr/bin/env python3
"""One process, N GPUs behind one Device (cfg gpus=N, csrc/group_api.cu): the reference-facing per-face loop of bench.py's e2e leg
(update primitives, commit, render, swap, map to the host) on the group device.   python tools/group_bench.py <workload> <N> [faces]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from yulio_raytracer_b200 import Device
wl, n = sys.argv[1], int(sys.argv[2]); faces = int(sys.argv[3]) if len(sys.argv) > 3 else 12
_, desc, size, spp, depth = bench.WORKLOADS[wl]
dev = Device.cuda(cfg=f"gpus={n}" if n > 1 else "")
s = bench.build_workload(dev, wl, size, spp, depth, "RGB8")
for i in range(3):
    bench.render_face(dev, s, bench.face_camera(dev, s, i)); dev.rtSwapBuffers(s.framebuffer); dev.rtMapFrameBuffer(s.framebuffer); dev.rtUnmapFrameBuffer(s.framebuffer)
rays = 0; dev_ms = 0.0; t0 = time.perf_counter()
for i in range(faces):
    cam = bench.face_camera(dev, s, i)
    bench.render_face(dev, s, cam)
    dev.rtSwapBuffers(s.framebuffer); dev.rtMapFrameBuffer(s.framebuffer); dev.rtUnmapFrameBuffer(s.framebuffer)
    st = dev.frame_stats(); rays += st.rays_closest + st.rays_shadow; dev_ms += st.render_ms
dt = time.perf_counter() - t0
print(json.dumps({"workload": wl, "gpus_in_one_process": n, "faces": faces, "e2e_Mrays_s": rays / dt / 1e6, "device_Mrays_s": rays / (dev_ms * 1e-3) / 1e6,
                  "e2e_s_per_stereo_cube_map": dt / faces * 12, "ms_per_face_device_max_over_gpus": dev_ms / faces}))
