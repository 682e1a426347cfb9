#!/usr/bin/env python3
"""Per-bounce stage times of one face (cfg verbose=2 prints the CUDA-event spans in launch order).  python tools/bounce_profile.py c3 [size]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from yulio_raytracer_b200 import Device
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
_, desc, size, spp, depth = bench.WORKLOADS[wl]
if len(sys.argv) > 2: size = int(sys.argv[2])
dev = Device.cuda(cfg="verbose=2")
s = bench.build_workload(dev, wl, size, spp, depth, "RGB8")
for i in range(3):
    bench.render_face(dev, s, bench.face_camera(dev, s, i))
