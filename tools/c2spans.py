import sys; sys.path.insert(0,'.')
import bench
from yulio_raytracer_b200 import Device
for cfg in ("trav=0","trav=1"):
    dev = Device.cuda(cfg=cfg+",verbose=2")
    s = bench.build_workload(dev, "c2", 1024, 64, 8, "RGB8")
    for f in (0,1):
        cam = bench.face_camera(dev, s, f)
        bench.render_face(dev, s, cam)
        bench.render_face(dev, s, cam)
        print(cfg, "face", f, flush=True)
    dev.close()
