# round 2, GPU call 9 (1 GPU): batched triangle phase (shared-memory work list) vs one triangle per lane, at 8 / 7 / 6 CTAs per SM
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -3
for V in tb8 tb7 tb6; do YRT_TEST_LIB=$V python - <<PY
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from yulio_raytracer_b200 import devapi, workloads as W
v = os.environ["YRT_TEST_LIB"]
ref = None
for path in (devapi.CUDA_LIB, os.path.join(os.path.dirname(devapi.CUDA_LIB), "variants", f"libyrt_{v}.so")):
    d = devapi.Device(path)
    s = W.atrium(d, 96, 80, 16, 8, face=4, detail=6, tex_size=32)
    d.rtRenderFrame(s.renderer, s.camera, s.scene, s.tonemapper, s.framebuffer, 0)
    img = d.read_framebuffer(s.framebuffer, "RGB_FLOAT32", 96, 80); st = d.frame_stats()
    if ref is None: ref = (img, st.rays_closest, st.rays_shadow)
    else: print(v, "bit-identical frame:", np.array_equal(img.view(np.uint32), ref[0].view(np.uint32)), "rays equal:", (st.rays_closest, st.rays_shadow) == ref[1:])
    d.close()
PY
done
python tools/ab.py c4 2048 1 default tb8 tb8:trinum=2 tb8:trinum=1 tb7:tracectas=7 tb7:tracectas=7,trinum=2 tb6:tracectas=6 tb6:tracectas=6,trinum=2 tb8:stats=1 default:stats=1 2>&1 | tee gpurun_out/r2i_ab_c4.txt
python tools/ab.py c3 1024 1 default tb8 tb7:tracectas=7 2>&1 | tee gpurun_out/r2i_ab_c3.txt
python tools/ab.py c2 1024 1 default tb8 tb7:tracectas=7 2>&1 | tee gpurun_out/r2i_ab_c2.txt
