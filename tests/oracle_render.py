"""Renders one of the test scenes on the reference CPU device with all host cores, in its own process (the reference keeps
process-wide state — task scheduler, g_device — so a second device with another thread count cannot share the pytest process),
and saves the linear float frame as .npy.   python -m tests.oracle_render <cornell|glass|atrium> <size> <spp> <depth> <out.npy>"""
import sys

import numpy as np


def build(d, name, size, spp, depth):
    from tests import scenes
    if name == "cornell":
        return scenes.cornell(d, size, size, spp, depth)
    if name == "glass":
        return scenes.spheres(d, "glass", size, size, spp, depth, face=3)
    return scenes.atrium(d, size, size, spp, depth, face=1, detail=6, tex_size=64)


def main():
    from oracle import oracle_device
    name, size, spp, depth, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    d = oracle_device.open_oracle(num_threads=0)
    s = build(d, name, size, spp, depth)
    d.rtRenderFrame(s.renderer, s.camera, s.scene, s.tonemapper, s.framebuffer, 0)
    np.save(out, d.read_framebuffer(s.framebuffer, "RGB_FLOAT32", size, size))


if __name__ == "__main__":
    main()
