#!/usr/bin/env python3
"""Traversal micro-benchmark (development tool, not a bench line): times yrtxTraceRays on device-resident rays for a
scene of tests/scenes.py, for primary (coherent) and diffuse-bounce (incoherent) rays, across scheduling knobs.

    python tools/trace_bench.py --scene atrium --detail 56 --size 512 --cfgs "refill=8,trinum=3,triden=1;refill=16,trinum=3,triden=1"
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tests import scenes
from yulio_raytracer_b200 import Device


def build(dev, a):
    if a.scene == "atrium":
        return scenes.atrium(dev, a.size, a.size, a.spp, 4, face=a.face, detail=a.detail)
    if a.scene == "spheres":
        return scenes.spheres(dev, "glass", a.size, a.size, a.spp, 4, face=a.face)
    return scenes.cornell(dev, a.size, a.size, a.spp, 2)


def bounce_rays(rays, hits, seed=1):
    rng = np.random.default_rng(seed)
    hit = hits.view(np.int32)[:, 3] >= 0
    r, h = rays[hit], hits[hit]
    P = r[:, 0:3] + h[:, 0:1] * r[:, 4:7]
    n = h[:, 5:8] / np.maximum(1e-20, np.linalg.norm(h[:, 5:8], axis=1, keepdims=True))
    n = np.where((np.sum(n * r[:, 4:7], axis=1, keepdims=True) > 0), -n, n)
    d = rng.normal(0, 1, P.shape).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = np.where(np.sum(d * n, axis=1, keepdims=True) < 0, -d, d)
    out = np.zeros((len(P), 8), np.float32)
    out[:, 0:3] = P; out[:, 3] = 1e-3 * np.maximum(1.0, np.abs(P).max(axis=1)); out[:, 4:7] = d; out[:, 7] = np.inf
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="atrium"); ap.add_argument("--detail", type=int, default=56)
    ap.add_argument("--size", type=int, default=512); ap.add_argument("--spp", type=int, default=4); ap.add_argument("--face", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cfgs", default="refill=8,trinum=3,triden=1")
    a = ap.parse_args()
    base = Device.cuda()
    s = build(base, a)
    prim, _ = base.primary_rays(s.renderer, s.camera, s.framebuffer, a.size, a.size, a.spp)
    hits, _ = base.trace_rays(s.scene, prim, closest=True)
    sec = bounce_rays(prim, hits)
    shadow = sec.copy(); shadow[:, 7] = 120.0
    sets = {"primary": (prim, True), "bounce": (sec, True), "shadow120": (shadow, False)}
    base.close()
    for cfg in a.cfgs.split(";"):
        dev = Device.cuda(cfg=cfg + ",stats=0")
        sc = build(dev, a)
        line = [f"{cfg:32s}"]
        for name, (rays, closest) in sets.items():
            d_r = torch.from_numpy(rays).cuda(); d_h = torch.zeros_like(d_r)
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(a.reps):
                ms = dev.trace_rays_device(sc.scene, d_r.data_ptr(), d_h.data_ptr(), len(rays), closest)
                best = min(best, ms)
            line.append(f"{name} {len(rays) / best / 1e3:8.1f} Mrays/s ({best:.3f} ms)")
        print(" | ".join(line), flush=True)
        dev.close()


if __name__ == "__main__":
    main()
