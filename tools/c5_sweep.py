#!/usr/bin/env python3
"""BASELINE config 5: synthetic triangle soup, incoherent random-ray traversal sweep.

    python tools/c5_sweep.py --sizes 1e5,1e6,1e7 [--rays 16777216] [--json gpurun_out/c5.json]

For every size: build the soup (blobs of 1000 small triangles, centres uniform in [0,1000]^3) through the Device API, commit
(GPU BVH build), trace 2^24 random segments closest-hit (tfar = inf) and any-hit (tfar ~ U(0,200)) with device-resident rays
(yrtxTraceRays onDevice), and a coherent pinhole sweep. Reports Mrays/s, N̄node/N̄tri from a stats replay, the algorithmic
bytes/ray = 64 (36 any-hit) + 80 N̄node + 48 N̄tri and the fraction of the measured HBM bandwidth.
"""
import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tests import scenes
from yulio_raytracer_b200 import Device


def pinhole_rays(n, extent):
    side = int(math.sqrt(n))
    y, x = np.mgrid[0:side, 0:side].astype(np.float32)
    d = np.stack([(x / side - 0.5), (y / side - 0.5), np.full_like(x, 0.8)], -1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.zeros((side * side, 8), np.float32)
    r[:, 0:3] = (extent * 0.5, extent * 0.5, -0.2 * extent); r[:, 4:7] = d; r[:, 7] = np.inf
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1e5,1e6,1e7")
    ap.add_argument("--rays", type=int, default=1 << 24)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--json", default="")
    ap.add_argument("--cfg", default="")
    a = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    extent = 1000.0
    sets = {
        "closest_incoherent": (scenes.random_rays(a.rays, seed=67890, extent=extent), True),
        "anyhit_incoherent": (scenes.random_rays(a.rays, seed=67891, extent=extent, tfar_uniform=200.0), False),
        "closest_coherent": (pinhole_rays(a.rays, extent), True),
    }
    out = []
    for size in [int(float(s)) for s in a.sizes.split(",")]:
        t0 = time.perf_counter()
        dev = Device.cuda(cfg=a.cfg)
        sc = scenes.soup(dev, size, seed=12345, extent=extent, meshes=max(1, size // 1_000_000))
        st = dev.frame_stats()
        build_wall = time.perf_counter() - t0
        sdev = Device.cuda(cfg="stats=1")
        ssc = scenes.soup(sdev, size, seed=12345, extent=extent, meshes=max(1, size // 1_000_000))
        row = {"triangles": size, "nodes": int(st.num_nodes), "bvh_build_ms": st.build_ms, "scene_setup_wall_s": build_wall,
               "bvh_bytes": int(st.num_nodes) * 80 + int(st.num_triangles) * 48}
        for name, (rays, closest) in sets.items():
            d_r = torch.from_numpy(rays).cuda(); d_h = torch.zeros_like(d_r)
            torch.cuda.synchronize()
            sub = min(len(rays), 1 << 20)
            sdev.trace_rays_device(ssc.scene, d_r.data_ptr(), d_h.data_ptr(), sub, closest)
            ss = sdev.frame_stats()
            nn, nt = ss.node_visits / sub, ss.tri_tests / sub
            best = 1e30
            for _ in range(a.reps + 1):
                best = min(best, dev.trace_rays_device(sc.scene, d_r.data_ptr(), d_h.data_ptr(), len(rays), closest))
            bpr = (64 if closest else 36) + 80 * nn + 48 * nt
            mr = len(rays) / best / 1e3
            hitfrac = float((d_h.view(torch.int32)[:, 3] >= 0).float().mean()) if closest else float((d_h.view(torch.int32)[:, 3] == 0).float().mean())
            row[name] = {"mrays_s": mr, "ms": best, "nodes_per_ray": nn, "tris_per_ray": nt, "bytes_per_ray": bpr,
                         "algorithmic_gb_s": mr * 1e6 * bpr / 1e9, "frac_of_hbm_peak": mr * 1e6 * bpr / 1e9 / peak, "hit_fraction": hitfrac}
            del d_r, d_h
        out.append(row)
        print(json.dumps(row), flush=True)
        dev.close(); sdev.close()
        torch.cuda.empty_cache()
    if a.json:
        json.dump({"hbm_peak_gb_s": peak, "rays": a.rays, "rows": out}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
