#!/usr/bin/env python3
"""bench.py — the reference's headline benchmark on device_cuda: Mrays/s (path segments) and seconds per stereo cube map,
driven through the C-ABI (include/yrt_device.h) as the reference's outputMode loop drives a device
(devices/renderer/renderer.cpp:543-632).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c3|c2|c1|c5] [--impl reference]

A "step" is ONE STEREO CUBE MAP of the workload: the 12 stereo cube cameras of a viewpoint (ColladaLoader.cpp:470-505), rendered by
yrtxRenderCubeMap as one wavefront after the camera-aligned primitives were turned and the scene committed (renderer.cpp:551-559).
Default workload = BASELINE.json configs[3], the north-star target: the sample-scene stand-in "office" (the .dae is stripped from the
reference mount; its 152 real texture images are not: data/sample_scene, read through rtNewImageFromFile), 2048x2048 per face, 16 spp
(rt_test_dll.cpp:17), DLL defaults otherwise (YulioRT.h:37-50). Metric = the reference's own counter and timer span:
(rtcIntersect + rtcOccluded equivalents) / render time (integratorrenderer.cpp:99-111, pathtraceintegrator.cpp:74,161).

  value   device-timed (CUDA events on the device's stream around the wavefront loop, + the band gather for N > 1), frames left in HBM
  e2e     host wall-clock of the reference-facing loop per cube map: 12 x rtNewCamera, rtUpdatePrimitive x prims, rtCommit(scene),
          yrtxRenderCubeMap, rtSwapBuffers + rtMapFrameBuffer x 12 -> every frame is in a (pinned) host buffer
  N > 1   one process per GPU (torchrun, the driver's launch): the scene is replicated, every face is split into the reference's 4-row
          bands dealt round-robin to the ranks (api/swapchain.h:57-70, the reference's own network-device partition), each rank renders
          its bands of all 12 faces as one wavefront, ONE NCCL all-gather per cube map brings them to rank 0; strong scaling.
          (--gpus N without torchrun: one process, the in-process group device, cfg gpus=N.)
--impl reference / cpu_baseline: the reference's own CPU path (oracle/_ref: devices/device_singleray sources built -O3 -ffast-math + the
embree2-API shim) on all host cores; each step is a bounded sample of the SAME faces at the SAME size: every B-th 4-row band of a face
(the reference's own serverID/serverCount partition), B chosen so that a step is ~2 M paths.
--workload c5: BASELINE configs[4], raw traversal of a synthetic triangle soup (--tris), rays split over the ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

METRIC = "Mrays/s (path segments)"
WORKLOADS = {
    # name: (builder, description, face size, spp, depth, faces per step)
    "c4": ("office", "C4 sample-scene stand-in (22 Frederick St .dae stripped from the mount): procedural office, 1.20 M triangles, the scene's 152 REAL "
                     "texture images (273 MB as RGBA8, 30 with alpha) on Uber / ThinDielectric materials through rtNewImageFromFile, billboard, dome light, "
                     "DLL defaults (depth 10, tMaxShadowRay 120, toe-in stereo); stereo cube map = 12 faces of 2048x2048, 16 spp (rt_test_dll.cpp:17)", 2048, 16, 10, 12),
    "c3": ("atrium", "C3 Sponza stand-in (Sponza.DAE stripped): procedural atrium, 294 k triangles, the 14 REAL models/Sponza JPG textures + 2 RGBA cut-outs, "
                     "Uber + alpha + dome light, tMaxShadowRay 120; stereo cube map = 12 faces of 1024x1024, 64 spp, depth 10", 1024, 64, 10, 12),
    "c2": ("spheres", "C2 sphere_glass.xml + sphere_view.ecs: stereo cube map = 12 faces of 1024x1024, 64 spp, depth 8 (procedural lines texture stands in for lines.ppm)", 1024, 64, 8, 12),
    "c1": ("cornell", "C1 cornell_box.ecs: pinhole 512x512, 16 spp, depth 2 (one frame per step)", 512, 16, 2, 1),
}


def build_workload(dev, name, size, spp, depth, fmt):
    from yulio_raytracer_b200 import workloads as W
    kind = WORKLOADS[name][0]
    if kind == "office":
        return W.office(dev, size, size, spp, depth, face=0, detail=112, fmt=fmt)
    if kind == "atrium":
        return W.atrium(dev, size, size, spp, depth, face=0, detail=56, fmt=fmt, tex_set="sponza")
    if kind == "spheres":
        return W.spheres(dev, "glass", size, size, spp, depth, face=0, fmt=fmt)
    s = W.cornell(dev, size, size, spp, depth, fmt=fmt)
    s.view = None
    return s


def make_cameras(dev, s, faces):
    from yulio_raytracer_b200 import workloads as W
    if s.view is None:
        return [s.camera]
    return W.cube_cameras(dev, s, faces=range(faces))


def render_step(dev, s, cams, fbs, per_face=False):
    """One stereo cube map: the per-viewpoint body of outputMode (renderer.cpp:551-580)."""
    from yulio_raytracer_b200 import workloads as W
    if s.view is None:
        dev.rtRenderFrame(s.renderer, cams[0], s.scene, s.tonemapper, fbs[0], 0)
    elif per_face:                                      # the reference's literal loop: update, commit, render — per face
        for cam, fb in zip(cams, fbs):
            org = dev.rtGetFloat3(cam, "origin")
            for j, p in enumerate(s.prims):
                dev.rtUpdatePrimitive(s.scene, j, p, org, s.view[2])
            dev.rtCommit(s.scene)
            dev.rtRenderFrame(s.renderer, cam, s.scene, s.tonemapper, fb, 0)
    else:
        W.render_cube_map_batched(dev, s, cams, fbs)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.rows, self.stop = [], threading.Event()
        self.gpu = gpu
        self.thread = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread.start(); return self

    def __exit__(self, *a):
        self.stop.set(); self.thread.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, n in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peak():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference(workload, steps, warmup):
    """The reference CPU path on the host cores. Same scene / cameras / face size / spp / depth as the GPU arm; one step = the rows of ONE
    face that server 0 of B renders under the reference's own band partition (every B-th 4-row band, api/swapchain.h:57-70)."""
    from oracle import oracle_device
    _, desc, size, spp, depth, faces = WORKLOADS[workload]
    cores = host_cores()
    fast = os.path.exists(oracle_device.ORACLE_FAST_LIB)
    dev = oracle_device.open_oracle(num_threads=cores, fast=fast)
    bands = max(1, (size * size * spp) >> 21)
    bands = min(bands, max(1, size // 4))
    dev.rtSetInt1(None, "serverID", 0); dev.rtSetInt1(None, "serverCount", bands)      # singleray_device.cpp:505-508
    s = build_workload(dev, workload, size, spp, depth, "RGB8")
    cams = make_cameras(dev, s, faces)
    rays = secs = wall = 0.0
    for i in range(warmup + steps):
        cam = cams[i % len(cams)]
        t0 = time.perf_counter()
        if s.view is not None:
            org = dev.rtGetFloat3(cam, "origin")
            for j, p in enumerate(s.prims):
                dev.rtUpdatePrimitive(s.scene, j, p, org, s.view[2])
            dev.rtCommit(s.scene)
        dev.rtRenderFrame(s.renderer, cam, s.scene, s.tonemapper, s.framebuffer, 0)
        dev.rtSwapBuffers(s.framebuffer); dev.rtMapFrameBuffer(s.framebuffer); dev.rtUnmapFrameBuffer(s.framebuffer)
        dt = time.perf_counter() - t0
        st = dev.frame_stats()
        if i >= warmup:
            rays += st.rays_closest; secs += st.render_ms * 1e-3; wall += dt
    sample = (f"reference devices/device_singleray sources, {'-O3 -ffast-math' if fast else '-O2 strict'} build + embree2-API shim (scalar BVH2; NOT Intel Embree), "
              f"{cores} threads; {steps} steps, each = one stereo cube face of the same scene at the SAME {size}x{size} px, {spp} spp, depth {depth}, "
              f"restricted to every {bands}-th 4-row band (serverCount={bands}: {size * size * spp // bands} paths per step)")
    return {"value": rays / secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference", "sample": sample,
            "e2e_value": rays / wall / 1e6, "ms_per_step": secs * 1e3 / steps, "rays": rays, "wall_s": wall,
            "s_per_stereo_cube_map_extrapolated": (wall / steps) * bands * faces}


def kernel_rooflines(agg, nbar, peaks, n_gpus_in_process):
    """Per-kernel algorithmic bytes / flops against the measured denominators. The dominant kernel's block becomes `roofline`."""
    out = {}
    hbm, l2, fp32 = peaks["hbm_gbs"], peaks.get("l2_gbs_measured"), peaks.get("fp32_tflops_measured")

    def trav(name, rays, ms, launches, nodes, tris, hit_bytes):
        if not rays or not ms:
            return
        rays /= n_gpus_in_process
        bpr = 32 + hit_bytes + 80 * nodes + 48 * tris
        fpr = 190 * nodes + 90 * tris
        gbs = rays * bpr / (ms * 1e-3) / 1e9
        tf = rays * fpr / (ms * 1e-3) / 1e12
        out[name] = {"bound": "hbm", "kernel": name, "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None,
                     "bytes_per_ray": bpr, "nodes_per_ray": nodes, "tris_per_ray": tris, "avg_launch_ms": ms / max(1, launches),
                     "l2_gbs_measured": l2, "frac_l2": (gbs / l2) if l2 else None,
                     "fp32_tflops_measured": fp32, "fp32_tflops_achieved": tf, "frac_fp32": (tf / fp32) if fp32 else None, "flops_per_ray": fpr,
                     "what_bounds_it": "instruction issue (ncu: profiles/): the BVH is L2-resident, so the memory-side bound is L2 read bandwidth (frac_l2), "
                                       "not HBM; `frac` is kept as the contract's algorithmic-bytes / HBM-peak figure"}
    trav("k_trace_closest", agg["closest_rays"], agg["closest_ms"], agg["closest_launches"], nbar["closest_nodes_per_ray"], nbar["closest_tris_per_ray"], 16)
    trav("k_trace_shadow", agg["shadow_rays"], agg["shadow_ms"], agg["shadow_launches"], nbar["shadow_nodes_per_ray"], nbar["shadow_tris_per_ray"], 4)
    if agg["shade_ms"] and agg["vertices"]:
        v = agg["vertices"] / n_gpus_in_process
        spv = agg["shadow_rays"] / max(1, agg["vertices"])
        # per path vertex: queue 4 + rayO/rayD/hit/throughput 64 in, 80 shading record, 16 texels, next ray + throughput + queue 52 out, 52 per shadow ray
        bpv = 4 + 64 + 80 + 16 + 52 + 52 * spv
        gbs = v * bpv / (agg["shade_ms"] * 1e-3) / 1e9
        out["k_shade"] = {"bound": "hbm", "kernel": "k_shade", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None,
                          "bytes_per_vertex": bpv, "shadow_rays_per_vertex": spv, "avg_launch_ms": agg["shade_ms"] / max(1, agg["shade_launches"]),
                          "what_bounds_it": "HBM stream of path state in the limit; measured: latency + divergence (ncu: profiles/)"}
    return out


def run_soup(args, rank, world, local_rank):
    """BASELINE configs[4]: incoherent closest-hit rays through a synthetic triangle soup, scene replicated, rays split over the ranks."""
    import torch
    import torch.distributed as dist
    from yulio_raytracer_b200 import Device, workloads as W
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = Device.cuda(cfg=f"gpu={local_rank}" + ("," + args.cfg if args.cfg else ""))
    t0 = time.perf_counter()
    s = W.soup(dev, args.tris, meshes=max(1, args.tris // 4_000_000))
    build_s = time.perf_counter() - t0
    st0 = dev.frame_stats()
    total_rays = 1 << args.log2_rays
    n = total_rays // world
    rays_h = W.random_rays(total_rays, tfar=float("inf"))[rank * n:(rank + 1) * n]
    rays_pin = torch.from_numpy(rays_h).pin_memory()
    rays_d = rays_pin.cuda()
    hits_d = torch.zeros((n, 8), dtype=torch.float32, device="cuda")
    hits_pin = torch.empty((n, 8), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    nbar = None
    if rank == 0:
        sdev = Device.cuda(cfg=f"gpu={local_rank},stats=1")
        ss = W.soup(sdev, args.tris, meshes=max(1, args.tris // 4_000_000))
        m = min(n, 1 << 22)
        sdev.trace_rays_device(ss.scene, rays_d.data_ptr(), hits_d.data_ptr(), m, True)
        sst = sdev.frame_stats()
        nbar = {"nodes_per_ray": sst.node_visits / m, "tris_per_ray": sst.tri_tests / m, "num_triangles": int(sst.num_triangles), "num_nodes": int(sst.num_nodes)}
        sdev.close()
    for _ in range(args.warmup):
        dev.trace_rays_device(s.scene, rays_d.data_ptr(), hits_d.data_ptr(), n, True)
    barrier()
    ms = 0.0
    with ClockSampler(local_rank) as clocks:
        for i in range(args.steps):
            flush.fill_(i & 255); torch.cuda.synchronize()
            ms += dev.trace_rays_device(s.scene, rays_d.data_ptr(), hits_d.data_ptr(), n, True)
        barrier()
    clk = clocks.summary()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):                              # e2e: host rays -> device, trace, hits -> host
        rays_d.copy_(rays_pin, non_blocking=True); torch.cuda.synchronize()
        dev.trace_rays_device(s.scene, rays_d.data_ptr(), hits_d.data_ptr(), n, True)
        hits_pin.copy_(hits_d, non_blocking=True); torch.cuda.synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    vals = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms_t, e2e_t = vals[0].item(), vals[1].item()
        value = total_rays * args.steps / (ms_t * 1e-3) / 1e6
        peak, peak_src = measured_peak()
        bpr = 64 + 80 * nbar["nodes_per_ray"] + 48 * nbar["tris_per_ray"]
        gbs = n * args.steps * bpr / (ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(REPO, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(f"c5_{args.tris}")
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_t / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"C5 synthetic soup: {args.tris} triangles ({nbar['num_nodes']} BVH8 nodes, {(nbar['num_nodes'] * 80 + nbar['num_triangles'] * 48) / 1e6:.0f} MB), "
                                       f"2^{args.log2_rays} incoherent closest-hit rays per step split over {world} rank(s), scene replicated",
                           "l2": "a 256 MiB buffer is rewritten between timed steps (L2 flush)"},
                "e2e": {"value": total_rays * args.steps / e2e_t / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": n * 32},
                "gpu_launches": args.steps * world, "clocks": clk, "traversal": nbar, "build_s_first_commit": build_s, "bvh_build_ms": st0.build_ms,
                "roofline": {"bound": "hbm", "kernel": "k_trace_user", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": traffic,
                             "peak_source": peak_src, "bytes_per_ray": bpr, "avg_launch_ms": ms / args.steps}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda")
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cfg", default="", help="extra device cfg keys (development A/B only)")
    ap.add_argument("--size", type=int, default=0, help="override the face resolution (profiling runs only; not a bench line)")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (profiling runs only; not a bench line)")
    ap.add_argument("--per-face", action="store_true", help="the reference's literal loop: 12 x (update, commit, rtRenderFrame) instead of yrtxRenderCubeMap (A/B)")
    ap.add_argument("--no-stats", action="store_true", help="skip the stats=1 replay (profiling runs under ncu; rooflines then print zeros)")
    ap.add_argument("--tris", type=int, default=10_000_000, help="c5: triangles in the soup")
    ap.add_argument("--log2-rays", type=int, default=24, help="c5: log2 of the rays per step")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "c5":
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "c5 (raw traversal sweep) has no reference-side driver; the CPU arm covers c1..c4"}))
            return 0
        return run_soup(args, rank, world, local_rank)
    _, desc, size, spp, depth, faces = WORKLOADS[args.workload]
    if args.size or args.spp:
        size = args.size or size; spp = args.spp or spp
        desc += f" [OVERRIDDEN for profiling: {size}x{size}, {spp} spp - not the benchmark configuration]"
    config = {"workload": desc, "step": f"one stereo cube map = {faces} faces" if faces > 1 else "one frame",
              "faces_per_cube_map": faces, "face_px": size, "spp": spp, "max_depth": depth,
              "partition": f"4-row bands round-robin over {world} rank(s), scene replicated, {'12 x rtRenderFrame' if args.per_face else 'all faces in one wavefront (yrtxRenderCubeMap)'}",
              "l2": "a 256 MiB buffer is rewritten between timed steps (L2 flush); wavefront state per step (>7 GB) also exceeds L2"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference(args.workload, args.steps, max(1, min(args.warmup, 2)))
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["e2e_value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "s_per_stereo_cube_map_extrapolated": r["s_per_stereo_cube_map_extrapolated"], "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from yulio_raytracer_b200 import Device
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # --gpus N without torchrun (one process): the group device (cfg gpus=N, csrc/group_api.cu) spreads every face over N GPUs behind the
    # one Device; under torchrun (the driver's launch) every rank owns one GPU and the bands are exchanged over NCCL.
    in_process = args.gpus if (world == 1 and args.gpus > 1) else 1
    if in_process > 1:
        config["partition"] = f"4-row bands round-robin over {in_process} GPUs of one process (group device), scene replicated"
    dev = Device.cuda(cfg=(f"gpus={in_process}" if in_process > 1 else f"gpu={local_rank},serverID={rank},serverCount={world}") + ("," + args.cfg if args.cfg else ""))
    s = build_workload(dev, args.workload, size, spp, depth, "RGB8")
    dev.set_option("rebuild", 1); dev.rtCommit(s.scene); dev.set_option("rebuild", 0)   # a second, full commit: the GPU BVH build without the first call's pool growth
    st_build = dev.frame_stats()                           # CUDA-event time of the build (Morton sort, PLOC, SAH-optimal BVH8 collapse)
    cams = make_cameras(dev, s, faces)
    fbs = [s.framebuffer] + [dev.rtNewFrameBuffer("RGB8", size, size, 1) for _ in range(len(cams) - 1)]
    stride = (3 * size + 3) // 4 * 4
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    # ---- measured roofline denominators (csrc/microbench.cu), rank 0 ----
    peaks = {}
    peaks["hbm_gbs"], peaks["hbm_source"] = measured_peak()
    if rank == 0:
        try:
            peaks["fp32_tflops_measured"] = dev.microbench(0)
            peaks["l2_gbs_measured"] = dev.microbench(1, 48 << 20)
            peaks["hbm_read_gbs_measured"] = dev.microbench(1, 4 << 30)
            peaks["issue_gwarpinst_s_measured"] = dev.microbench(2)
        except Exception as e:                             # a group device answers through member 0; anything else is reported, not hidden
            peaks["microbench_error"] = str(e)

    # ---- multi-GPU gather of the row bands (the only inter-GPU traffic; NCCL over NVLink): one all-gather per cube map ----
    class _DevView:                      # zero-copy view of the device framebuffer for torch
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    fb_t = gatherer = full_host = None
    if world > 1:
        fb_t = []
        for fb in fbs:
            ptr, nbytes, _ = dev.framebuffer_device(fb)
            fb_t.append(torch.as_tensor(_DevView(ptr, nbytes), device="cuda"))
        from yulio_raytracer_b200 import bands
        gatherer = bands.CubeBandGather(len(fbs), size, stride, rank, world, "cuda")
        if rank == 0:
            full_host = torch.empty(len(fbs) * size * stride, dtype=torch.uint8).pin_memory()

    def gather_bands(to_host=False):
        if world == 1:
            return 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        full = gatherer.gather(fb_t)
        e1.record()
        if to_host and rank == 0:
            full_host.copy_(full, non_blocking=True)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- traversal statistics of the workload (stats=1 replay of one cube map on a second device handle, untimed) ----
    nbar = None
    if rank == 0 and args.no_stats:
        nbar = {k: 0.0 for k in ("nodes_per_ray", "tris_per_ray", "closest_nodes_per_ray", "closest_tris_per_ray", "shadow_nodes_per_ray", "shadow_tris_per_ray")}
    elif rank == 0:
        sdev = Device.cuda(cfg=f"gpu={local_rank},stats=1,serverID={rank},serverCount={world * in_process}")
        ss = build_workload(sdev, args.workload, size, spp, depth, "RGB8")
        sdev.set_readback(False)
        scams = make_cameras(sdev, ss, faces)
        sfbs = [ss.framebuffer] + [sdev.rtNewFrameBuffer("RGB8", size, size, 1) for _ in range(len(scams) - 1)]
        render_step(sdev, ss, scams, sfbs)
        st = sdev.frame_stats()
        nrays = st.rays_closest + st.rays_shadow
        cn, ct = st.node_visits - st.node_visits_shadow, st.tri_tests - st.tri_tests_shadow
        nbar = {"nodes_per_ray": st.node_visits / max(1, nrays), "tris_per_ray": st.tri_tests / max(1, nrays),
                "closest_nodes_per_ray": cn / max(1, st.rays_closest), "closest_tris_per_ray": ct / max(1, st.rays_closest),
                "shadow_nodes_per_ray": st.node_visits_shadow / max(1, st.rays_shadow), "shadow_tris_per_ray": st.tri_tests_shadow / max(1, st.rays_shadow),
                "closest_fraction": st.rays_closest / max(1, nrays), "num_triangles": int(st.num_triangles), "num_nodes": int(st.num_nodes),
                "bvh_mb": (st.num_nodes * 80 + st.num_triangles * 48) / 1e6, "shading_records_mb": st.num_triangles * 80 / 1e6}
        sdev.close()

    # ---- (1) device-timed: frames stay in HBM ----
    dev.set_readback(False)
    for i in range(args.warmup):
        render_step(dev, s, cams, fbs, args.per_face); gather_bands()
    keys = ("rays", "ms", "closest_ms", "shadow_ms", "shade_ms", "resolve_ms", "rf_ms", "launches", "closest_rays", "shadow_rays", "closest_launches",
            "shadow_launches", "shade_launches", "gather_ms", "sort_ms", "vertices", "miss_ms")
    agg = {k: 0.0 for k in keys}
    barrier()
    with ClockSampler(local_rank) as clocks:
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.fill_(i & 255); torch.cuda.synchronize()
            if args.per_face and s.view is not None:
                rows = []
                for cam, fb in zip(cams, fbs):
                    render_step(dev, s, [cam], [fb], True); rows.append(dev.frame_stats())
            else:
                render_step(dev, s, cams, fbs); rows = [dev.frame_stats()]
            g = gather_bands()
            agg["ms"] += g; agg["gather_ms"] += g
            for st in rows:
                agg["rays"] += st.rays_closest + st.rays_shadow; agg["ms"] += st.render_ms
                agg["closest_ms"] += st.closest_ms; agg["shadow_ms"] += st.shadow_ms; agg["shade_ms"] += st.shade_ms; agg["resolve_ms"] += st.resolve_ms
                agg["rf_ms"] += st.raygen_film_ms; agg["sort_ms"] += st.sort_ms; agg["miss_ms"] += st.miss_ms
                agg["launches"] += st.kernel_launches; agg["closest_rays"] += st.rays_closest; agg["shadow_rays"] += st.rays_shadow
                agg["closest_launches"] += st.closest_launches; agg["shadow_launches"] += st.shadow_launches; agg["shade_launches"] += st.shade_launches
                agg["vertices"] += st.path_vertices
        barrier()
        wall_dev = time.perf_counter() - t_wall0
    clk = clocks.summary()

    # ---- (1b) per-kernel stage times for the rooflines: ONE cube map with one chunk lane (cfg lanes=1), so that every launch has the
    # GPU to itself. The timed region above runs two lanes: kernels of two chunks share the SMs and their CUDA-event spans overlap.
    ser = None
    if rank == 0 and not args.per_face and "lanes=2" in args.cfg:
        try:
            dev.set_option("lanes", 1)
            render_step(dev, s, cams, fbs)
            render_step(dev, s, cams, fbs); st = dev.frame_stats()
            ser = {"ms": st.render_ms, "closest_ms": st.closest_ms, "shadow_ms": st.shadow_ms, "shade_ms": st.shade_ms, "resolve_ms": st.resolve_ms,
                   "rf_ms": st.raygen_film_ms, "closest_rays": st.rays_closest, "shadow_rays": st.rays_shadow, "closest_launches": st.closest_launches,
                   "shadow_launches": st.shadow_launches, "shade_launches": st.shade_launches, "vertices": st.path_vertices,
                   "rays": st.rays_closest + st.rays_shadow}
        finally:
            dev.set_option("lanes", 2)
    if world > 1:
        barrier()

    # ---- (2) end to end through the reference-facing loop, every frame read back to the host every step ----
    dev.set_readback(world == 1)           # N > 1: the bands are gathered on the GPU first, rank 0 copies the assembled frames to the host
    for i in range(2):                                     # the same body as the timed loop (first maps allocate staging / pinned buffers)
        render_step(dev, s, cams, fbs, args.per_face)
        for fb in fbs:
            dev.rtSwapBuffers(fb); dev.rtMapFrameBuffer(fb); dev.rtUnmapFrameBuffer(fb)
        gather_bands(to_host=True)
    barrier()
    e2e_rays = 0; h2d = d2h = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        c2 = make_cameras(dev, s, faces)                   # camera creation + commit is part of the caller's per-viewpoint work
        render_step(dev, s, c2, fbs, args.per_face)
        st = dev.frame_stats()
        for fb in fbs:
            dev.rtSwapBuffers(fb); dev.rtMapFrameBuffer(fb); dev.rtUnmapFrameBuffer(fb)
        gather_bands(to_host=True)
        if args.per_face and s.view is not None:
            e2e_rays += agg["rays"] / args.steps
        else:
            e2e_rays += st.rays_closest + st.rays_shadow
        h2d += st.h2d_bytes + 4096
        # bytes that reached the host this step: a plain device copies every frame inside the render call, a group device when its frames are mapped
        d2h += (len(fbs) * size * stride) if (world > 1 and rank == 0) else (max(st.d2h_bytes, dev.frame_stats().d2h_bytes) if world == 1 else 0)
        if world == 1 and args.per_face:
            d2h += (len(fbs) - 1) * size * stride
        if s.view is not None:                             # C1 renders with the scene's one pinhole camera
            for c in c2:
                dev.rtDecRef(c)
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- reduce over ranks: max time, sum rays ----
    vals = torch.tensor([agg["ms"], e2e_s, float(agg["rays"]), float(e2e_rays), float(agg["launches"]), float(d2h), float(h2d)],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, e2e_total = mx[0].item(), mx[1].item()
        rays_total, e2e_rays_total, launches_total, d2h, h2d = sm[2].item(), sm[3].item(), sm[4].item(), sm[5].item(), sm[6].item()
    else:
        ms_total, e2e_total, rays_total, e2e_rays_total, launches_total = agg["ms"], e2e_s, agg["rays"], e2e_rays, agg["launches"]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = rays_total / (ms_total * 1e-3) / 1e6
    e2e_value = e2e_rays_total / e2e_total / 1e6
    ms_per_step = ms_total / args.steps
    src = ser if ser else agg                              # single-lane pass when there is one (exclusive kernel times)
    kr = kernel_rooflines(src, nbar, peaks, in_process)
    stage = {"k_trace_closest": src["closest_ms"], "k_trace_shadow": src["shadow_ms"], "k_shade": src["shade_ms"]}
    dominant = max(stage, key=stage.get)
    roofline = dict(kr.get(dominant, {"bound": "hbm", "kernel": dominant, "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": None, "traffic": None}))
    roofline["peak_source"] = peaks["hbm_source"]
    roofline["share_of_step"] = stage[dominant] / max(1e-9, src["ms"])
    roofline["timed_with"] = ("one cube map rendered with one chunk lane (cfg lanes=1) after the timed region: exclusive kernel times; "
                              f"that pass ran at {src['rays'] / src['ms'] / 1e3:.0f} Mrays/s") if ser else "the timed region"
    roofline["measured_peaks"] = {k: v for k, v in peaks.items() if k.endswith("_measured") or k == "microbench_error"}
    roofline["other_kernels"] = {k: {kk: v[kk] for kk in ("achieved", "frac", "frac_l2", "frac_fp32", "avg_launch_ms") if kk in v} for k, v in kr.items() if k != dominant}
    try:
        with open(os.path.join(REPO, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        roofline["traffic"] = (tj.get(args.workload) or {}).get(dominant)
        roofline["ncu"] = (tj.get(args.workload) or {}).get(dominant + "_ncu")
    except Exception:
        pass
    line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world * in_process, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config,
            "s_per_stereo_cube_map": ms_per_step * 1e-3, "e2e_s_per_stereo_cube_map": e2e_total / args.steps,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d // args.steps), "d2h_bytes_per_step": int(d2h // args.steps)},
            "gpu_launches": int(launches_total), "clocks": clk, "roofline": roofline,
            "stage_ms_per_step": {"closest": agg["closest_ms"] / args.steps, "shadow": agg["shadow_ms"] / args.steps,
                                  "shade": agg["shade_ms"] / args.steps, "resolve": agg["resolve_ms"] / args.steps, "miss": agg["miss_ms"] / args.steps,
                                  "raygen_film": agg["rf_ms"] / args.steps, "sort": agg["sort_ms"] / args.steps, "gather": agg["gather_ms"] / args.steps},
            "stage_ms_note": "sums of CUDA-event spans per kernel kind; launches are serialised on one stream (cfg lanes=1, the default)",
            "stage_ms_single_lane": ({k: ser[k] for k in ("ms", "closest_ms", "shadow_ms", "shade_ms", "resolve_ms", "rf_ms")} if ser else None),
            "bvh_build": {"ms": st_build.build_ms, "triangles": int(st_build.num_triangles), "nodes": int(st_build.num_nodes),
                          "mtris_per_s": (st_build.num_triangles / st_build.build_ms / 1e3) if st_build.build_ms > 0 else None,
                          "note": "second full scene commit of the process (CUDA-event time of build_bvh: Morton sort, PLOC, SAH-optimal BVH8 collapse)"},
            "traversal": nbar, "rays_per_step": rays_total / args.steps, "wall_s_timed_region": wall_dev}
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference(args.workload, 3, 1)
        line["cpu_baseline"] = {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                "s_per_stereo_cube_map_extrapolated": r["s_per_stereo_cube_map_extrapolated"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
