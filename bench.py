#!/usr/bin/env python3
"""bench.py — the reference's headline benchmark on device_cuda: Mrays/s (path segments) and seconds per stereo
cube map, driven through the C-ABI (include/yrt_device.h) exactly as the reference's outputMode loop drives a device
(devices/renderer/renderer.cpp:543-632).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c1] [--impl reference]

A "step" is one rtRenderFrame of one stereo cube face of the workload (BASELINE.json configs[1] by default: the
sphere_glass scene, 1024x1024 per face, 64 spp, depth 8, 12 stereo cube cameras); K steps walk the faces 0..11
cyclically, so the default K = 12 is exactly one stereo cube map. Metric = the reference's own counter and timer span:
(rtcIntersect + rtcOccluded equivalents) / render time (integratorrenderer.cpp:99-111, pathtraceintegrator.cpp:74,161).

  value   device-timed (CUDA events on the device's stream around the wavefront loop), frame left in HBM
  e2e     host wall-clock of the reference-facing loop per face: rtUpdatePrimitive x prims, rtCommit(scene),
          rtRenderFrame, rtSwapBuffers, rtMapFrameBuffer -> the frame is in the (pinned) host buffer; host->device
          bytes = sample table / constants actually uploaded, device->host bytes = the frame
  N > 1   (without torchrun: one process, the group device of csrc/group_api.cu, cfg gpus=N)
          one process per GPU (torchrun, the driver's launch); the scene is replicated, every face is split into the reference's 4-row
          bands dealt round-robin to the ranks (api/swapchain.h:57-70, the reference's own network-device partition) and
          the bands are gathered on rank 0 over NCCL; strong scaling (total work fixed).
--impl reference runs the reference's own CPU path (oracle/_ref: devices/device_singleray sources + embree2 shim) on all
host cores on a bounded sample (same scene, camera, spp and depth at a reduced face resolution).
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
    # name: (builder, description, face size, spp, depth)
    "c2": ("spheres", "C2 sphere_glass.xml + sphere_view.ecs: stereo cube face 1024x1024, 64 spp, depth 8 (procedural lines texture stands in for lines.ppm)", 1024, 64, 8),
    "c3": ("atrium", "C3 stand-in (Sponza.DAE stripped): procedural atrium ~276k tris, Uber+alpha+dome light, stereo cube face 1024x1024, 64 spp, depth 10, tMaxShadowRay 120", 1024, 64, 10),
    "c4": ("atrium4", "C4 stand-in (22 Frederick St .dae stripped): procedural atrium ~1.1M tris, Uber+alpha+thin glass+billboard+dome light, stereo cube face 2048x2048, 16 spp (rt_test_dll.cpp:17), depth 10, tMaxShadowRay 120", 2048, 16, 10),
    "c1": ("cornell", "C1 cornell_box.ecs: pinhole 512x512, 16 spp, depth 2", 512, 16, 2),
}


def build_workload(dev, name, size, spp, depth, fmt):
    from tests import scenes
    kind = WORKLOADS[name][0]
    if kind == "spheres":
        return scenes.spheres(dev, "glass", size, size, spp, depth, face=0, fmt=fmt)
    if kind == "atrium":
        return scenes.atrium(dev, size, size, spp, depth, face=0, detail=56, fmt=fmt)
    if kind == "atrium4":
        return scenes.atrium(dev, size, size, spp, depth, face=0, detail=112, fmt=fmt, tex_size=512)
    s = scenes.cornell(dev, size, size, spp, depth, fmt=fmt)
    s.view = None
    return s


def face_camera(dev, s, face):
    from tests import scenes
    if s.view is None:
        return s.camera
    pos, target, up = s.view
    return scenes.stereo_camera(dev, face % 12, pos, target, up)


def render_face(dev, s, cam, update=True):
    """One iteration of the outputMode loop (renderer.cpp:551-580)."""
    if update and s.view is not None:
        org = dev.rtGetFloat3(cam, "origin")
        for j, p in enumerate(s.prims):
            dev.rtUpdatePrimitive(s.scene, j, p, org, s.view[2])
        dev.rtCommit(s.scene)
    dev.rtRenderFrame(s.renderer, cam, s.scene, s.tonemapper, s.framebuffer, 0)


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
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, n in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peak():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference(workload, steps, warmup, sample_size):
    """The reference CPU path on the host cores, bounded sample: same scene/camera/spp/depth, face resolution sample_size."""
    from oracle import oracle_device
    _, desc, size, spp, depth = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    dev = oracle_device.open_oracle(num_threads=cores)
    s = build_workload(dev, workload, sample_size, spp, depth, "RGB8")
    rays = secs = wall = 0.0
    for i in range(warmup + steps):
        cam = face_camera(dev, s, i)
        t0 = time.perf_counter()
        render_face(dev, s, cam)
        dev.rtSwapBuffers(s.framebuffer); dev.rtMapFrameBuffer(s.framebuffer); dev.rtUnmapFrameBuffer(s.framebuffer)
        dt = time.perf_counter() - t0
        st = dev.frame_stats()
        if i >= warmup:
            rays += st.rays_closest; secs += st.render_ms * 1e-3; wall += dt
    sample = (f"reference devices/device_singleray + embree2-API shim (NOT Intel Embree), {cores} threads, {steps} faces of the same "
              f"scene/camera/spp/depth at {sample_size}x{sample_size} px per face instead of {size}x{size}")
    return {"value": rays / secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference", "sample": sample,
            "e2e_value": rays / wall / 1e6, "ms_per_step": secs * 1e3 / steps, "rays": rays, "wall_s": wall}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=256, help="face resolution of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cfg", default="", help="extra device cfg keys (development A/B only)")
    ap.add_argument("--size", type=int, default=0, help="override the face resolution (profiling runs only; not a bench line)")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (profiling runs only; not a bench line)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _, desc, size, spp, depth = WORKLOADS[args.workload]
    if args.size or args.spp:
        size = args.size or size; spp = args.spp or spp
        desc += f" [OVERRIDDEN for profiling: {size}x{size}, {spp} spp - not the benchmark configuration]"
    config = {"workload": desc, "faces_per_cube_map": 12, "partition": f"4-row bands round-robin over {world} rank(s), scene replicated",
              "l2": "a 256 MiB buffer is rewritten between timed steps (L2 flush); wavefront state per step (>1 GB) also exceeds L2"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference(args.workload, args.steps, max(1, min(args.warmup, 1)), args.cpu_sample)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["e2e_value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
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
    stride = (3 * size + 3) // 4 * 4
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    # ---- multi-GPU gather of the row bands (the only inter-GPU traffic; NCCL over NVLink) ----
    class _DevView:                      # zero-copy view of the device framebuffer for torch
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    fb_t = gatherer = None
    if world > 1:
        ptr, nbytes, _ = dev.framebuffer_device(s.framebuffer)
        fb_t = torch.as_tensor(_DevView(ptr, nbytes), device="cuda")
        from yulio_raytracer_b200 import bands
        gatherer = bands.BandGather(size, stride, rank, world, "cuda")

    def gather_bands():
        if world == 1:
            return 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gatherer.gather(fb_t)
        e1.record(); e1.synchronize()
        return e0.elapsed_time(e1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- traversal statistics of the workload (stats=1 replay of one face on a second device handle, untimed) ----
    nbar = None
    if rank == 0:
        sdev = Device.cuda(cfg=f"gpu={local_rank},stats=1,serverID={rank},serverCount={world * in_process}")
        ss = build_workload(sdev, args.workload, size, spp, depth, "RGB8")
        sdev.set_readback(False)
        render_face(sdev, ss, face_camera(sdev, ss, 0))
        st = sdev.frame_stats()
        nrays = st.rays_closest + st.rays_shadow
        nbar = {"nodes_per_ray": st.node_visits / max(1, nrays), "tris_per_ray": st.tri_tests / max(1, nrays),
                "closest_fraction": st.rays_closest / max(1, nrays), "num_triangles": int(st.num_triangles), "num_nodes": int(st.num_nodes)}
        sdev.close()

    # ---- (1) device-timed: frame stays in HBM ----
    dev.set_readback(False)
    cams = [face_camera(dev, s, f) for f in range(12)]
    for i in range(args.warmup):
        render_face(dev, s, cams[i % 12]); gather_bands()
    agg = {"rays": 0, "ms": 0.0, "closest_ms": 0.0, "shadow_ms": 0.0, "shade_ms": 0.0, "rf_ms": 0.0, "launches": 0, "closest_rays": 0,
           "shadow_rays": 0, "closest_launches": 0, "shadow_launches": 0, "gather_ms": 0.0, "sort_ms": 0.0}
    barrier()
    with ClockSampler(local_rank) as clocks:
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.fill_(i & 255); torch.cuda.synchronize()
            render_face(dev, s, cams[i % 12])
            st = dev.frame_stats()
            g = gather_bands()
            agg["rays"] += st.rays_closest + st.rays_shadow; agg["ms"] += st.render_ms + g; agg["gather_ms"] += g
            agg["closest_ms"] += st.closest_ms; agg["shadow_ms"] += st.shadow_ms; agg["shade_ms"] += st.shade_ms; agg["rf_ms"] += st.raygen_film_ms; agg["sort_ms"] += st.sort_ms
            agg["launches"] += st.kernel_launches; agg["closest_rays"] += st.rays_closest; agg["shadow_rays"] += st.rays_shadow
            agg["closest_launches"] += st.closest_launches; agg["shadow_launches"] += st.shadow_launches
        barrier()
        wall_dev = time.perf_counter() - t_wall0
    clk = clocks.summary()

    # ---- (2) end to end through the reference-facing loop, frame read back to the host every step ----
    dev.set_readback(True)
    for i in range(2):
        render_face(dev, s, cams[i % 12]); dev.rtSwapBuffers(s.framebuffer); dev.rtMapFrameBuffer(s.framebuffer); dev.rtUnmapFrameBuffer(s.framebuffer)
    barrier()
    e2e_rays = 0; h2d = d2h = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        cam = face_camera(dev, s, i)                       # camera creation + commit is part of the caller's per-face work
        render_face(dev, s, cam)
        dev.rtSwapBuffers(s.framebuffer)
        p = dev.rtMapFrameBuffer(s.framebuffer); dev.rtUnmapFrameBuffer(s.framebuffer)
        gather_bands()
        st = dev.frame_stats()
        e2e_rays += st.rays_closest + st.rays_shadow; h2d += st.h2d_bytes + 4096; d2h += st.d2h_bytes
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- reduce over ranks: max time, sum rays ----
    vals = torch.tensor([agg["ms"], e2e_s, float(agg["rays"]), float(e2e_rays), float(agg["launches"]), agg["closest_ms"], agg["shadow_ms"]],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, e2e_total = mx[0].item(), mx[1].item()
        rays_total, e2e_rays_total, launches_total = sm[2].item(), sm[3].item(), sm[4].item()
    else:
        ms_total, e2e_total, rays_total, e2e_rays_total, launches_total = agg["ms"], e2e_s, agg["rays"], e2e_rays, agg["launches"]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = rays_total / (ms_total * 1e-3) / 1e6
    e2e_value = e2e_rays_total / e2e_total / 1e6
    ms_per_step = ms_total / args.steps
    peak, peak_src = measured_peak()
    # dominant kernel: closest-hit traversal. Algorithmic bytes/ray = 32 (ray in) + 16 (hit out: t,u,v,triangle index — SURVEY §8d planned
    # 32, the record shrank when the shading data moved into per-triangle records) + 80*Nnode + 48*Ntri
    bpr = 48 + 80 * nbar["nodes_per_ray"] + 48 * nbar["tris_per_ray"]
    per_gpu_closest = agg["closest_rays"] / in_process      # the roofline is one GPU's (group device: the counters are sums over its GPUs)
    achieved = per_gpu_closest * bpr / (agg["closest_ms"] * 1e-3) / 1e9 if agg["closest_ms"] > 0 else None
    roofline = {"bound": "hbm", "kernel": "k_trace_closest", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src,
                "bytes_per_ray": bpr, "avg_launch_ms": agg["closest_ms"] / max(1, agg["closest_launches"]),
                "note": "BVH of this workload is L2-resident; node/triangle bytes are served by L1/L2, not HBM (DESIGN.md)"}
    # secondary roofline: FP32 issue. flops/ray = 190*Nnode + 90*Ntri against 148 SMs x 128 lanes x 2 flop x SM clock
    flops_per_ray = 190 * nbar["nodes_per_ray"] + 90 * nbar["tris_per_ray"]
    fp32_peak = 148 * 128 * 2 * (clk["sm_mhz"] or 1965.0) * 1e6 / 1e12
    fp32_ach = per_gpu_closest * flops_per_ray / (agg["closest_ms"] * 1e-3) / 1e12 if agg["closest_ms"] > 0 else None
    roofline["fp32"] = {"achieved": fp32_ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": (fp32_ach / fp32_peak) if fp32_ach else None,
                        "flops_per_ray": flops_per_ray, "peak_source": "nominal: 148 SMs x 128 FP32 lanes x 2 x sampled SM clock"}
    try:
        with open(os.path.join(REPO, "profiles", "traffic.json")) as f:
            roofline["traffic"] = json.load(f).get(args.workload)
    except Exception:
        pass
    line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world * in_process, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config,
            "s_per_stereo_cube_map": ms_per_step * 12e-3, "e2e_s_per_stereo_cube_map": e2e_total / args.steps * 12,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps},
            "gpu_launches": int(launches_total), "clocks": clk, "roofline": roofline,
            "stage_ms_per_step": {"closest": agg["closest_ms"] / args.steps, "shadow": agg["shadow_ms"] / args.steps,
                                  "shade": agg["shade_ms"] / args.steps, "raygen_film": agg["rf_ms"] / args.steps, "sort": agg["sort_ms"] / args.steps,
                                  "gather": agg["gather_ms"] / args.steps},
            "traversal": nbar, "rays_per_step": rays_total / args.steps, "wall_s_timed_region": wall_dev}
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference(args.workload, 4, 1, args.cpu_sample)
        line["cpu_baseline"] = {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
