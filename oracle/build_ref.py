#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY (oracle).  Builds oracle/_ref/liboracle_singleray.so:

    the reference's own devices/device_singleray + common/{sys,image} sources, compiled where
    they lie (through the patched scratch overlay made by make_overlay.py), linked with
    oracle/embree2_shim.cpp (stand-in for the un-vendored Intel Embree 2.15 binary) and
    oracle/oracle_capi.cpp (the C symbols of include/yrt_device.h forwarded to embree::Device).

Needs /root/reference (present in the build container only). The GPU box uses the prebuilt
.so that travels with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).
Flags: -O2 -msse4.2 -fno-fast-math -ffp-contract=off for the reference sources (strict IEEE so the
oracle is reproducible; the shipped build uses fast-math); the shim adds -mfma
because its arithmetic contract YRT-PLUECKER-1 uses explicit fused multiply-adds.

A second library, oracle/_ref/liboracle_singleray_fast.so, is the BASELINE build bench.py times (cpu_baseline, --impl reference):
the reference's release flags -O3 -ffast-math (common/cmake/gcc.cmake:27) plus -msse4.2, its own approximate rcp / rsqrt (overlay
without pin P3), -fno-finite-math-only so that the infinite tfar / tMaxShadowRay the path relies on stay defined. It is never used
for parity.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("YRT_REFERENCE", "/root/reference")
OVL = os.path.join(REPO, "build", "oracle_overlay")
OBJ = os.path.join(REPO, "build", "oracle_obj")
OUT = os.path.join(HERE, "_ref")

SINGLERAY = """api/singleray_device.cpp lights/hdrilight.cpp shapes/trianglemesh_normals.cpp
shapes/trianglemesh_full.cpp samplers/sampler.cpp samplers/distribution1d.cpp samplers/distribution2d.cpp
integrators/pathtraceintegrator.cpp filters/filter.cpp renderers/debugrenderer.cpp
renderers/integratorrenderer.cpp renderers/progress.cpp""".split()
COMMON = """sys/platform.cpp sys/sysinfo.cpp sys/filename.cpp sys/library.cpp sys/thread.cpp
sys/taskscheduler.cpp sys/taskscheduler_sys.cpp sys/sync/mutex.cpp sys/sync/condition.cpp sys/stl/string.cpp
image/image.cpp image/pfm.cpp image/ppm.cpp image/tga.cpp""".split()


def main():
    if not os.path.isdir(REF):
        print("build_ref: no reference tree at", REF, "- keeping prebuilt oracle/_ref", file=sys.stderr)
        return 0 if os.path.exists(os.path.join(OUT, "liboracle_singleray.so")) else 1
    build(False)
    build(True)
    return 0


def build(fast):
    global OVL, OBJ
    OVL = os.path.join(REPO, "build", "oracle_overlay" + ("_fast" if fast else ""))
    OBJ = os.path.join(REPO, "build", "oracle_obj" + ("_fast" if fast else ""))
    subprocess.check_call([sys.executable, os.path.join(HERE, "make_overlay.py")] + (["--fast"] if fast else []))
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(OUT, exist_ok=True)
    inc = ["-I" + OVL, "-I" + os.path.join(OVL, "common"), "-I" + os.path.join(OVL, "devices"),
           "-I" + os.path.join(OVL, "devices", "device_singleray"), "-I" + HERE,
           "-I" + os.path.join(REF, "3rd party", "Embree v2.15.0 x64", "include"),
           "-I" + os.path.join(REF, "3rd party", "glm-0.9.8.4")]
    opt = ["-O3", "-ffast-math", "-fno-finite-math-only"] if fast else ["-O2", "-fno-fast-math", "-ffp-contract=off"]
    base = ["g++", "-std=c++14", "-msse4.2"] + opt + ["-fPIC", "-fpermissive", "-w", "-DNDEBUG", "-pthread"] + inc
    strict = ["g++", "-std=c++14", "-msse4.2", "-O3" if fast else "-O2", "-fno-fast-math", "-ffp-contract=off", "-fPIC", "-fpermissive", "-w", "-DNDEBUG", "-pthread"] + inc
    jobs = []
    for s in SINGLERAY:
        jobs.append((os.path.join(OVL, "devices", "device_singleray", s), base))
    for s in COMMON:
        jobs.append((os.path.join(OVL, "common", s), base))
    jobs.append((os.path.join(HERE, "embree2_shim.cpp"), strict + ["-mfma"]))     # the shim keeps its arithmetic contract in both builds
    jobs.append((os.path.join(HERE, "oracle_capi.cpp"), strict))
    jobs.append((os.path.join(HERE, "oracle_stubs.cpp"), base))

    def cc(job):
        src, flags = job
        obj = os.path.join(OBJ, os.path.relpath(src, REPO).replace("/", "_") + ".o")
        r = subprocess.run(flags + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(f"--- {src}\n{r.stderr[-4000:]}\n")
            raise SystemExit(1)
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(cc, jobs))
    so = os.path.join(OUT, "liboracle_singleray_fast.so" if fast else "liboracle_singleray.so")
    subprocess.check_call(["g++", "-shared", "-o", so] + objs + ["-ldl", "-pthread", "-Wl,--no-undefined"])
    print("built", so)


if __name__ == "__main__":
    sys.exit(main())
